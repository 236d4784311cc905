mkdir -p gpurun_out
timeout 200 python tools/variant_sweep.py base --streams 8 --slots 8388608 --cases cbox_glossy:1024:64 > gpurun_out/c2_slots.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gpt_bounce -s 12 -c 1 -o gpurun_out/r01b_bounce python tools/gpt_sweep.py cbox_glossy:1024:4 > gpurun_out/c2_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gpt_generate -s 12 -c 1 -o gpurun_out/r01b_generate python tools/gpt_sweep.py cbox_glossy:1024:4 >> gpurun_out/c2_ncu.log 2>&1
cat gpurun_out/c2_slots.log; tail -5 gpurun_out/c2_ncu.log; ls -la gpurun_out
