mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
timeout 600 python bench.py > gpurun_out/c8_bench_n1.json 2> gpurun_out/c8_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/c8_bench_ref.json 2> gpurun_out/c8_bench_ref.err
timeout 300 python tools/band_probe.py cbox_glossy:1024:64 --world 8 --streams 8 > gpurun_out/c8_bands.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 1 --warmup 1 --spp 4 > gpurun_out/c8_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gpt_bounce -s 12 -c 1 -o gpurun_out/r01c_bounce python tools/gpt_sweep.py cbox_glossy:1024:4 > gpurun_out/c8_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gpt_generate -s 12 -c 1 -o gpurun_out/r01c_generate python tools/gpt_sweep.py cbox_glossy:1024:4 >> gpurun_out/c8_ncu.log 2>&1
tail -4 gpurun_out/c8_pytest.log; cat gpurun_out/c8_bench_n1.json gpurun_out/c8_bench_ref.json gpurun_out/c8_bands.log; tail -3 gpurun_out/c8_bench_n1.err
