mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpt_gpu.py -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
timeout 300 python tools/variant_sweep.py base rmw > gpurun_out/c5_variants.log 2>&1
timeout 200 python tools/variant_sweep.py base --streams 8 --slots 1048576,8388608 --cases cbox_glossy:1024:64 >> gpurun_out/c5_variants.log 2>&1
tail -5 gpurun_out/c5_pytest.log; cat gpurun_out/c5_variants.log
