mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpt_gpu.py -m gpu -x -q > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c12_pytest.log
timeout 300 python bench.py > gpurun_out/c12_bench_n1.json 2> gpurun_out/c12_bench_n1.err
tail -4 gpurun_out/c12_pytest.log; cat gpurun_out/c12_bench_n1.json; tail -3 gpurun_out/c12_bench_n1.err
