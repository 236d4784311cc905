mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/c11_bench_n8.json 2> gpurun_out/c11_bench_n8.err
cat gpurun_out/c11_bench_n8.json; tail -5 gpurun_out/c11_bench_n8.err
