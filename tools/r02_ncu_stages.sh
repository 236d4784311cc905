#!/usr/bin/env bash
# ncu of one steady-state tick of the staged wavefront (13 launches: 2 compactions, 7 stage kernels, 2 casts + the tick's
# queue lengths on stderr), 1 M resident slots: --set full capture, then the launch list of a whole short render.
set -u
mkdir -p gpurun_out
TICK=${TICK:-40}
SKIP=$((1 + 11 * TICK))
GDB200_PRINT_QUEUES=$TICK timeout 1200 ncu --set full --clock-control none --import-source on -s $SKIP -c 11 -f -o gpurun_out/r02_stages \
    env GDB200_SWEEP_STREAMS=8 GDB200_SWEEP_SLOTS=1048576 python tools/gpt_sweep.py cbox_glossy:1024:16 > gpurun_out/r02_ncu_stages.log 2>&1
grep -h "gdb200 tick" gpurun_out/r02_ncu_stages.log | head -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --spp 16 > gpurun_out/r02_bench_launches.log 2>&1
tail -2 gpurun_out/r02_bench_launches.log | cut -c1-300
