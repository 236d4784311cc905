#!/usr/bin/env bash
# ncu --set full of one steady-state tick of the staged wavefront (11 launches), 1 M resident slots.
set -u
mkdir -p gpurun_out
SKIP=${SKIP:-441}
timeout 1200 ncu --set full --clock-control none --import-source on -s $SKIP -c 11 -f -o gpurun_out/r02_stages \
    env GDB200_SWEEP_STREAMS=8 GDB200_SWEEP_SLOTS=1048576 python tools/gpt_sweep.py cbox_glossy:1024:16 > gpurun_out/r02_ncu_stages.log 2>&1
tail -5 gpurun_out/r02_ncu_stages.log
ls -la gpurun_out/
