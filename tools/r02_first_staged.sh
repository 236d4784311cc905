#!/usr/bin/env bash
# Round-2 GPU probe: staged wavefront vs the round-1 fused bounce kernel (parity first, then throughput per kernel family).
set -u
mkdir -p gpurun_out
{
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()"; echo "rc=$?"
  echo "== gpu parity (tracer)"; timeout 900 python -m pytest tests/test_gpt_gpu.py -q -m gpu -x 2>&1 | tail -15
  echo "== sweep staged"; GDB200_SWEEP_STREAMS=8 timeout 600 python tools/gpt_sweep.py cbox_glossy:1024:64 cbox_diffuse:512:64
  echo "== sweep fused";  GDB200_SWEEP_FUSED=1 GDB200_SWEEP_STREAMS=8 timeout 600 python tools/gpt_sweep.py cbox_glossy:1024:64
  for slots in 1048576 2097152 4194304; do
    echo "== sweep staged slots=$slots"; GDB200_SWEEP_SLOTS=$slots GDB200_SWEEP_STREAMS=8 timeout 600 python tools/gpt_sweep.py cbox_glossy:1024:64
  done
} > gpurun_out/r02_first_staged.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_first_staged_launches.csv \
    env GDB200_SWEEP_STREAMS=8 GDB200_SWEEP_SLOTS=1048576 python tools/gpt_sweep.py cbox_glossy:1024:8 > gpurun_out/r02_first_staged_ncu.log 2>&1
tail -60 gpurun_out/r02_first_staged.log
