#!/usr/bin/env bash
# One GPU call that re-establishes the state of the tree on a B200 (run through gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh'
# Rebuild every library first (python -c "import __graft_entry__ as g; g.build()") -- the built .so files travel with the snapshot.
# Everything is written under gpurun_out/; nothing here changes GPU clocks.
set -u
mkdir -p gpurun_out
{
  echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()"; echo "rc=$?"
  echo "== gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
  echo "== bench N=1"; timeout 600 python bench.py --steps 3 --warmup 3
  echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1
} > gpurun_out/round_check.log 2>&1
# launch list of one bench step (cold-cache, serialised: shares of the step only, never a bench value)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/round_check_launches.csv \
    python bench.py --steps 1 --warmup 1 --spp 32 > gpurun_out/round_check_ncu.log 2>&1
tail -40 gpurun_out/round_check.log
