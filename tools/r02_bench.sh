#!/usr/bin/env bash
# Round-2: full bench (both arms) + GPU test suite on one box.
set -u
mkdir -p gpurun_out
{
  echo "== bench gdb200"; timeout 900 python bench.py --steps ${STEPS:-3} --warmup 3
  echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1
  if [ -n "${TESTS:-}" ]; then echo "== gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8; fi
} > gpurun_out/r02_bench.log 2>&1
cat gpurun_out/r02_bench.log | cut -c1-3000
