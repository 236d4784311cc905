#!/usr/bin/env bash
# One GPU session: tracer A/B sweep, full GPU test suite, solver sweep vs the reference's CUDA backend.
set -u
mkdir -p gpurun_out
{
  echo "== variants"; timeout 900 python tools/variant_sweep.py ${VARIANTS:-base} --streams 8 --cases ${CASES:-cbox_glossy:1024:64}
  if [ -n "${TESTS:-}" ]; then echo "== gpu tests"; timeout 2400 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -v "^$" | tail -40; fi
  if [ -n "${SOLVER:-}" ]; then echo "== solver sweep"; timeout 1200 python tools/solver_sweep.py ${SOLVER_SIZES:-1024x1024 1920x1080 3840x2160 7680x4320}; fi
} > gpurun_out/r02_session.log 2>&1
cat gpurun_out/r02_session.log | cut -c1-1500
