#!/usr/bin/env bash
# Round-2 A/B on the GPU: build variants of the staged wavefront (tools/variant_sweep.py), parity first.
#   VARIANTS="base x y"  NCU_FULL=1 (one steady-state tick with --set full)  NCU_LIST=1 (launch list)
set -u
mkdir -p gpurun_out
{
  echo "== gpu parity (tracer)"; timeout 900 python -m pytest tests/test_gpt_gpu.py -q -m gpu -x 2>&1 | tail -5
  echo "== variants"; timeout 1200 python tools/variant_sweep.py ${VARIANTS:-base} --streams 8 --cases ${CASES:-cbox_glossy:1024:64}
} > gpurun_out/r02_variants.log 2>&1
cat gpurun_out/r02_variants.log
if [ -n "${NCU_LIST:-}" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 1100 --csv --log-file gpurun_out/r02_launches.csv \
      env GDB200_SWEEP_STREAMS=8 GDB200_SWEEP_SLOTS=1048576 python tools/gpt_sweep.py cbox_glossy:1024:16 > gpurun_out/r02_launches.log 2>&1
fi
if [ -n "${NCU_FULL:-}" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -s ${SKIP:-441} -c 11 -f -o gpurun_out/r02_stages \
      env GDB200_SWEEP_STREAMS=8 GDB200_SWEEP_SLOTS=1048576 python tools/gpt_sweep.py cbox_glossy:1024:16 > gpurun_out/r02_ncu_stages.log 2>&1
fi
