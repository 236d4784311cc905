mkdir -p gpurun_out
timeout 35 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c15_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c15_smoke.log
tail -3 gpurun_out/c15_smoke.log
