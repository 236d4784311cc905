mkdir -p gpurun_out
SPP=64 timeout 60 python tools/e2e_phases.py > gpurun_out/c14_e2e_phases.log 2>&1
cat gpurun_out/c14_e2e_phases.log
