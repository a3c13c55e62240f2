mkdir -p gpurun_out
timeout 300 python tools/tc_check.py 120 > gpurun_out/tc_check23.log 2>&1
timeout 300 python tools/tc_check.py 120 cores > gpurun_out/tc_check23_cores.log 2>&1
