mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 72 > gpurun_out/tc_check14.log 2>&1
timeout 300 python tools/tc_check.py 72 pair > gpurun_out/tc_check14_pair.log 2>&1
