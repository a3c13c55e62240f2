mkdir -p gpurun_out
timeout 300 python tools/tc_check.py 120 > gpurun_out/tc_check26.log 2>&1
timeout 300 python tools/tc_check.py 120 bn128 > gpurun_out/tc_check26_bn128.log 2>&1
