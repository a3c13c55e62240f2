mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 3 > gpurun_out/tc_check.log 2>&1
echo "rc=$?" >> gpurun_out/tc_check.log
