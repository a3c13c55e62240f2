mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 0 72 > gpurun_out/tc_check12.log 2>&1
timeout 900 python tools/e2e_err.py 256 > gpurun_out/e2e_err256b.log 2>&1
