mkdir -p gpurun_out
timeout 900 python tools/e2e_err.py 128 > gpurun_out/e2e_err.log 2>&1
