mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 72 > gpurun_out/tc_check13.log 2>&1
timeout 900 python tools/e2e_err.py 256 > gpurun_out/e2e_err13.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=line 2>&1 | tail -20 > gpurun_out/t13.log
timeout 300 python tools/tc_check.py 72 pair > gpurun_out/tc_check13_pair.log 2>&1
