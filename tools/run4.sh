mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 3 > gpurun_out/tc_check4.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=line 2>&1 | tail -40 > gpurun_out/t4.log
