mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 120 > gpurun_out/tc_check17.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=line 2>&1 | tail -20 > gpurun_out/t17.log
timeout 900 python tools/e2e_err.py 256 > gpurun_out/e2e_err17.log 2>&1
timeout 900 python tools/e2e_err.py 128 > gpurun_out/e2e_err17_128.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-frames 0 > gpurun_out/bench17.log 2>gpurun_out/bench17.err
