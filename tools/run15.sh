mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 72 > gpurun_out/tc_check15.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=line 2>&1 | tail -20 > gpurun_out/t15.log
timeout 900 python tools/e2e_err.py 256 > gpurun_out/e2e_err15.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-frames 0 > gpurun_out/bench15.log 2>gpurun_out/bench15.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python tools/profile_step.py 8 > gpurun_out/ncu_launches15.log 2>&1
