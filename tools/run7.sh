mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 3 > gpurun_out/tc_check7.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=line 2>&1 | tail -40 > gpurun_out/t7.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench7.log 2>gpurun_out/bench7.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python tools/profile_step.py 8 > gpurun_out/ncu_launches7.log 2>&1
