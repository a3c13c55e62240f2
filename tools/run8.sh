mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --tb=line 2>&1 | tail -40 > gpurun_out/t8.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench8.log 2>gpurun_out/bench8.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python tools/profile_step.py 8 > gpurun_out/ncu_launches8.log 2>&1
