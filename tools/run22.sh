mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=60 --tb=short 2>&1 | tail -30 > gpurun_out/t22.log
timeout 400 python bench.py --steps 6 --warmup 3 --cpu-frames 0 > gpurun_out/bench22.log 2>gpurun_out/bench22.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1g.csv python tools/profile_step.py 8 > gpurun_out/ncu_launches22.log 2>&1
timeout 400 python tools/e2e_err.py 256 > gpurun_out/e2e_err22.log 2>&1
