mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 3 > gpurun_out/tc_check11.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-frames 0 > gpurun_out/bench11.log 2>gpurun_out/bench11.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python tools/profile_step.py 8 > gpurun_out/ncu_launches11.log 2>&1
timeout 900 python tools/e2e_err.py 256 > gpurun_out/e2e_err256.log 2>&1
