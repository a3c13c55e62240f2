mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python tools/profile_step.py 8 > gpurun_out/ncu_launches21.log 2>&1
