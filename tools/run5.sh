mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench5.log 2>gpurun_out/bench5.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench5_ref.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py 8 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 2 -f -o gpurun_out/prof_conv_tc_r1 python tools/profile_step.py 8 > gpurun_out/ncu_full.log 2>&1
