mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=60 --tb=short 2>&1 | tail -15 > gpurun_out/t25.log
timeout 600 python bench.py > gpurun_out/bench25.log 2>gpurun_out/bench25.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r1_final.csv python tools/profile_step.py 8 > gpurun_out/ncu25a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 15 -c 2 -f -o gpurun_out/prof_conv_tc_r1_final python tools/profile_step.py 8 > gpurun_out/ncu25b.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv7_tc_kernel -c 1 -f -o gpurun_out/prof_conv7_r1_final python tools/profile_step.py 8 > gpurun_out/ncu25c.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv3s_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_conv3s_r1_final python tools/profile_step.py 8 > gpurun_out/ncu25d.log 2>&1
