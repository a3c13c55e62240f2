#!/bin/bash
# Round-1 (second half) profiles of one cs_frame step (B = 8, 512 px); run on the GPU box, results land in gpurun_out/.
set -x
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f --kernel-name-base demangled"
$NCU -k "regex:conv_tc_kernel<\(bool\)0, \(bool\)0, \(int\)2" -s 10 -c 1 -o gpurun_out/prof_conv_tc_adaptive_r1b python tools/profile_step.py 8 > gpurun_out/ncu2.log 2>&1
$NCU -k "regex:conv_tc_kernel<\(bool\)0, \(bool\)1, \(int\)2, \(bool\)1" -s 1 -c 1 -o gpurun_out/prof_conv_tc_spade_r1b python tools/profile_step.py 8 > gpurun_out/ncu3.log 2>&1
$NCU -k "regex:conv7_tc_kernel" -s 0 -c 1 -o gpurun_out/prof_conv7_r1b python tools/profile_step.py 8 > gpurun_out/ncu4.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r1b_launches.csv python tools/profile_step.py 8 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
