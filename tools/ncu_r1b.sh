#!/bin/bash
# ncu --set full captures of the three tcgen05 kernels in one cs_frame step (B = 8, 512 px); run on the GPU box.
set -x
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
$NCU -k regex:conv3s_tc_kernel -s 0 -c 2 -o gpurun_out/prof_conv3s_r1b python tools/profile_step.py 8 > gpurun_out/ncu1.log 2>&1
$NCU -k "regex:conv_tc_kernel<false, false, 2" -s 10 -c 1 -o gpurun_out/prof_conv_tc_adaptive_r1b python tools/profile_step.py 8 > gpurun_out/ncu2.log 2>&1
$NCU -k "regex:conv_tc_kernel<false, true, 2, true" -s 1 -c 1 -o gpurun_out/prof_conv_tc_spade_r1b python tools/profile_step.py 8 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
