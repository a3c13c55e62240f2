#!/bin/bash
# Round-1 final profiles of one cs_frame step (B = 8, 512 px) with the Winograd convs; run on the GPU box.
set -x
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f --kernel-name-base demangled"
# the Winograd GEMM of an adaptive conv: conv_tc_kernel<0,0,2,0> with grid 148 -- the 11th launch of that variant is in the swap stage
$NCU -k "regex:conv_tc_kernel<\(bool\)0, \(bool\)0, \(int\)2" -s 10 -c 1 -o gpurun_out/prof_conv_tc_wino_gemm_r1c python tools/profile_step.py 8 > gpurun_out/ncu6.log 2>&1
$NCU -k "regex:wino_in_kernel|wino_out_blend_kernel" -s 0 -c 2 -o gpurun_out/prof_wino_transforms_r1c python tools/profile_step.py 8 > gpurun_out/ncu7.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r1c_launches.csv python tools/profile_step.py 8 > gpurun_out/ncu8.log 2>&1
ls -la gpurun_out/*r1c*
