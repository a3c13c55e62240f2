#!/bin/bash
# Round-2 profiles of one cs_frame step (B = 8, 512 px); run on the GPU box:  bash tools/ncu_r2.sh
set -x
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f --kernel-name-base demangled"
# the HBM-bound kernels around the convs: Winograd transforms, input transform (prep), flow-warp, hourglass input
$NCU -k "regex:wino_in_kernel|wino_out_kernel|wino_out_blend_kernel|prep_kernel_fast|softmax_flow_warp_kernel|dm_input_operand_kernel" -s 0 -c 12 -o gpurun_out/prof_hbm_kernels_r2 python tools/profile_step.py 8 > gpurun_out/ncu_r2_a.log 2>&1
# the persistent depth-stacked 3x3x3 kernel
$NCU -k "regex:conv3s_tc_kernel" -s 2 -c 2 -o gpurun_out/prof_conv3s_r2 python tools/profile_step.py 8 > gpurun_out/ncu_r2_b.log 2>&1
# launch list of the whole step: duration + DRAM bytes per launch
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py 8 > gpurun_out/ncu_r2_c.log 2>&1
ls -la gpurun_out/*r2*
