#!/bin/bash
# Round-2 (final build) profiles of one cs_frame step (B = 8, 512 px); run on the GPU box:  bash tools/ncu_r2b.sh
# Kernels are selected by their MANGLED names (template arguments are unambiguous there).
set -x
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f --kernel-name-base mangled"
# Winograd GEMM of an adaptive conv (pair mode): first launch of the swap stage
timeout 300 $NCU -k "regex:conv_tc_kernelILb0ELb0ELi2ELb0ELb0E" -s 0 -c 1 -o gpurun_out/prof_r2b_wino_gemm python tools/profile_step.py 8 > gpurun_out/ncu_r2b_a.log 2>&1
# SPADE gamma|beta conv with the modulation epilogue
timeout 300 $NCU -k "regex:conv_tc_kernelILb0ELb1ELi2ELb1ELb0E" -s 1 -c 1 -o gpurun_out/prof_r2b_spade python tools/profile_step.py 8 > gpurun_out/ncu_r2b_b.log 2>&1
# the depth-stacked 3x3x3 kernel: emit / residual+emit variants after the epilogue clean-up
timeout 300 $NCU -k "regex:conv3s_tc_kernel" -s 2 -c 2 -o gpurun_out/prof_r2b_conv3s python tools/profile_step.py 8 > gpurun_out/ncu_r2b_c.log 2>&1
# hourglass input, occlusion gather, Winograd input transform with the mask conv
timeout 300 $NCU -k "regex:dm_input_operand_kernel|occlusion_gather_kernel|wino_in_kernelILb1E" -s 0 -c 3 -o gpurun_out/prof_r2b_hbm python tools/profile_step.py 8 > gpurun_out/ncu_r2b_d.log 2>&1
# launch list of the whole step: duration + DRAM bytes per launch
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv python tools/profile_step.py 8 > gpurun_out/ncu_r2b_e.log 2>&1
ls -la gpurun_out/*r2b*
