#!/bin/bash
# Round-2 (late) profiles of the tcgen05 kernels after the epilogue clean-up; run on the GPU box:  bash tools/ncu_r2b.sh
set -x
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f --kernel-name-base demangled"
# Winograd GEMM of an adaptive conv (pair mode), SPADE gamma|beta conv, refine Winograd GEMM: one launch each
$NCU -k "regex:conv_tc_kernel<false, false, 2, false, false>" -s 4 -c 1 -o gpurun_out/prof_r2b_wino_gemm python tools/profile_step.py 8 > gpurun_out/ncu_r2b_a.log 2>&1
$NCU -k "regex:conv_tc_kernel<false, true, 2, true, false>" -s 2 -c 2 -o gpurun_out/prof_r2b_spade python tools/profile_step.py 8 > gpurun_out/ncu_r2b_b.log 2>&1
$NCU -k "regex:conv3s_tc_kernel" -s 2 -c 2 -o gpurun_out/prof_r2b_conv3s python tools/profile_step.py 8 > gpurun_out/ncu_r2b_c.log 2>&1
ls -la gpurun_out/*r2b*
