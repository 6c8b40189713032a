#!/bin/bash
# `ncu --set full` captures of single launches of the s0 variant's decoder stage (ordinals from the launch list of tools/profile_s0.sh;
# ncu's -k matches the function name without template arguments): the 512^2 x 256-channel upsample conv (gemm ordinal 286), a 512^2 x
# 128-channel ResBlock conv2 with the identity K segment (ordinal 290) and the largest 16-bit GroupNorm apply (ordinal 106).
O=gpurun_out
mkdir -p $O
cap() {  # name kernel launch-skip
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$2" --launch-skip $3 -c 1 \
    -f -o $O/full_s0_$1 python tools/ncu_step.py --variant s0 > $O/ncu_full_s0_$1.log 2>&1
}
cap up_conv256 gemm_tc_kernel 286
cap res_conv128 gemm_tc_kernel 290
cap gn_apply16 gn_apply_kernel 106
ls -la $O/*.ncu-rep
