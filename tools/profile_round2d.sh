#!/bin/bash
# `ncu --set full` captures of single launches addressed by their ordinal among the gemm_tc_kernel / fa_tc_kernel / attn_bwd_tc_kernel launches (ordinals from
# the launch list + plan dump of tools/profile_round2c.sh): the TMA-store epilogue variants, the d = 80 attention kernel, the tcgen05 attention backward.
O=gpurun_out
mkdir -p $O
cap() {  # name kernel-regex launch-skip script...
  local name=$1 kern=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$kern" --launch-skip $skip -c 1 \
    -o $O/r02f_full_$name "$@" > /dev/null 2>&1
}
cap gemm160_tma_red_to_out gemm_tc_kernel 60 python tools/ncu_step.py
cap gemm192_tma_h16_qkv gemm_tc_kernel 59 python tools/ncu_step.py
cap gemm128x2_tma_geglu gemm_tc_kernel 63 python tools/ncu_step.py
cap gemm128x2_tma_stats_conv_in gemm_tc_kernel 0 python tools/ncu_step.py
cap gemm128x2_tma_stats_vae_conv gemm_tc_kernel 1 python tools/ncu_step.py
cap gemm160_tma_stats_unet_conv gemm_tc_kernel 204 python tools/ncu_step.py
cap fa80_self fa_tc_kernel 4 python tools/ncu_step.py
cap attn_bwd_tc_dkv attn_bwd_tc_kernel 0 python tools/profile_train.py ncu
cap attn_bwd_tc_dq attn_bwd_tc_kernel 1 python tools/profile_train.py ncu
python tools/ncu_summary.py $O/r02f_full_*.ncu-rep > $O/r02f_ncu_full_summary.txt 2>&1
ls -la $O/r02f_full_*.ncu-rep
rm -f $O/r02f_full_gemm*.ncu-rep $O/r02f_full_fa80*.ncu-rep  # (20 MB each: the summary travels back; the backward reports stay for the source view)
