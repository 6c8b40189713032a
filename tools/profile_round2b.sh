#!/bin/bash
# `ncu --set full` captures of single launches of the forward step, addressed by their ordinal among the gemm_tc_kernel / fa_tc_kernel
# launches (from the launch list of tools/profile_round2.sh): the final CTA-pair GEMM variants and the d = 40 attention kernel.
O=gpurun_out
mkdir -p $O
cap() {  # name kernel-regex launch-skip
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$2" --launch-skip $3 -c 1 \
    -o $O/r02p_full_$1 python tools/ncu_step.py > /dev/null 2>&1
}
cap gemm256_pair_vae_conv gemm_tc_kernel 12
cap gemm128x2_pair_vae_conv gemm_tc_kernel 2
cap gemm160_pair_unet_conv gemm_tc_kernel 172
cap gemm160_pair_shortk_linear gemm_tc_kernel 64
cap gemm128x2_pair_geglu gemm_tc_kernel 63
cap fa40_self gemm_never_matches 0
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:fa_tc_kernel" --launch-skip 0 -c 1 \
    -o $O/r02p_full_fa40_self python tools/ncu_step.py > /dev/null 2>&1
python tools/ncu_summary.py $O/r02p_full_gemm*.ncu-rep $O/r02p_full_fa40_self.ncu-rep > $O/r02p_ncu_forward_summary.txt 2>&1
rm -f $O/r02p_full_gemm*.ncu-rep  # (20 MB each: the summary travels back, one small report stays)
ls -la $O/*.ncu-rep
