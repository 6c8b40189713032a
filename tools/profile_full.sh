#!/bin/bash
# `ncu --set full` captures of single launches (ordinals from the launch list of tools/profile_round.sh): the 256-row x 128 VAE
# conv tile, the 128 x 256 tile, the 128 x 160 UNet tile, and the d = 40 self-attention launch.  Outputs -> gpurun_out/.
O=gpurun_out
mkdir -p $O
cap() {  # name kernel-regex launch-skip
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$2" --launch-skip $3 -c 1 \
    -o $O/full_$1 python tools/ncu_step.py > /dev/null 2>&1
}
cap gemm128x2 gemm_tc_kernel 2
cap gemm256 gemm_tc_kernel 7
cap gemm160 gemm_tc_kernel 172
cap fa40 fa_tc_kernel 0
ls -la $O/*.ncu-rep
