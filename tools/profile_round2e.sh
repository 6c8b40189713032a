#!/bin/bash
# Final-state captures of the round's training kernels: `ncu --set full` of the tcgen05 attention backward (d = 40 both roles, d = 80 dK/dV role) and
# the fused LoRA gradient kernel, plus the launch list of one training pass.  Outputs -> gpurun_out/r02g_*.
O=gpurun_out
mkdir -p $O
cap() {  # name kernel-regex launch-skip
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$2" --launch-skip $3 -c 1 \
    -o $O/r02g_full_$1 python tools/profile_train.py ncu > /dev/null 2>&1
}
cap attn_bwd_tc_d40_dkv attn_bwd_tc_kernel 0
cap attn_bwd_tc_d40_dq attn_bwd_tc_kernel 1
cap attn_bwd_tc_d80_dkv attn_bwd_tc_kernel 6
cap lora_grad_64x64 lora_grad_kernel 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02g_train_launches.csv \
  python tools/profile_train.py ncu > /dev/null 2>&1
python tools/launch_summary.py $O/r02g_train_launches.csv 45 > $O/r02g_train_launch_summary.txt
python tools/ncu_summary.py $O/r02g_full_*.ncu-rep > $O/r02g_ncu_full_summary.txt 2>&1
rm -f $O/r02g_full_*.ncu-rep
grep -E "^==|duration|tensor pipe cycles|XU|issue slots|regs" $O/r02g_ncu_full_summary.txt
