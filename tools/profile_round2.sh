#!/bin/bash
# Round-2 profiling pass on a B200 (run under gpurun): launch list + per-layer GEMM report + stage profile of the base step, and one
# `ncu --set full` capture each of the final CTA-pair GEMM variants, the d = 40 attention kernel, GroupNorm apply and the two largest
# backward kernels.  Outputs -> gpurun_out/r02p_*.
set -u
O=gpurun_out
mkdir -p $O
MADM_DUMP_PLAN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02p_launches.csv python tools/ncu_step.py > $O/r02p_step.log 2> $O/r02p_plan.log
python tools/layer_report.py $O/r02p_launches.csv $O/r02p_plan.log > $O/r02p_gemm_layers.txt
python tools/launch_summary.py $O/r02p_launches.csv 40 > $O/r02p_launch_summary.txt
python tools/stage_profile.py 8 > $O/r02p_stage.txt 2>&1
for spec in "gemm256pair:regex:gemm_tc_kernel<256, 1, 0, 1:3" "gemm160pair:regex:gemm_tc_kernel<160, 1, 0, 1:6" "gemm128x2pair:regex:gemm_tc_kernel<128, 2, 1, 1:2" "fa40:regex:fa_tc_kernel<40:1" "gn_apply:regex:gn_apply:4"; do
  name=${spec%%:*}; rest=${spec#*:}; kern=${rest%:*}; cnt=${rest##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$kern" -c $cnt -o $O/r02p_full_$name \
    python tools/ncu_step.py > /dev/null 2>&1
done
for spec in "attn_bwd_dkv:regex:attn_bwd_dkv_kernel:2" "gn_bwd_apply:regex:gn_bwd_apply:3"; do
  name=${spec%%:*}; rest=${spec#*:}; kern=${rest%:*}; cnt=${rest##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$kern" -c $cnt -o $O/r02p_full_$name \
    python tools/profile_train.py ncu > /dev/null 2>&1
done
python tools/ncu_summary.py $O/r02p_full_*.ncu-rep > $O/r02p_ncu_full_summary.txt 2>&1
ls -la $O | tail -20
