#!/bin/bash
# Round-2 final profiling pass on a B200 (run under gpurun) after the TMA-store epilogue, the d = 80 attention change and the tcgen05 attention
# backward: launch list + per-layer GEMM report + stage profile of the base step, launch list of one training pass, and one `ncu --set full`
# capture each of the TMA-epilogue GEMM variants, the d = 80 attention kernel and the two roles of the backward attention kernel.
# Outputs -> gpurun_out/r02f_*.
set -u
O=gpurun_out
mkdir -p $O
MADM_DUMP_PLAN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02f_launches.csv python tools/ncu_step.py > $O/r02f_step.log 2> $O/r02f_plan.log
python tools/layer_report.py $O/r02f_launches.csv $O/r02f_plan.log > $O/r02f_gemm_layers.txt
python tools/launch_summary.py $O/r02f_launches.csv 45 > $O/r02f_launch_summary.txt
python tools/stage_profile.py 8 > $O/r02f_stage.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02f_train_launches.csv \
  python tools/profile_train.py ncu > $O/r02f_train.log 2>&1
python tools/launch_summary.py $O/r02f_train_launches.csv 45 > $O/r02f_train_launch_summary.txt
for spec in "gemm160tma_pair:regex:gemm_tc_kernel<160, 1, 4, 1:4" "gemm192tma_pair:regex:gemm_tc_kernel<192, 1, 4, 1:2" "gemm128x2tma_pair:regex:gemm_tc_kernel<128, 2, 4, 1:2" \
            "gemm128x2tma_stats_pair:regex:gemm_tc_kernel<128, 2, 5, 1:2" "gemm256tma_stats_pair:regex:gemm_tc_kernel<256, 1, 5, 1:2" "fa80:regex:fa_tc_kernel<80:1"; do
  name=${spec%%:*}; rest=${spec#*:}; kern=${rest%:*}; cnt=${rest##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$kern" -c $cnt -o $O/r02f_full_$name \
    python tools/ncu_step.py > /dev/null 2>&1
done
for spec in "attn_bwd_tc_dkv:regex:attn_bwd_tc_kernel<0:1" "attn_bwd_tc_dq:regex:attn_bwd_tc_kernel<1:1"; do
  name=${spec%%:*}; rest=${spec#*:}; kern=${rest%:*}; cnt=${rest##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$kern" -c $cnt -o $O/r02f_full_$name \
    python tools/profile_train.py ncu > /dev/null 2>&1
done
python tools/ncu_summary.py $O/r02f_full_*.ncu-rep > $O/r02f_ncu_full_summary.txt 2>&1
ls -la $O | grep r02f
