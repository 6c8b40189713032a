#!/bin/bash
# Profiling pass of the s0 variant on a B200 (run under gpurun): launch list + plan, per-layer GEMM report, and `ncu --set full`
# captures of one decoder 512^2 conv tile (16-bit stream, identity K segment) and one 16-bit GroupNorm apply.  Outputs -> gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
MADM_DUMP_PLAN=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/launches_s0.csv python tools/ncu_step.py --variant s0 > $O/step_s0.log 2> $O/plan_s0.log
python tools/layer_report.py $O/launches_s0.csv $O/plan_s0.log > $O/gemm_layers_s0.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gemm_tc_kernel<128, 2" --launch-skip 26 -c 1 \
  -o $O/full_s0_dec_gemm128x2 python tools/ncu_step.py --variant s0 > $O/ncu_full_s0_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gn_apply_kernel<1>" --launch-skip 60 -c 1 \
  -o $O/full_s0_dec_gn_apply16 python tools/ncu_step.py --variant s0 > $O/ncu_full_s0_b.log 2>&1
ls -la $O | tail -8
