#!/bin/bash
# One profiling pass on a B200 (run under gpurun): launch list + plan, per-layer report, stage profile, and one `ncu --set full`
# capture each of the dominant GEMM variant, the d=40 attention kernel and the GroupNorm apply kernel.  Outputs -> gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
MADM_DUMP_PLAN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/launches.csv python tools/ncu_step.py > $O/step.log 2> $O/plan.log
python tools/layer_report.py $O/launches.csv $O/plan.log > $O/gemm_layers.txt
python tools/stage_profile.py 8 > $O/stage.txt 2>&1
for spec in "gemm256:regex:gemm_tc_kernel<256:3" "gemm128:regex:gemm_tc_kernel<128, 2:2" "gemm160:regex:gemm_tc_kernel<160:4" "fa40:regex:fa_tc_kernel<40:1" "gn_apply:regex:gn_apply:4"; do
  name=${spec%%:*}; rest=${spec#*:}; kern=${rest%:*}; cnt=${rest##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$kern" -c $cnt -o $O/full_$name \
    python tools/ncu_step.py > /dev/null 2>&1
done
ls -la $O
