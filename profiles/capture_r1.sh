#!/bin/bash
# ncu captures of one bench-workload step (c2).  Run under gpurun; outputs land in gpurun_out/.
# usage: [PATS="regex1 regex2"] bash profiles/capture_r1.sh <tag> [mode]
TAG=${1:-r1}
MODE=${2:-bf16}
CMD="python profiles/profile_step.py --mode $MODE"
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
PATS=${PATS:-"lstm_step|vocab_sample conv_pool|head_fwd|head_bwd|dz_fused|clip_adam gemm_p_kernel|gemm_tf32_kernel|gemm_pair_kernel"}
for pat in $PATS; do
  name=$(echo "$pat" | tr -c 'a-zA-Z0-9' '_' | cut -c1-24)
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$pat" -c 10 \
      -f -o gpurun_out/prof_${TAG}_$name $CMD > gpurun_out/ncu_full_${TAG}_$name.log 2>&1
done
ls -la gpurun_out/*${TAG}*.ncu-rep
