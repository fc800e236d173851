#!/bin/bash
# Round-2 ncu captures.  Run under gpurun; outputs land in gpurun_out/.
#   1. launch list (gpu__time_duration.sum) of bench.py's own command line (the first launches: the eager step that precedes capture)
#   2. launch list of ONE eager c2 step (profiles/profile_step.py, between cudaProfilerStart/Stop)
#   3. --set full of the top kernels of that step
TAG=${1:-r2}
MODE=${2:-bf16}
BENCH="python bench.py --steps 3 --warmup 3 --no-modes --no-workloads --no-sustained --no-cpu-baseline"
$BENCH > gpurun_out/bench_plain_$TAG.log 2>&1 || { echo "bench plain run failed"; tail -5 gpurun_out/bench_plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_$TAG.csv $BENCH > gpurun_out/ncu_bench_$TAG.log 2>&1
CMD="python profiles/profile_step.py --mode $MODE"
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
PATS=${PATS:-"vocab_sample|lstm_step gemm_pair_kernel|gemm_p_kernel conv_pool|head_fwd|head_bwd|dz_fused|clip_adam"}
for pat in $PATS; do
  name=$(echo "$pat" | tr -c 'a-zA-Z0-9' '_' | cut -c1-24)
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$pat" -c 12 \
      -f -o gpurun_out/prof_${TAG}_$name $CMD > gpurun_out/ncu_full_${TAG}_$name.log 2>&1
  # gpurun merges at most 64 MiB back: keep the raw-metric table of every capture, drop the report itself
  ncu -i gpurun_out/prof_${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$name.raw.csv 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$name.ncu-rep
done
ls -la gpurun_out/*${TAG}*
