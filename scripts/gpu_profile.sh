#!/bin/bash
# Per-phase timing, ncu launch list of one bench run and full-set captures of
# the two tcgen05 kernels.  Outputs -> gpurun_out/ (summaries are copied into
# profiles/ by hand).
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 300 python scripts/eval_breakdown.py > gpurun_out/breakdown_${TAG}.json 2> gpurun_out/breakdown_${TAG}.err
echo "breakdown rc=$?"; tail -1 gpurun_out/breakdown_${TAG}.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu \
  > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc2_suffstats -s 1 -c 1 \
  -f -o gpurun_out/prof_suffstats_${TAG} python scripts/eval_breakdown.py \
  > gpurun_out/prof_suffstats_${TAG}.log 2>&1
echo "suffstats capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gradpass -s 300 -c 1 \
  -f -o gpurun_out/prof_gradpass_${TAG} python scripts/eval_breakdown.py \
  > gpurun_out/prof_gradpass_${TAG}.log 2>&1
echo "gradpass capture rc=$?"
ls -la gpurun_out
