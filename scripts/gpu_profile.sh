#!/bin/bash
# ncu launch list of one bench run + full-set capture of the fused value-pass
# kernel.  Outputs -> gpurun_out/ (copy summaries into profiles/).
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
KREGEX=${2:-tc_suffstats}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu \
  > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 1 -c 1 \
  -f -o gpurun_out/prof_${TAG} python scripts/eval_breakdown.py \
  > gpurun_out/prof_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
