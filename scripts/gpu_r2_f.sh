#!/bin/bash
# round 2, run F: full GPU tests after the tf32x3 GEMM / xtq change, GLM bench
set -u
mkdir -p gpurun_out
TAG=${1:-r02f}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1800 python -m pytest tests -m gpu -q -x >> $L 2>&1; echo "rc=$?" >> $L
echo "== glm bench" >> $L
timeout 600 python bench.py --workload config4 --steps 30 --warmup 5 > gpurun_out/bench_glm_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}.log >> $L
echo "== glm launch list" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/launches_glm_${TAG}.csv python scripts/glm_step_timing.py \
  > gpurun_out/glm_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | cut -c1-1200 | tail -30
