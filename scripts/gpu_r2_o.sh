#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02o}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 >> $L 2>&1; echo "rc=$?" >> $L
echo "== glm bench" >> $L
timeout 600 python bench.py --workload config4 --steps 200 --warmup 10 > gpurun_out/bench_glm_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}.log >> $L
echo "== glm launch list" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 250 --csv \
  --log-file gpurun_out/launches_glm_${TAG}.csv python scripts/glm_step_timing.py \
  > gpurun_out/glm_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
python scripts/launch_summary.py gpurun_out/launches_glm_${TAG}.csv 2>&1 | head -14 >> $L
echo "== bench config2" >> $L
timeout 900 python bench.py --no-cpu > gpurun_out/bench_${TAG}_n1.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-1300 >> $L
grep -v "^$" $L | cut -c1-2500 | tail -120
