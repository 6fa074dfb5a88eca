#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02n}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests (gemm3, glm, predict)" > $L
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=300 -k "gemm or glm or predict" >> $L 2>&1; echo "rc=$?" >> $L
echo "== solve variants" >> $L
timeout 600 python scripts/solve_variants.py 2>&1 | grep "batched, leaf\|full solve\|value-only solve\|blocked_spd" >> $L; echo "rc=$?" >> $L
echo "== glm bench" >> $L
timeout 600 python bench.py --workload config4 --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_glm_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}.log >> $L
REVRAND_B200_GLM_DEVICE_GRAPH=0 timeout 600 python bench.py --workload config4 --steps 100 --warmup 10 --no-cpu > gpurun_out/bench_glm_${TAG}_nograph.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}_nograph.log | cut -c1-300 >> $L
REVRAND_B200_GLM_DEVICE_LOOP=0 timeout 600 python bench.py --workload config4 --steps 50 --warmup 10 --no-cpu > gpurun_out/bench_glm_${TAG}_host.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}_host.log | cut -c1-300 >> $L
echo "== glm launch list" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 250 --csv \
  --log-file gpurun_out/launches_glm_${TAG}.csv python scripts/glm_step_timing.py \
  > gpurun_out/glm_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
python scripts/launch_summary.py gpurun_out/launches_glm_${TAG}.csv 2>&1 | head -14 >> $L
grep -v "^$" $L | cut -c1-2500 | tail -120
