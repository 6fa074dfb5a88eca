#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02m}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests (glm + fit)" > $L
timeout 1800 python -m pytest tests -m gpu -q -x --timeout=300 -k "glm or fit or pipelined" >> $L 2>&1; echo "rc=$?" >> $L
echo "== solve variants" >> $L
timeout 600 python scripts/solve_variants.py 2>&1 | grep "batched, leaf\|full solve\|value-only solve\|blocked_spd" >> $L; echo "rc=$?" >> $L
echo "== midsize fit debug" >> $L
timeout 600 python scripts/fit_midsize_debug.py 2>&1 | grep -v "^coord" | cut -c1-400 >> $L; echo "rc=$?" >> $L
echo "== glm bench" >> $L
timeout 600 python bench.py --workload config4 --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_glm_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}.log >> $L
REVRAND_B200_GLM_DEVICE_GRAPH=0 timeout 600 python bench.py --workload config4 --steps 100 --warmup 10 --no-cpu > gpurun_out/bench_glm_${TAG}_nograph.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}_nograph.log | cut -c1-300 >> $L
echo "== bench fit" >> $L
timeout 900 python bench.py --workload fit > gpurun_out/bench_fit_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_fit_${TAG}.log >> $L
grep -v "^$" $L | cut -c1-2500 | tail -120
