#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02g}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1800 python -m pytest tests -m gpu -q -x -s >> $L 2>&1; echo "rc=$?" >> $L
echo "== quick" >> $L
timeout 600 python scripts/r2_quick.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== glm bench" >> $L
timeout 600 python bench.py --workload config4 --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_glm_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}.log >> $L
grep -v "^$" $L | grep -v "^E  \|^    " | cut -c1-1200 | tail -40
