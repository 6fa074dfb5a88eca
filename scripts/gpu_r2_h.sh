#!/bin/bash
# round 2, run H (1 GPU): all GPU tests, quick value-pass check, config-2 bench, fit bench
set -u
mkdir -p gpurun_out
TAG=${1:-r02h}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1800 python -m pytest tests -m gpu -q -x -s >> $L 2>&1; echo "rc=$?" >> $L
echo "== quick" >> $L
timeout 600 python scripts/r2_quick.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench config2 N=1" >> $L
timeout 900 python bench.py --no-cpu > gpurun_out/bench_${TAG}_n1.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n1.log >> $L
echo "== bench fit" >> $L
timeout 900 python bench.py --workload fit > gpurun_out/bench_fit_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -3 gpurun_out/bench_fit_${TAG}.log >> $L
grep -v "^$" $L | grep -v "^E  \|^    " | cut -c1-1500 | tail -60
