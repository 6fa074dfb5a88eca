#!/bin/bash
# round 2, run I (1 GPU): all GPU tests, fit trace, config-2 bench
set -u
mkdir -p gpurun_out
TAG=${1:-r02i}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1800 python -m pytest tests -m gpu -q -s >> $L 2>&1; echo "rc=$?" >> $L
echo "== fit debug" >> $L
timeout 600 python scripts/fit_debug.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench config2 N=1" >> $L
timeout 900 python bench.py --no-cpu > gpurun_out/bench_${TAG}_n1.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n1.log >> $L
grep -v "^$" $L | grep -v "^E  \|^    " | cut -c1-1500 | tail -80
