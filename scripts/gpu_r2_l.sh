#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02l}
L=gpurun_out/final_${TAG}.log
echo "== solve variants" > $L
timeout 600 python scripts/solve_variants.py 2>&1 | grep -v "^rank\|^   [a-z]" >> $L; echo "rc=$?" >> $L
echo "== midsize fit debug" >> $L
timeout 600 python scripts/fit_midsize_debug.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== gpu tests (slm)" >> $L
timeout 1800 python -m pytest tests -m gpu -q -x -k "slm or fit or pipelined" >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench config2 N=1" >> $L
timeout 900 python bench.py --no-cpu > gpurun_out/bench_${TAG}_n1.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-1400 >> $L
grep -v "^$" $L | cut -c1-1500 | tail -120
