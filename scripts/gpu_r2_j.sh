#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02j}
L=gpurun_out/final_${TAG}.log
echo "== fit debug" > $L
timeout 600 python scripts/fit_debug.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== solve variants" >> $L
timeout 600 python scripts/solve_variants.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== e2e breakdown" >> $L
timeout 600 python scripts/e2e_breakdown.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench fit" >> $L
timeout 900 python bench.py --workload fit > gpurun_out/bench_fit_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_fit_${TAG}.log >> $L
echo "== gpu tests" >> $L
timeout 1800 python -m pytest tests -m gpu -q -x >> $L 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | cut -c1-1500 | tail -120
