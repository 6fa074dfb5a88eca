#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02k}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1800 python -m pytest tests -m gpu -q -x >> $L 2>&1; echo "rc=$?" >> $L
echo "== solve variants" >> $L
timeout 600 python scripts/solve_variants.py 2>&1 | grep -v "^rank\|^   " >> $L; echo "rc=$?" >> $L
echo "== e2e breakdown" >> $L
timeout 600 python scripts/e2e_breakdown.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench fit" >> $L
timeout 900 python bench.py --workload fit > gpurun_out/bench_fit_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_fit_${TAG}.log >> $L
echo "== bench config2 N=1" >> $L
timeout 900 python bench.py --no-cpu > gpurun_out/bench_${TAG}_n1.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n1.log >> $L
grep -v "^$" $L | cut -c1-2500 | tail -120
