#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02a2}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests (kept features)" > $L
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x \
  -k "kept or split_covariance or config2_posterior or mid_size or concatenated or slm_elbo" >> $L 2>&1; echo "rc=$?" >> $L
cat $L | tail -30; exit 0
timeout 300 python scripts/keep_breakdown.py > gpurun_out/keep_breakdown_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -2 gpurun_out/keep_breakdown_${TAG}.log >> $L
echo "== bench" >> $L
timeout 300 python bench.py --no-cpu > gpurun_out/bench_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}.log | cut -c1-1800 >> $L
cat $L | cut -c1-1900 | tail -30
