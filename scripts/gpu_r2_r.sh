#!/bin/bash
# kept-feature gradient pass: parity + A/B timing
set -u
mkdir -p gpurun_out
TAG=${1:-r02r}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests (kept features)" > $L
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x \
  -k "kept or config2 or mid_size or concatenated or polynomial_basis_rides or finite_differences" >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench config2 kept" >> $L
timeout 300 python bench.py --no-cpu > gpurun_out/bench_${TAG}_keep.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_keep.log | cut -c1-2500 >> $L
echo "== bench config2 regenerate" >> $L
REVRAND_B200_KEEP_FEATURES_MAX_GB=0 timeout 300 python bench.py --no-cpu > gpurun_out/bench_${TAG}_regen.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_regen.log | cut -c1-2500 >> $L
echo "== phases" >> $L
timeout 240 python bench.py --workload config5 --N 1000000 --Ks 2048 --steps 4 > gpurun_out/phases_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep '^{' gpurun_out/phases_${TAG}.log | cut -c1-1500 >> $L
grep -v "^$" $L | cut -c1-2600 | tail -80
