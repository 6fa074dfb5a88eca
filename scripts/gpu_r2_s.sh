#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02s}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests (kept features)" > $L
timeout 900 python -m pytest tests -m gpu -q --timeout=300 \
  -k "kept or config2_posterior or slm_elbo or slm_fit or pipelined or ill_conditioned or mid_size" >> $L 2>&1; echo "rc=$?" >> $L
echo "== breakdown" >> $L
timeout 300 python scripts/keep_breakdown.py > gpurun_out/keep_breakdown_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -2 gpurun_out/keep_breakdown_${TAG}.log >> $L
echo "== bench" >> $L
timeout 300 python bench.py --no-cpu > gpurun_out/bench_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}.log | cut -c1-1800 >> $L
grep -v "^$" $L | cut -c1-2600 | tail -60; exit 0
echo "== launch list (one evaluation kept, one regenerated)" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_keep_${TAG}.csv python scripts/keep_breakdown.py --once \
  > gpurun_out/ncu_keep_${TAG}.log 2>&1; echo "rc=$?" >> $L
python scripts/launch_summary.py gpurun_out/launches_keep_${TAG}.csv >> $L 2>&1
grep -v "^$" $L | cut -c1-2600 | tail -80
