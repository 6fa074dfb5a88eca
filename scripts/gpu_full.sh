#!/bin/bash
# Full round-end style pass: all GPU tests, smoke, bench (ours + reference arm),
# launch list under ncu.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
L=gpurun_out/full_${TAG}.log
echo "== gpu tests" > $L
timeout 900 python -m pytest tests -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench ours" >> $L
timeout 900 python bench.py > gpurun_out/bench_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}.log >> $L
echo "== bench reference" >> $L
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_ref_${TAG}.log >> $L
echo "== launch list" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu \
  > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | tail -40
