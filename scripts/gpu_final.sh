#!/bin/bash
# what the driver runs at round end, on one GPU: GPU suite, smoke, default bench,
# config-4 bench
set -u
mkdir -p gpurun_out
TAG=${1:-r02end}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 >> $L 2>&1; echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench" >> $L
timeout 600 python bench.py > gpurun_out/bench_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}.log >> $L
echo "== bench config4" >> $L
timeout 600 python bench.py --workload config4 > gpurun_out/bench_glm_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}.log >> $L
grep -v "^$" $L | cut -c1-2600 | tail -30
