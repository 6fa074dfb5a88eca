#!/bin/bash
# End-of-round evidence: all GPU tests, smoke, bench, ncu launch list and full
# captures of the two tcgen05 kernels.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 900 python -m pytest tests -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$?" >> $L
echo "== breakdown" >> $L
timeout 300 python scripts/eval_breakdown.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench ours" >> $L
timeout 900 python bench.py > gpurun_out/bench_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}.log >> $L
echo "== launch list" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu \
  > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== ncu full: value pass" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc2_suffstats -s 1 -c 1 \
  -f -o gpurun_out/prof_suffstats_${TAG} python scripts/run_suffstats.py 1000000 2 \
  > gpurun_out/prof_suffstats_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== ncu full: gradient pass" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp2_kernel -s 20 -c 1 \
  -f -o gpurun_out/prof_gp2_${TAG} python scripts/eval_breakdown.py \
  > gpurun_out/prof_gp2_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | tail -40
