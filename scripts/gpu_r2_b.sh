#!/bin/bash
# round 2, run B: all GPU tests, smoke, breakdown, bench, launch list, ncu captures
set -u
mkdir -p gpurun_out
TAG=${1:-r02b}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1500 python -m pytest tests -m gpu -q -x >> $L 2>&1; echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$?" >> $L
echo "== breakdown" >> $L
timeout 300 python scripts/eval_breakdown.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench ours" >> $L
timeout 900 python bench.py > gpurun_out/bench_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}.log >> $L
echo "== launch list" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu \
  > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== ncu full: int8 SYRK" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t3_syrk -s 5 -c 1 \
  -f -o gpurun_out/prof_syrk_${TAG} python scripts/run_suffstats.py 1000000 3 \
  > gpurun_out/prof_syrk_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== ncu full: digit generator" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t3_digits -s 5 -c 1 \
  -f -o gpurun_out/prof_digits_${TAG} python scripts/run_suffstats.py 1000000 3 \
  > gpurun_out/prof_digits_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | tail -60
