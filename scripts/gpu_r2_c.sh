#!/bin/bash
# round 2, run C: new generator + new bases; GLM step baseline timing
set -u
mkdir -p gpurun_out
TAG=${1:-r02c}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1500 python -m pytest tests -m gpu -q -x >> $L 2>&1; echo "rc=$?" >> $L
echo "== quick" >> $L
timeout 600 python scripts/r2_quick.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== breakdown" >> $L
timeout 300 python scripts/eval_breakdown.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== glm step" >> $L
timeout 300 python scripts/glm_step_timing.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== launch list of 2 value passes" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_vp_${TAG}.csv python scripts/run_suffstats.py 1000000 2 \
  > gpurun_out/vp_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | tail -60
