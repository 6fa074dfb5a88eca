#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp2_kernel -s 20 -c 1 \
  -f -o gpurun_out/prof_gp2_${TAG} python scripts/eval_breakdown.py \
  > gpurun_out/prof_gp2_${TAG}.log 2>&1
echo "gp2 capture rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gp2_kernel|phi_err|prep_c|absmax" -c 1200 --csv \
  --log-file gpurun_out/launches_gp2_${TAG}.csv python scripts/eval_breakdown.py > /dev/null 2>&1
echo "launch list rc=$?"
