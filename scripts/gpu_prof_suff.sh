#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc2_suffstats -s 1 -c 1 \
  -f -o gpurun_out/prof_suffstats_${TAG} python scripts/run_suffstats.py 1000000 2 \
  > gpurun_out/prof_suffstats_${TAG}.log 2>&1
echo "suffstats capture rc=$?"; tail -3 gpurun_out/prof_suffstats_${TAG}.log
