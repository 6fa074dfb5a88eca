#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02a3}
timeout 300 python scripts/keep_breakdown.py --reps 6 > gpurun_out/keep_breakdown_${TAG}.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/keep_breakdown_${TAG}.log
