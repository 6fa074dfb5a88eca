#!/bin/bash
# Quick GPU loop: value-pass parity tests, then the per-phase breakdown.
set -u
mkdir -p gpurun_out
L=gpurun_out/quick.log
echo "== fused suffstats + full size" > $L
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_suffstats or full_size or auto_engine or finite_diff" >> $L 2>&1; echo "rc=$?" >> $L
echo "== breakdown" >> $L
timeout 300 python scripts/eval_breakdown.py >> $L 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | tail -40
