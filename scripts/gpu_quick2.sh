#!/bin/bash
# GPU loop for the gradient pass: small parity tests first (each stage under its
# own timeout so that a hung kernel cannot eat the lease), then the breakdown.
set -u
mkdir -p gpurun_out
L=gpurun_out/quick2.log
echo "== slm auto engine + finite diff" > $L
timeout 240 python -m pytest tests/test_gpu_parity.py -q -x -k "auto_engine or finite_diff" >> $L 2>&1; rc=$?; echo "rc=$rc" >> $L
if [ $rc -ne 0 ] && [ $rc -ne 1 ]; then grep -v "^$" $L | tail -30; exit 0; fi
echo "== breakdown" >> $L
timeout 300 python scripts/eval_breakdown.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== all gpu tests" >> $L
timeout 900 python -m pytest tests -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | tail -60
