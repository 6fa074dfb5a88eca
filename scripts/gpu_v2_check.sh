#!/bin/bash
# Staged GPU check of the CTA-pair fused value pass; each stage under its own
# timeout so a hung kernel cannot eat the lease.
set -u
mkdir -p gpurun_out
L=gpurun_out/v2_check.log
echo "== fused suffstats small" > $L
timeout 180 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_suffstats" >> $L 2>&1; rc=$?; echo "rc=$rc" >> $L
if [ $rc -ne 0 ] && [ $rc -ne 1 ]; then tail -30 $L; exit 0; fi
echo "== full size properties" >> $L
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "full_size" >> $L 2>&1; echo "rc=$?" >> $L
echo "== all gpu tests" >> $L
timeout 900 python -m pytest tests -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
echo "== breakdown" >> $L
timeout 300 python scripts/eval_breakdown.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | tail -150
