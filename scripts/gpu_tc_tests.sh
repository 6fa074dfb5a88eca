#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/tc_tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tcgen05" > $L 2>&1; echo "rc=$?" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | tail -60
