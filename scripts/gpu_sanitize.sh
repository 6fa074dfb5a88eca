#!/bin/bash
# compute-sanitizer memcheck over one small evaluation through both tcgen05 kernels
# (smoke: N=3000, d=5, K=96) and the ragged-size value-pass tests.
set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_smoke.log 2>&1
echo "memcheck smoke rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|smoke\]" gpurun_out/sanitize_smoke.log | head -12
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 \
  python -m pytest tests/test_gpu_parity.py -x -q -k "fused_suffstats" > gpurun_out/sanitize_fused.log 2>&1
echo "memcheck fused rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" gpurun_out/sanitize_fused.log | head -12
