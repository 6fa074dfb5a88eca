#!/bin/bash
# first GPU contact of round 2: int8 engine self-test, parity subset, timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python scripts/r2_quick.py > gpurun_out/r2a_quick.log 2>&1
echo "quick rc=$?" >> gpurun_out/r2a_quick.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s \
  -k "i8_selftest or fused_suffstats or partition_invariant or tcgen05_engine or fused16 or mid_size or ill_conditioned or config2_posterior or is_selected" \
  > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_quick.log
tail -30 gpurun_out/r2a_pytest.log
