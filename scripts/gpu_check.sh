#!/bin/bash
# First-contact GPU run: each stage under its own timeout so that a hung
# kernel in one stage does not eat the whole lease.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== selftest" | tee gpurun_out/check.log
timeout 120 python -c "
import __graft_entry__ as g; g.build()
from revrand_b200 import _engine
print('selftest err', _engine.tcgen05_selftest())
" >> gpurun_out/check.log 2>&1; echo "rc=$?" >> gpurun_out/check.log
echo "== simt + api tests" >> gpurun_out/check.log
timeout 900 python -m pytest tests -m gpu -q -k "not auto_engine and not fused and not full_size and not finite_diff and not fit_config1 and not glm_fit and not selftest" >> gpurun_out/check.log 2>&1; echo "rc=$?" >> gpurun_out/check.log
echo "== tcgen05 tests" >> gpurun_out/check.log
timeout 900 python -m pytest tests -m gpu -q -k "auto_engine or fused or full_size or finite_diff or fit_config1 or glm_fit" >> gpurun_out/check.log 2>&1; echo "rc=$?" >> gpurun_out/check.log
echo "== smoke" >> gpurun_out/check.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/check.log 2>&1; echo "rc=$?" >> gpurun_out/check.log
echo "== bench" >> gpurun_out/check.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/check.log
tail -5 gpurun_out/bench.log >> gpurun_out/check.log
tail -120 gpurun_out/check.log
