#!/bin/bash
# memcheck of the kept-feature path + 2-rank bench and phases
set -u
mkdir -p gpurun_out
TAG=${1:-r02u}
L=gpurun_out/final_${TAG}.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== compute-sanitizer memcheck (kept features tests)" > $L
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --timeout=800 \
  -k "kept_feature or split_covariance" > gpurun_out/memcheck_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -5 gpurun_out/memcheck_${TAG}.log >> $L
echo "== bench config2 N=2" >> $L
timeout 240 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --no-cpu \
  > gpurun_out/bench_${TAG}_n2.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n2.log | cut -c1-1500 >> $L
echo "== multi-rank test" >> $L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=300 >> $L 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | cut -c1-1600 | tail -40
