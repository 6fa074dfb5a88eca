#!/bin/bash
# full GPU suite + smoke + 2-rank bench/phases
set -u
mkdir -p gpurun_out
TAG=${1:-r02v}
L=gpurun_out/final_${TAG}.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== gpu tests" > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 >> $L 2>&1; echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench config2 N=2" >> $L
timeout 240 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --no-cpu \
  > gpurun_out/bench_${TAG}_n2.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n2.log | cut -c1-1500 >> $L
echo "== phases at 2 ranks" >> $L
timeout 240 $TR --nproc-per-node 2 --master-port 29531 bench.py --workload config5 --gpus 2 \
  --N 1000000 --Ks 2048 --steps 4 > gpurun_out/phases_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep '^{' gpurun_out/phases_${TAG}.log | cut -c1-1500 >> $L
grep -v "^$" $L | cut -c1-1600 | tail -50
