#!/bin/bash
# round 2, 8-GPU run: 2-rank equality test, config-2 bench at 8 / 4 / 2 ranks, per-phase
# breakdown at 8 ranks, config-5 sweep (N = 1e7, K = 512 .. 8192)
set -u
mkdir -p gpurun_out
TAG=${1:-r02x8}
L=gpurun_out/final_${TAG}.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > $L
for n in 8 4; do
  echo "== bench config2 N=$n" >> $L
  timeout 240 $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --no-cpu \
    > gpurun_out/bench_${TAG}_n$n.log 2>&1; echo "rc=$?" >> $L
  tail -1 gpurun_out/bench_${TAG}_n$n.log >> $L
done
echo "== phases at 8 ranks (config-2 shape: N=1e6, K=2048)" >> $L
timeout 240 $TR --nproc-per-node 8 --master-port 29531 bench.py --workload config5 --gpus 8 \
  --N 1000000 --Ks 2048 --steps 4 > gpurun_out/phases_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep '^{' gpurun_out/phases_${TAG}.log >> $L
echo "== config5 sweep, N=1e7, 8 ranks" >> $L
timeout 600 $TR --nproc-per-node 8 --master-port 29532 bench.py --workload config5 --gpus 8 \
  --N 10000000 --Ks 512,1024,2048,4096,8192 --steps 3 > gpurun_out/config5_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep '^{' gpurun_out/config5_${TAG}.log >> $L
grep -v "^$" $L | cut -c1-1800 | tail -60
