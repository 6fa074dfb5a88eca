#!/bin/bash
# round 2, run E (2 GPUs): config5 script smoke at reduced N, e2e fix check, GLM launch list
set -u
mkdir -p gpurun_out
TAG=${1:-r02e}
L=gpurun_out/final_${TAG}.log
echo "== config5 smoke on 2 GPUs (N=2e6)" > $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
  --master-port 29512 bench.py --workload config5 --gpus 2 --N 2000000 --Ks 512,2048 --steps 2 \
  > gpurun_out/config5_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep "^{" gpurun_out/config5_${TAG}.log >> $L
tail -3 gpurun_out/config5_${TAG}.log >> $L
echo "== bench config2 N=1" >> $L
timeout 900 python bench.py --no-cpu > gpurun_out/bench_${TAG}_n1.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n1.log >> $L
echo "== glm launch list" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/launches_glm_${TAG}.csv python scripts/glm_step_timing.py \
  > gpurun_out/glm_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== gpu tests (subset)" >> $L
timeout 900 python -m pytest tests -m gpu -q -x -k "multi or partition or fused_suffstats or i8_selftest" >> $L 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | cut -c1-1500 | tail -40
