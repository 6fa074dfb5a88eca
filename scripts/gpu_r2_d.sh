#!/bin/bash
# round 2, run D (2 GPUs): all GPU tests incl. the 2-rank equality test, GLM config-4
# bench + launch list, config-2 bench at N=1 and N=2
set -u
mkdir -p gpurun_out
TAG=${1:-r02d}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1800 python -m pytest tests -m gpu -q -x -s >> $L 2>&1; echo "rc=$?" >> $L
echo "== glm bench" >> $L
timeout 600 python bench.py --workload config4 --steps 20 --warmup 5 > gpurun_out/bench_glm_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_${TAG}.log >> $L
timeout 300 python bench.py --workload config4 --impl reference --steps 2 > gpurun_out/bench_glm_ref_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_glm_ref_${TAG}.log >> $L
echo "== glm launch list" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv \
  --log-file gpurun_out/launches_glm_${TAG}.csv python scripts/glm_step_timing.py \
  > gpurun_out/glm_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== bench config2 N=1" >> $L
timeout 900 python bench.py --no-cpu > gpurun_out/bench_${TAG}_n1.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n1.log >> $L
echo "== bench config2 N=2" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
  --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_${TAG}_n2.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}_n2.log >> $L
grep -v "^$" $L | tail -80
