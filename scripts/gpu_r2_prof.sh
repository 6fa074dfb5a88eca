#!/bin/bash
# round 2 evidence (1 GPU): launch lists, DRAM traffic of one value pass, ncu --set full
# captures of the four tensor-core kernels, smoke
set -u
mkdir -p gpurun_out
TAG=${1:-r02final}
L=gpurun_out/prof_${TAG}.log
echo "== smoke" > $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$?" >> $L
echo "== breakdown" >> $L
timeout 300 python scripts/eval_breakdown.py >> $L 2>&1; echo "rc=$?" >> $L
echo "== launch list (config-2 bench, 2 timed steps)" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu \
  > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== DRAM traffic of the value pass (2 passes)" >> $L
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none -k regex:t3_ --csv --log-file gpurun_out/traffic_vp_${TAG}.csv \
  python scripts/run_suffstats.py 1000000 2 > gpurun_out/traffic_vp_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== ncu full: int8 SYRK" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t3_syrk -s 6 -c 1 \
  -f -o gpurun_out/prof_syrk_${TAG} python scripts/run_suffstats.py 1000000 3 \
  > gpurun_out/prof_syrk_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== ncu full: digit generator" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t3_digits -s 6 -c 1 \
  -f -o gpurun_out/prof_digits_${TAG} python scripts/run_suffstats.py 1000000 3 \
  > gpurun_out/prof_digits_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== ncu full: gradient GEMM" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp2_kernel -s 3 -c 1 \
  -f -o gpurun_out/prof_gp2_${TAG} python scripts/run_gradpass.py \
  > gpurun_out/prof_gp2_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== ncu full: tf32x3 GEMM (GLM step, F = Phi Ws^T)" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:g3_gemm -s 8 -c 1 \
  -f -o gpurun_out/prof_g3_${TAG} python scripts/glm_step_timing.py \
  > gpurun_out/prof_g3_${TAG}.log 2>&1; echo "rc=$?" >> $L
echo "== glm launch list" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 250 --csv \
  --log-file gpurun_out/launches_glm_${TAG}.csv python scripts/glm_step_timing.py \
  > gpurun_out/glm_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
grep -v "^$" $L | cut -c1-600 | tail -40
