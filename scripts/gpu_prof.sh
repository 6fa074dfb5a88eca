#!/bin/bash
# round-2 final evidence (1 GPU): launch list of the bench, DRAM traffic of one kept
# evaluation, ncu --set full captures of the kernels of the evaluation
set -u
mkdir -p gpurun_out
TAG=${1:-r02final}
L=gpurun_out/prof_${TAG}.log
echo "== launch list (config-2 bench, 2 timed steps)" > $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu \
  > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "rc=$?" >> $L
python scripts/launch_summary.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}.txt 2>&1
head -12 gpurun_out/launches_${TAG}.txt >> $L
echo "== DRAM traffic, our kernels of one kept + one regenerated evaluation" >> $L
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none -k regex:"t3_|gp2_|fit16|phi_fit" --csv --log-file gpurun_out/traffic_${TAG}.csv \
  python scripts/keep_breakdown.py --once > gpurun_out/traffic_${TAG}.log 2>&1; echo "rc=$?" >> $L
for spec in "syrk:t3_syrk:2" "digits:t3_digits:2" "gp2:gp2_kernel:0" "fit16:fit16_kernel:0"; do
  name=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; skip=${rest##*:}
  echo "== ncu full: $name" >> $L
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
    -f -o gpurun_out/prof_${name}_${TAG} python scripts/keep_breakdown.py --once \
    > gpurun_out/prof_${name}_${TAG}.log 2>&1; echo "rc=$?" >> $L
  python scripts/ncu_summary.py gpurun_out/prof_${name}_${TAG}.ncu-rep > gpurun_out/ncu_${name}_${TAG}.txt 2>&1
  grep -E "gpu__time_duration.sum|pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes_read.sum |dram__bytes_write.sum |gpu__dram_throughput" gpurun_out/ncu_${name}_${TAG}.txt >> $L
done
grep -v "^$" $L | cut -c1-300 | tail -70
