#!/bin/bash
# full GPU suite, smoke, bench, re-capture of the two kernels that changed
set -u
mkdir -p gpurun_out
TAG=${1:-r02b2}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 >> $L 2>&1; echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench" >> $L
timeout 600 python bench.py > gpurun_out/bench_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}.log >> $L
for spec in "digits:t3_digits:2" "gp2:gp2_kernel:0"; do
  name=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; skip=${rest##*:}
  echo "== ncu full: $name" >> $L
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
    -f -o gpurun_out/prof_${name}_${TAG} python scripts/keep_breakdown.py --once \
    > gpurun_out/prof_${name}_${TAG}.log 2>&1; echo "rc=$?" >> $L
  python scripts/ncu_summary.py gpurun_out/prof_${name}_${TAG}.ncu-rep > gpurun_out/ncu_${name}_${TAG}.txt 2>&1
  grep -E "gpu__time_duration.sum|pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes_read.sum |dram__bytes_write.sum |gpu__dram_throughput|shared_mem_per_block" gpurun_out/ncu_${name}_${TAG}.txt >> $L
done
grep -v "^$" $L | cut -c1-2500 | tail -40
