set -u
mkdir -p gpurun_out
TAG=r02end3
L=gpurun_out/final_${TAG}.log
echo "== gpu tests" > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench" >> $L
timeout 600 python bench.py --no-cpu > gpurun_out/bench_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}.log >> $L
echo "== ncu full: syrk" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:t3_syrk -s 2 -c 1 \
  -f -o gpurun_out/prof_syrk_${TAG} python scripts/keep_breakdown.py --once \
  > gpurun_out/prof_syrk_${TAG}.log 2>&1; echo "rc=$?" >> $L
python scripts/ncu_summary.py gpurun_out/prof_syrk_${TAG}.ncu-rep > gpurun_out/ncu_syrk_${TAG}.txt 2>&1
grep -E "gpu__time_duration.sum|pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes_read.sum |dram__bytes_write.sum |gpu__dram_throughput" gpurun_out/ncu_syrk_${TAG}.txt >> $L
grep -v "^$" $L | cut -c1-700 | tail -16
