#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02z}
L=gpurun_out/final_${TAG}.log
echo "== gpu tests (gradient pass)" > $L
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x \
  -k "kept or config2_posterior or mid_size or concatenated or polynomial_basis_rides or finite_differences or slm_elbo" >> $L 2>&1; echo "rc=$?" >> $L
for mode in keep regen; do
  echo "== $mode" >> $L
  if [ $mode = regen ]; then export REVRAND_B200_KEEP_FEATURES_MAX_GB=0; else unset REVRAND_B200_KEEP_FEATURES_MAX_GB; fi
  timeout 600 python bench.py --workload config5 --N 1250000 --Ks 2048,4096,8192 --steps 3 \
    > gpurun_out/ab_${TAG}_$mode.log 2>&1; echo "rc=$?" >> $L
  grep '^{' gpurun_out/ab_${TAG}_$mode.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config']['workload'][:70], 'ms/eval %.2f' % d['ms_per_step'], {k: round(v, 2) for k, v in d['phases_ms'].items()})
" >> $L
done
unset REVRAND_B200_KEEP_FEATURES_MAX_GB
echo "== bench config2" >> $L
timeout 300 python bench.py --no-cpu > gpurun_out/bench_${TAG}.log 2>&1; echo "rc=$?" >> $L
tail -1 gpurun_out/bench_${TAG}.log | cut -c1-400 >> $L
cat $L | cut -c1-400 | tail -40
