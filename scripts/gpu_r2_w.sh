#!/bin/bash
# A/B: kept feature image vs regenerated features at config-5 per-rank sizes (1 GPU)
set -u
mkdir -p gpurun_out
TAG=${1:-r02w}
L=gpurun_out/final_${TAG}.log
: > $L
for mode in keep regen keep regen; do
  echo "== $mode" >> $L
  if [ $mode = regen ]; then export REVRAND_B200_KEEP_FEATURES_MAX_GB=0; else unset REVRAND_B200_KEEP_FEATURES_MAX_GB; fi
  timeout 600 python bench.py --workload config5 --N 1250000 --Ks 512,2048,4096 --steps 3 \
    > gpurun_out/ab_${TAG}_$mode.log 2>&1; echo "rc=$?" >> $L
  grep '^{' gpurun_out/ab_${TAG}_$mode.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config']['workload'][:70], 'ms/eval %.2f' % d['ms_per_step'], {k: round(v, 2) for k, v in d['phases_ms'].items()})
" >> $L
done
cat $L
