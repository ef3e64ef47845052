#!/bin/bash
# A/B on one box: 32-bit vs 64-bit per-thread ray counters (ab_u32.so / ab_u64.so built from the two sources).
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_az.txt; : > $S
L=raytracer_go_b200/csrc
cp $L/librt_b200.so /tmp/keep.so
for rep in 1 2; do for v in u64 u32; do
  cp $L/ab_$v.so $L/librt_b200.so
  for cfg in C2 CB C3; do
    timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$cfg', round(d['value'],1), 'Msamples/s')" | tee -a $S
  done
done; done
cp /tmp/keep.so $L/librt_b200.so
