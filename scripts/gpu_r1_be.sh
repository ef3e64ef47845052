#!/bin/bash
# C2 / Cornell: SAH traversal-cost sweep under the final kernels (the slab test got cheaper since the last sweep).
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_be.txt; : > $S
for ct in 1.2 0.8 1.8 2.5 1.2; do
  for cfg in C2; do
  RT_B200_BVH_CTRAV=$ct timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('ctrav $ct $cfg', round(d['value'],1), 'Msamples/s box/ray', round(r['box_tests_per_ray'],2), 'sph/ray', round(r['sphere_tests_per_ray'],2))" | tee -a $S
  done
done
