#!/bin/bash
# C4 (1 M spheres, scene in global memory): leaf size and SAH traversal-cost sweep.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_ax.txt; : > $S
run() { local label=$1; shift
  env "$@" timeout 300 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$label', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms/step box/ray', round(r['box_tests_per_ray'],1), 'sph/ray', round(r['sphere_tests_per_ray'],2))" | tee -a $S; }
run default A=1
run leaf2 RT_B200_MAX_LEAF=2
run leaf8 RT_B200_MAX_LEAF=8
run ctrav0.6 RT_B200_BVH_CTRAV=0.6
run ctrav2.5 RT_B200_BVH_CTRAV=2.5
run leaf8_ctrav2.5 RT_B200_MAX_LEAF=8 RT_B200_BVH_CTRAV=2.5
