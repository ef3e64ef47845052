#!/bin/bash
# Parameter sweep in the default two-stage mode.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/variants_u.txt
run() { label="$1"; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', round(d['value'],1),'Msamples/s ms/step', round(d['ms_per_step'],2), 'box/ray', round(r['box_tests_per_ray'],2), 'sph/ray', round(r['sphere_tests_per_ray'],2))" >> gpurun_out/variants_u.txt 2>&1
}
run "default" A=1
run "default again" A=1
for c in 64 128 512 1024 4096; do run "chunk $c" RT_B200_CHUNK=$c; done
for l in 1 2 8; do run "leaf $l" RT_B200_MAX_LEAF=$l; done
for ct in 0.6 2.0 3.0; do run "ctrav $ct" RT_B200_BVH_CTRAV=$ct; done
for b in 8 32 64; do run "bins $b" RT_B200_BVH_BINS=$b; done
run "pass 64M" RT_B200_PASS_PATHS=67108864
run "pass 16M" RT_B200_PASS_PATHS=16777216
run "minb4" RT_B200_MINB=4
run "b512m2" RT_B200_BLOCK=512 RT_B200_MINB=2
run "regen2" RT_B200_REGEN_MIN=2
run "regen4" RT_B200_REGEN_MIN=4
run "nosmem" RT_B200_NO_SMEM=1
cat gpurun_out/variants_u.txt
