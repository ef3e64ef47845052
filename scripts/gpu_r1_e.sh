#!/bin/bash
# Fifth GPU pass: speculative while-while traversal; C3 / C4 configs.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1e.so
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_e.txt
tail -5 gpurun_out/pytest_gpu_e.log
run() { label="$1"; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s frac', round(r['frac'],4), 'box/ray', round(r['box_tests_per_ray'],2))" >> gpurun_out/variants_e.txt 2>&1
}
for sm in 0 4 8 12 16 20 24 32; do run "spec_min $sm" RT_B200_SPEC_MIN=$sm; done
run "spec 12 regen 1" RT_B200_SPEC_MIN=12 RT_B200_REGEN_MIN=1
run "spec 12 regen 16" RT_B200_SPEC_MIN=12 RT_B200_REGEN_MIN=16
run "spec 12 b256m4" RT_B200_SPEC_MIN=12 RT_B200_MINB=4
run "spec 12 b512m2" RT_B200_SPEC_MIN=12 RT_B200_BLOCK=512 RT_B200_MINB=2
run "spec 12 leaf1" RT_B200_SPEC_MIN=12 RT_B200_MAX_LEAF=1
run "spec 12 leaf2" RT_B200_SPEC_MIN=12 RT_B200_MAX_LEAF=2
cat gpurun_out/variants_e.txt
timeout 600 python bench.py --config C3 --spp 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_e.json 2> gpurun_out/bench_c3_e.err; echo "C3 rc=$?" >> gpurun_out/summary_e.txt
timeout 600 python bench.py --config C4 --spp 8 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_e.json 2> gpurun_out/bench_c4_e.err; echo "C4 rc=$?" >> gpurun_out/summary_e.txt
tail -3 gpurun_out/bench_c4_e.err
python -c "
import json
for f in ['gpurun_out/bench_c3_e.json','gpurun_out/bench_c4_e.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, round(d['value'],1),'Msamples/s', round(d['mrays_s'],1), 'Mrays/s box/ray', round(r['box_tests_per_ray'],1), 'sph/ray', round(r['sphere_tests_per_ray'],2), 'seg', round(r['segments_per_sample'],2), 'ms', round(d['ms_per_step'],1))
    except Exception as e: print(f, 'ERR', e)
"
CMD="python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r1e $CMD > gpurun_out/ncu_full_e.log 2>&1
cat gpurun_out/summary_e.txt
