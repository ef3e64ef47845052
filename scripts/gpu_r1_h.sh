#!/bin/bash
# Seventh GPU pass: quads / Cornell box; full test-suite; final-ish bench + profiles.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp raytracer_go_b200/csrc/librt_b200.so gpurun_out/librt_b200_r1h.so
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_h.txt
tail -30 gpurun_out/pytest_gpu_h.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_h.json 2>gpurun_out/bench_c2_h.err; echo "bench rc=$?" >> gpurun_out/summary_h.txt
timeout 300 python bench.py --config CB --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_cb_h.json 2>gpurun_out/bench_cb_h.err; echo "bench CB rc=$?" >> gpurun_out/summary_h.txt
python - <<'PY'
import json
for f in ['gpurun_out/bench_c2_h.json','gpurun_out/bench_cb_h.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['value'],1),'Msamples/s', round(d['mrays_s'],1),'Mrays/s e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],4), 'seg', round(r['segments_per_sample'],2), 'box/ray', round(r['box_tests_per_ray'],2), 'prim/ray', round(r['sphere_tests_per_ray'],2))
    except Exception as e: print(f, 'ERR', e)
PY
mkdir -p out && ./raytracer_go_b200/host/rt_demo cornell 600 200 out/cornell.ppm && python - <<'PY'
from PIL import Image
import numpy as np
t=open('out/cornell.ppm').read().split()
w,h=int(t[1]),int(t[2]); a=np.array(t[4:],dtype=np.uint8).reshape(h,w,3); Image.fromarray(a).save('gpurun_out/cornell_600_200spp.png'); print('cornell png', a.mean())
PY
cat gpurun_out/summary_h.txt
