#!/bin/bash
# 2-GPU pass: in-library multi-device render + full gpu suite on a 2-GPU box + timing of rt_render_multi.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_k.txt
tail -6 gpurun_out/pytest_gpu_k.log
python - <<'PY' > gpurun_out/render_multi_k.txt 2>&1
import time
from raytracer_go_b200 import api, scenes
s = scenes.random_scene(); cam = api.camera_from_options(scenes.camera_options(1200, 500))
for devs in ([0], [0, 1], [0, 1]):
    t = time.perf_counter(); rgb, _, st = api.render_multi(s, cam, devs); dt = time.perf_counter() - t
    print(devs, f"wall {dt*1e3:.1f} ms  device max {st.ms_render:.1f} ms  {st.samples/dt/1e6:.0f} Msamples/s e2e  rays {st.rays}")
PY
cat gpurun_out/render_multi_k.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_k.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_c2_k.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], d['roofline']['traffic'])"
cat gpurun_out/summary_k.txt
