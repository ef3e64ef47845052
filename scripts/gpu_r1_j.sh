#!/bin/bash
# Ninth GPU pass: noise-texture scenes; compute-sanitizer memcheck on small cases.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary_j.txt
tail -12 gpurun_out/pytest_gpu_j.log
cat > /tmp/mc.py <<'PY'
import numpy as np
from raytracer_go_b200 import api, scenes
for s, o in [(scenes.random_scene(), scenes.camera_options(96, 2)), (scenes.cornell_box_scene(), scenes.cornell_camera_options(48, 2)),
             (scenes.simple_light_scene(), scenes.simple_light_camera_options(64, 2)), (scenes.earth_scene(), scenes.camera_options(64, 2, look_from=(0, 0, -12), defocus_deg=0.0)),
             (scenes.random_scene(half=30), scenes.camera_options(64, 1))]:
    cam = api.camera_from_options(o)
    with api.Scene(s) as sc:
        rgb, acc, st = sc.render(cam, 1, want_accum=True)
        ids, ts = sc.trace(np.zeros((100, 3), np.float32) + (1, 2, 3), np.random.default_rng(0).normal(size=(100, 3)).astype(np.float32))
    print(s.name, st.rays, rgb.mean())
PY
RT_B200_PASS_PATHS=5000 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/mc.py > gpurun_out/memcheck_j.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/summary_j.txt
tail -8 gpurun_out/memcheck_j.log
cat gpurun_out/summary_j.txt
