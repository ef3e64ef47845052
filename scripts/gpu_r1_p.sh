#!/bin/bash
# BASELINE config C5 at full size: 3840x2160, 4096 spp, depth 50, sample-split (strong) over all GPUs of the box.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $NG --config C5 --split strong --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5_full_n$NG.json 2> gpurun_out/bench_c5_full_n$NG.err; echo "C5 strong n=$NG rc=$?" > gpurun_out/summary_p.txt
tail -3 gpurun_out/bench_c5_full_n$NG.err
python - <<'PY'
import json, glob
for f in glob.glob('gpurun_out/bench_c5_full_n*.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['config']['workload'], round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],1), 'ms/step e2e', round(d['e2e']['value'],1), d['scaling'])
PY
python - <<'PY'
import time, subprocess
import numpy as np
from raytracer_go_b200 import api, scenes
n = len(subprocess.check_output(['nvidia-smi','-L']).decode().strip().splitlines())
scene, o = scenes.build_config("C5")
cam = api.camera_from_options(o)
api.render_multi(scene, api.camera_from_options(scenes.camera_options(3840, 8)), list(range(n)))  # warm-up: contexts, workspaces
t = time.perf_counter(); rgb, _, st = api.render_multi(scene, cam, list(range(n))); dt = time.perf_counter() - t
print(f"rt_render_multi C5 full ({cam.width}x{cam.height}x{cam.spp}) on {n} GPUs: wall {dt:.3f} s, device max {st.ms_render:.0f} ms, {st.samples/dt/1e6:.0f} Msamples/s, {st.rays/dt/1e6:.0f} Mrays/s")
from PIL import Image
Image.fromarray(rgb).resize((960,540), Image.LANCZOS).save('gpurun_out/c5_4k_4096spp_960px.png', optimize=True)
PY
cat gpurun_out/summary_p.txt
