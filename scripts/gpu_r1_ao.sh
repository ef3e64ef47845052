#!/bin/bash
# 8 GPUs, final round-1 build: N = 1/2/4/8 scaling launched as the driver does (torchrun, NCCL), BASELINE
# config C5 at full size (3840x2160, 4096 spp) strong sample-split, rt_render_multi on the C2 and C5 frames.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
S=gpurun_out/summary_ao.txt; echo "gpus=$NG" > $S
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_ao_n$n.json 2> gpurun_out/scale_ao_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_ao_n$n.json 2> gpurun_out/scale_ao_n$n.err
  fi
  echo "n=$n rc=$?" >> $S
done
python - <<'PY' | tee -a gpurun_out/summary_ao.txt
import json
base=None
for n in (1,2,4,8):
    try: d=json.loads(open(f'gpurun_out/scale_ao_n{n}.json').read().strip().splitlines()[-1])
    except Exception: continue
    if n==1: base=d['value']
    print(f"N={n} value {d['value']:.1f} Msamples/s  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.1f}  eff {d['value']/(n*base):.3f}")
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $NG --config C5 --split strong --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ao_c5_full_n$NG.json 2> gpurun_out/bench_ao_c5.err; echo "C5 strong n=$NG rc=$?" | tee -a $S
python - <<'PY' | tee -a gpurun_out/summary_ao.txt
import json, glob
for f in glob.glob('gpurun_out/bench_ao_c5_full_n*.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['config']['workload'], round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],1), 'ms/step e2e', round(d['e2e']['value'],1), d['scaling'])
PY
python - <<'PY' 2>&1 | tee -a gpurun_out/summary_ao.txt
import time, subprocess
from raytracer_go_b200 import api, scenes
n = len(subprocess.check_output(['nvidia-smi','-L']).decode().strip().splitlines())
for cfg in ("C2", "C5"):
    scene, o = scenes.build_config(cfg)
    cam = api.camera_from_options(o)
    api.render_multi(scene, api.camera_from_options(scenes.camera_options(cam.width, 8)), list(range(n)))  # warm-up
    for tile in (False, True):
        t = time.perf_counter(); rgb, _, st = api.render_multi(scene, cam, list(range(n)), tile_split=tile); dt = time.perf_counter() - t
        print(f"rt_render_multi {cfg} ({cam.width}x{cam.height}x{cam.spp}, {'tile' if tile else 'sample'}-split) on {n} GPUs: wall {dt*1e3:.1f} ms, device max {st.ms_render:.1f} ms, {st.samples/dt/1e6:.0f} Msamples/s, {st.rays/dt/1e6:.0f} Mrays/s")
PY
