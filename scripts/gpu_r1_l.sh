#!/bin/bash
# N-GPU scaling pass, launched exactly as the driver does (torchrun, one rank per GPU, NCCL).
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "gpus=$NG" > gpurun_out/summary_l.txt
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "n=$n rc=$?" >> gpurun_out/summary_l.txt
done
python - <<'PY'
import json, glob
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/scale_n{n}.json').read().strip().splitlines()[-1])
    except Exception as e:
        continue
    if n==1: base=d['value']
    print(f"N={n} value {d['value']:.1f} Msamples/s  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.1f}  eff {d['value']/(n*base):.3f}")
PY
python - <<'PY'
import time
from raytracer_go_b200 import api, scenes
import subprocess
n = len(subprocess.check_output(['nvidia-smi','-L']).decode().strip().splitlines())
s = scenes.random_scene(); cam = api.camera_from_options(scenes.camera_options(1200, 500))
for devs in ([0], list(range(n)), list(range(n))):
    t = time.perf_counter(); rgb, _, st = api.render_multi(s, cam, devs); dt = time.perf_counter() - t
    print('rt_render_multi', devs, f"wall {dt*1e3:.1f} ms  device max {st.ms_render:.1f} ms  {st.samples/dt/1e6:.0f} Msamples/s")
PY
cat gpurun_out/summary_l.txt
