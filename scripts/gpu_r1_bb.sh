#!/bin/bash
# 2 GPUs under the 512 x 2 default: gpu tests (incl. the two-device ones and the launch-shape test), torchrun
# bench in sample-split and tile-split mode, reference arm under torchrun, rt_render_multi in both modes.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_bb.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_bb.txt
tail -3 gpurun_out/pytest_bb.txt
for split in weak tile; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 3 --warmup 3 --split $split --no-cpu-baseline > gpurun_out/bench_bb_$split.json 2> gpurun_out/bench_bb_$split.err
  echo "$split rc=$?"; cat gpurun_out/bench_bb_$split.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_bb_reference.json 2> gpurun_out/bench_bb_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_bb_reference.json
timeout 600 python - > gpurun_out/multi_bb.txt 2>&1 <<'PY'
import time, numpy as np
from raytracer_go_b200 import api, scenes
data, opts = scenes.build_config("C2")
cam = api.camera_from_options(opts)
with api.Scene(data) as sc:
    rgb1, acc1, st1 = sc.render(cam, want_accum=True)
for tile in (False, True):
    for _ in range(2):
        t0 = time.perf_counter()
        rgb, acc, st = api.render_multi(data, cam, [0, 1], want_accum=True, tile_split=tile)
        dt = time.perf_counter() - t0
    print("tile" if tile else "sample", "wall %.1f ms, device %.1f ms," % (dt * 1e3, st.ms_render),
          "bitwise equal to 1 GPU:", bool(np.array_equal(rgb, rgb1) and np.array_equal(acc.view(np.uint32), acc1.view(np.uint32))),
          "max |drgb|", int(np.abs(rgb.astype(int) - rgb1.astype(int)).max()))
PY
cat gpurun_out/multi_bb.txt
