#!/bin/bash
# TEST INFRASTRUCTURE (lives under tests/ because it runs the oracle as the checker): run from the repo root on a GPU box.
# north_star check at the named size: converged 4096-spp image of the random scene at 1200x675, device vs the
# oracle running the reference's own algorithm (random-axis BVH, recursive radiance) on an independent sample set.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/psnr_c2_4096spp.txt 2>&1
import time, numpy as np
from raytracer_go_b200 import api, scenes
from oracle import pyoracle as orc
from PIL import Image
scene = scenes.random_scene()
cam = api.camera_from_options(scenes.camera_options(1200, 4096))
t = time.perf_counter()
with api.Scene(scene) as sc:
    rgb, acc, st = sc.render(cam, 0xA11CE, want_accum=True)
t_dev = time.perf_counter() - t
t = time.perf_counter()
rrgb, racc, rst = orc.render(scene, cam, 0xB0B, mode=orc.MODE_REF_BVH, order=orc.ORDER_RECURSIVE, bvh_seed=3)
t_cpu = time.perf_counter() - t
def psnr(a, b, peak):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2); return 10 * np.log10(peak ** 2 / mse)
m, rm = np.clip(acc / 4096, 0, 1), np.clip(racc / 4096, 0, 1)
print(f"C2 frame 1200x675, 4096 spp, depth 50: device {t_dev:.2f} s ({st.samples/t_dev/1e6:.0f} Msamples/s e2e), "
      f"oracle (reference algorithm, {rst.threads} threads) {t_cpu:.1f} s ({rst.samples/t_cpu/1e6:.1f} Msamples/s)")
print(f"PSNR linear radiance {psnr(m, rm, 1.0):.2f} dB, PSNR RGB8 {psnr(rgb, rrgb, 255.0):.2f} dB")
d = np.abs(m - rm)
print(f"per-pixel |mean difference|: max {d.max():.4f}, p99.9 {np.percentile(d, 99.9):.4f}, p99 {np.percentile(d, 99):.4f}, mean {d.mean():.5f}")
print(f"image means: device {m.mean():.6f} oracle {rm.mean():.6f} (bias {m.mean()-rm.mean():+.2e}); RGB8 max level difference {np.abs(rgb.astype(int)-rrgb.astype(int)).max()}")
print(f"segments/sample device {st.rays/st.samples:.4f} oracle {rst.rays/rst.samples:.4f}")
both = np.concatenate([rgb, rrgb], axis=1)
Image.fromarray(both).resize((1200, 338), Image.LANCZOS).save('gpurun_out/psnr_c2_device_left_oracle_right.jpg', quality=85)
PY
cat gpurun_out/psnr_c2_4096spp.txt
