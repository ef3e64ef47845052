#!/usr/bin/env python
"""Generates tests/golden/c2_4096spp_reference_algorithm_u16.npz: config C2's frame (1200x675) at 4096 spp, depth 50,
rendered by the CPU oracle running the REFERENCE's algorithm — random-axis median-split BVH (bvh.go:142-249), recursive
radiance (ray.go:32-54) — on its own sample set (seed 0xB0B, drawn from the 10-ROUND Philox stream: the committed frame predates the
switch of the render streams to 7 rounds, and comparing the device's 7-round samples with it also checks one round
count against the other).  The `-m gpu` test
test_c2_named_size_converged_psnr compares the device's 4096-spp render (seed 0xA11CE) with it (north_star: PSNR >= 40 dB).

The oracle needs ~5 minutes on 16 cores for the 3.3 G samples, too long for the test suite, hence the committed frame.
Stored as the clipped per-pixel mean quantised to 16 bits (4.9 MB instead of 9.7 MB of float32; the quantisation
error 7.6e-6 is 100 dB below the peak).  Run from the repo root:  python tests/golden/make_c2_converged.py [threads]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402
from raytracer_go_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SPP, SEED, BVH_SEED = 4096, 0xB0B, 3


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    s = scenes.random_scene()
    orc.set_philox_rounds(10)
    cam = orc.camera_from_options(scenes.camera_options(1200, SPP))
    t0 = time.time()
    rgb, acc, st = orc.render(s, cam, SEED, mode=orc.MODE_REF_BVH, order=orc.ORDER_RECURSIVE, bvh_seed=BVH_SEED, threads=threads)
    dt = time.time() - t0
    mean = np.clip(acc / SPP, 0.0, 1.0)
    q = np.round(mean * 65535.0).astype(np.uint16)
    np.savez_compressed(os.path.join(HERE, "c2_4096spp_reference_algorithm_u16.npz"), scene_sha=s.sha256(), seed=SEED,
                        bvh_seed=BVH_SEED, spp=SPP, mean_u16=q, rgb=rgb, rays=st.rays, samples=st.samples,
                        image_mean=float(mean.mean()))
    print(f"{cam.width}x{cam.height}x{SPP}: {dt:.1f} s on {st.threads} threads, {st.samples / dt / 1e6:.2f} Msamples/s, "
          f"segments/sample {st.rays / st.samples:.4f}, image mean {mean.mean():.6f}")


if __name__ == "__main__":
    main()
