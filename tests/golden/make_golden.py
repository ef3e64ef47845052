#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/oracle.cpp).

The reference ships no golden vectors and cannot be run here (Go is not installed), so these
fixtures pin the ORACLE's outputs on small seeded inputs: a regression net for the oracle itself
(`-m "not gpu"`) and a second, file-based parity target for the CUDA path (`-m gpu`).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402
from raytracer_go_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    s = scenes.random_scene()
    # 1. closest hits of 4096 primary + 4096 scattered rays (World.Hit semantics)
    cam = orc.camera_from_options(scenes.camera_options(128, 1))
    ro, rd = orc.primary_rays(cam, 2024, 0, 4096, 0, 1)
    rng = np.random.default_rng(5)
    so = (rng.uniform([-12, 0.01, -12], [12, 2.5, 12], size=(4096, 3))).astype(np.float32)
    sd = rng.normal(size=(4096, 3)).astype(np.float32)
    o, d = np.concatenate([ro, so]), np.concatenate([rd, sd])
    ids, ts = orc.trace(s, o, d)
    np.savez_compressed(os.path.join(HERE, "trace_random_scene.npz"), scene_sha=s.sha256(), origins=o, dirs=d,
                        ids=ids, ts=ts)
    # 2. a 64x36 render, 4 spp, samples [3, 7): accumulators in both radiance orders + RGB8
    cam = orc.camera_from_options(scenes.camera_options(64, 4))
    rgb_i, acc_i, st = orc.render(s, cam, 77, sample_offset=3, sample_count=4, order=orc.ORDER_ITERATIVE)
    rgb_r, acc_r, _ = orc.render(s, cam, 77, sample_offset=3, sample_count=4, order=orc.ORDER_RECURSIVE)
    np.savez_compressed(os.path.join(HERE, "render_random_64x36.npz"), scene_sha=s.sha256(), seed=77, sample_offset=3,
                        sample_count=4, acc_iterative=acc_i, acc_recursive=acc_r, rgb_iterative=rgb_i,
                        rgb_recursive=rgb_r, rays=st.rays, hits=st.hits)
    # 3. the image-textured sphere (main.go:80-104) seen from -z so the out-of-bounds stripe is visible
    e = scenes.earth_scene()
    cam = orc.camera_from_options(scenes.camera_options(64, 2, look_from=(0, 0, -12), defocus_deg=0.0))
    rgb, acc, _ = orc.render(e, cam, 5, order=orc.ORDER_ITERATIVE)
    np.savez_compressed(os.path.join(HERE, "render_earth_64x36.npz"), scene_sha=e.sha256(), seed=5, acc=acc, rgb=rgb)
    # 4. camera rays
    cam = orc.camera_from_options(scenes.camera_options(400, 2))
    po, pd = orc.primary_rays(cam, 9, 400 * 100, 512, 1, 2)
    np.savez_compressed(os.path.join(HERE, "primary_rays_400.npz"), seed=9, pixel_begin=400 * 100, n_pixels=512,
                        sample_offset=1, sample_count=2, origins=po, dirs=pd, camera=np.frombuffer(bytes(cam), np.uint8))
    # 5. the other four scenes of main.go (quads, boxes, diffuse light, Perlin noise): small renders and,
    #    for the quad scenes, closest hits of the primary rays (Quad.Hit, hittables.go:167-194)
    out = {}
    for name, (sc, opts) in other_scenes().items():
        cam = orc.camera_from_options(opts)
        rgb, acc, st = orc.render(sc, cam, 31, order=orc.ORDER_ITERATIVE)
        out[name + "_sha"], out[name + "_acc"], out[name + "_rgb"] = sc.sha256(), acc, rgb
        out[name + "_rays"] = st.rays
        if len(sc.quads):
            ro, rd = orc.primary_rays(cam, 31, 0, cam.width * cam.height, 0, 1)
            ids, ts = orc.trace(sc, ro, rd)
            out[name + "_ids"], out[name + "_ts"] = ids, ts
    np.savez_compressed(os.path.join(HERE, "render_other_scenes.npz"), **out)
    print("golden fixtures written to", HERE)


def other_scenes():
    return {
        "cornell": (scenes.cornell_box_scene(), scenes.cornell_camera_options(48, 8)),
        "quads": (scenes.quad_demo_scene(), scenes.quad_demo_camera_options(64, 4)),
        "perlin": (scenes.perlin_demo_scene(), scenes.perlin_camera_options(64, 4)),
        "simple_light": (scenes.simple_light_scene(), scenes.simple_light_camera_options(64, 8)),
    }


if __name__ == "__main__":
    main()
