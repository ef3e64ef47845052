#!/usr/bin/env python
"""rt_render_multi (one call, in-library threads, all GPUs of the box) on the C2 and C5 frames, both split modes."""
import subprocess
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from raytracer_go_b200 import api, scenes  # noqa: E402

n = len(subprocess.check_output(["nvidia-smi", "-L"]).decode().strip().splitlines())
cfgs = sys.argv[1:] or ["C2", "C5"]
for cfg in cfgs:
    scene, o = scenes.build_config(cfg)
    cam = api.camera_from_options(o)
    for nd in sorted({1, 2, 4, 8, n}):
        if nd > n:
            continue
        devs = list(range(nd))
        api.render_multi(scene, api.camera_from_options(scenes.camera_options(cam.width, 8)), devs)  # warm-up
        for tile in (False, True):
            best = None
            for _ in range(3):
                t = time.perf_counter()
                rgb, _, st = api.render_multi(scene, cam, devs, tile_split=tile)
                dt = time.perf_counter() - t
                if best is None or dt < best[0]:
                    best = (dt, st)
            dt, st = best
            print(f"rt_render_multi {cfg} ({cam.width}x{cam.height}x{cam.spp}, {'tile' if tile else 'sample'}-split) on {nd} GPUs: "
                  f"wall {dt*1e3:.1f} ms, device max {st.ms_render:.1f} ms, {st.samples/dt/1e6:.0f} Msamples/s, {st.rays/dt/1e6:.0f} Mrays/s")
