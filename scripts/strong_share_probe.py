#!/usr/bin/env python
"""What one GPU of N does in a strong split of a config's frame, measured on ONE GPU: the sample share (whole frame, spp/N
samples) and the tile share (rows r, r+N, ... at all samples), against the whole frame.  efficiency = t(1) / (N * t(N)):
what the split itself costs before any exchange between devices."""
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from raytracer_go_b200 import api, scenes  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
scene, o = scenes.build_config(cfg)
cam = api.camera_from_options(o)


def best(f, n=4):
    r = None
    for _ in range(n):
        t = time.perf_counter()
        st = f()
        dt = (time.perf_counter() - t) * 1e3
        if r is None or dt < r[0]:
            r = (dt, st.ms_render, st.ms_megakernel, st.kernel_launches)
    return r


with api.Scene(scene) as sc:
    sc.render(cam, sample_count=8)
    full = best(lambda: sc.render(cam)[2])
    print(f"{cfg} {cam.width}x{cam.height}x{cam.spp}: wall {full[0]:.2f} ms, device {full[1]:.2f}, kernels {full[2]:.2f}, launches {full[3]}")
    for n in (2, 4, 8):
        s = best(lambda: sc.render(cam, sample_offset=0, sample_count=cam.spp // n)[2])
        t = best(lambda: sc.render(cam, rows=(0, (cam.height + n - 1) // n, n))[2])
        b = best(lambda: sc.render(cam, rows=(0, (cam.height + n - 1) // n, 1))[2])
        for name, r in (("sample share", s), ("tile share (interleaved rows)", t), ("band share (top rows)", b)):
            print(f"  N={n} {name:30s}: wall {r[0]:7.2f} ms, device {r[1]:7.2f}, kernels {r[2]:7.2f}, launches {r[3]:3d}, "
                  f"efficiency (device) {full[1] / (n * r[1]):.3f}")
