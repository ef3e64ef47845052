#!/usr/bin/env python
"""Design exploration: lock-step SIMT replay of the secondary megakernel on real C2 path segments under
different traversal algorithms (tests/hostsim/travsim.cpp).  Output: profiles/experiments/r02_traversal_variants.txt"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracer_go_b200 import api, scenes  # noqa: E402

subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests/hostsim"), "libtravsim.so"])
ts = C.CDLL(os.path.join(ROOT, "tests/hostsim/libtravsim.so"))
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
scene, opts = scenes.build_config(cfg)
desc, keep = scene.to_desc()
cam = api.camera_from_options(opts)
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ts.ts_run(C.byref(desc), C.byref(cam), C.c_uint64(scenes.RENDER_SEED), 32, stride, int(os.environ.get("RT_B200_MAX_LEAF", 4)))
