#!/usr/bin/env python
"""Design exploration (not part of the product or the tests): what a per-pixel beam walk of the BVH would
cost for the primary stage.  For 32-sample warps of pixels spread over the C2 frame,
tests/hostsim's hs_beam_stats walks the flat BVH once with the interval slab test of {origin box} x
{direction box} and reports the node pairs visited, the candidate spheres, how many candidates survive a
front-to-back early exit, and the per-ray traversal's work for comparison.
Run from the repo root after `make -C tests/hostsim`."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracer_go_b200 import api, scenes  # noqa: E402

hs = C.CDLL(os.path.join(ROOT, "tests/hostsim/libhostsim.so"))
hs.hs_beam_stats.restype = C.c_int64
for name, (scene, opts) in {"C2": scenes.build_config("C2"),
                             "C4 at 1/25 size (40 K spheres)": scenes.build_config("C4", stress_half=100)}.items():
    desc, keep = scene.to_desc()
    cam = api.camera_from_options(opts)
    n_pix = cam.width * cam.height
    n = 20000 if name == "C2" else 4000
    stride = max(1, n_pix // n)
    out = np.zeros((n, 5), np.float32)
    got = hs.hs_beam_stats(C.byref(desc), C.byref(cam), C.c_uint64(7), C.c_int64(0), C.c_int64(n), C.c_int64(stride), 4,
                           out.ctypes.data_as(C.c_void_p))
    o = out[:got]
    print(f"{name}: {got} warps (32 samples of one pixel each)")
    print("  beam walk      : node pairs %.1f (p90 %.0f, max %.0f), candidate spheres %.1f (p90 %.0f, max %.0f)"
          % (o[:, 0].mean(), np.percentile(o[:, 0], 90), o[:, 0].max(), o[:, 1].mean(), np.percentile(o[:, 1], 90), o[:, 1].max()))
    print("  tested with the front-to-back early exit: %.1f per ray (p90 %.0f)" % (o[:, 2].mean(), np.percentile(o[:, 2], 90)))
    print("  per-ray traversal today: node pairs %.1f, sphere tests %.2f per ray" % (o[:, 3].mean(), o[:, 4].mean()))
    # issue slots per warp: today 32 lanes walk in lock-step (48 per pair, ~30 per sphere test, ~0.83 lane occupancy)
    today = (o[:, 3] * 48 + o[:, 4] * 30) / 0.83
    beam = o[:, 0] / 32 * 60 * 2 + 50 + o[:, 2] * 30 + 40   # pairs tested 32 at a time (2 boxes each), sort, exact tests, vote
    print("  issue slots per warp (model): today %.0f, beam %.0f (x%.2f)" % (today.mean(), beam.mean(), today.mean() / beam.mean()))
