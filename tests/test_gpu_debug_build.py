"""The bounds-checked debug build (csrc/rt_debug.h, librt_b200_debug.so) on the GPU: it must render what the release
build renders, bit for bit, without tripping a single check on scenes that exercise every array the kernels index
(shared-memory and global-memory scenes, quads, image and Perlin textures, ragged sizes, multi-pass renders), and it
must report a violation when one is provoked.  compute-sanitizer is closed on the development pool; this build stands
in for it.  The library is chosen when raytracer_go_b200.lib is imported, so each build runs in its own process.

The complete `-m gpu` suite also runs under the debug build (scripts/gpu_run.sh validate-debug; log under profiles/).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
from raytracer_go_b200 import api, scenes, lib
sys.path.insert(0, %(root)r + "/tests")
import fuzz_scenes
out = {}
def run(name, scene, opts, seed=7, **kw):
    cam = api.camera_from_options(opts)
    with api.Scene(scene) as sc:
        rgb, acc, st = sc.render(cam, seed, want_accum=True, **kw)
        n = cam.width * cam.height
        ro, rd = api.primary_rays(cam, seed, 0, n, 0, 1)
        ids, ts = sc.trace(ro, rd)
    out[name + "_acc"], out[name + "_rgb"], out[name + "_ids"], out[name + "_t"] = acc, rgb, ids, ts
import os
os.environ["RT_B200_PASS_PATHS"] = "50000"            # several passes that do not divide the path count
run("random", scenes.random_scene(), scenes.camera_options(157, 5))
run("random_depth1", scenes.random_scene(), scenes.camera_options(33, 3, max_depth=1))
del os.environ["RT_B200_PASS_PATHS"]
run("cornell", scenes.cornell_box_scene(), scenes.cornell_camera_options(61, 4))
run("mixed", scenes.mixed_scene(), scenes.camera_options(64, 3))
run("earth", scenes.earth_scene(), scenes.camera_options(64, 2, look_from=(0, 0, -12), defocus_deg=0.0))
run("perlin", scenes.perlin_demo_scene(), scenes.perlin_camera_options(64, 2))
run("light", scenes.simple_light_scene(), scenes.simple_light_camera_options(64, 2))
run("one_pixel", scenes.random_scene(), scenes.camera_options(1, 1))
for k in range(6):                                      # 1..400 spheres, coincident / nested / negative radii
    fs, fo, fd, _ = fuzz_scenes.fuzz_scene_and_rays(k, 3000)
    with api.Scene(fs) as sc:
        out["fuzz%%d_ids" %% k], out["fuzz%%d_t" %% k] = sc.trace(fo, fd)
s = scenes.stress_scene(60)                             # 14 k spheres: global-memory scene, local-memory stack
os.environ["RT_B200_NO_SMEM"] = "1"
run("global", s, scenes.camera_options(96, 2, look_from=(52, 24, 12)))
np.savez(sys.argv[1], **out)
print("child ok", lib.LIB_PATH)
'''

TRIP = r'''
import sys
sys.path.insert(0, %(root)r)
from raytracer_go_b200 import api, scenes, lib, abi
cam = api.camera_from_options(scenes.camera_options(32, 1))
try:
    with api.Scene(scenes.random_scene()) as sc:
        sc.render(cam, 1)
except lib.RtError as e:
    assert e.code == abi.RT_ERR_INTERNAL and "traversal stack depth" in str(e), str(e)
    print("tripped:", e)
    sys.exit(0)
sys.exit("the provoked stack overflow went unnoticed")
'''


def _child(code, env_extra, *args):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, "-c", code % dict(root=ROOT), *args], env=env, capture_output=True, text=True,
                          timeout=900)


def test_debug_build_is_clean_and_bit_identical(gpu, tmp_path):
    rel, dbg = str(tmp_path / "release.npz"), str(tmp_path / "debug.npz")
    r = _child(CHILD, {"RT_B200_DEBUG": "0"}, rel)
    assert r.returncode == 0 and "librt_b200.so" in r.stdout, r.stdout + r.stderr
    d = _child(CHILD, {"RT_B200_DEBUG": "1"}, dbg)
    assert d.returncode == 0 and "librt_b200_debug.so" in d.stdout, d.stdout + d.stderr   # no check tripped
    a, b = np.load(rel), np.load(dbg)
    assert sorted(a.files) == sorted(b.files) and len(a.files) >= 44
    for k in a.files:
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k


def test_debug_build_reports_a_provoked_violation(gpu):
    r = _child(TRIP, {"RT_B200_DEBUG": "1", "RT_B200_DEBUG_TRIP": "1"})
    assert r.returncode == 0 and "tripped" in r.stdout, r.stdout + r.stderr
