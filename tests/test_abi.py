"""C-ABI checks that need no GPU: the library loads, exports every symbol include/rt_b200.h
declares, agrees with the ctypes struct layouts, fails loudly without a device, and its host-only
helper (Camera.init) equals the oracle's."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

from raytracer_go_b200 import abi, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rt_b200.h")


def test_exports_every_declared_symbol(rtlib):
    text = open(HEADER).read()
    declared = set(re.findall(r"^(?:int|void|const char \*)\s*(rt_[a-z0-9_]+)\(", text, re.M))
    assert declared == set(abi.PROTOTYPES), declared ^ set(abi.PROTOTYPES)
    for name in declared:
        assert hasattr(rtlib, name), f"librt_b200.so does not export {name}"
    assert rtlib.rt_abi_version() == abi.RT_B200_ABI_VERSION


def test_struct_layouts_match_the_header():
    """Compile a probe against the real header and compare sizeof/offsetof with ctypes."""
    structs = {n: getattr(abi, n) for n in ("rt_sphere", "rt_quad", "rt_material", "rt_texture", "rt_image", "rt_perlin", "rt_scene_desc",
                                           "rt_camera", "rt_camera_options", "rt_render_opts", "rt_stats",
                                           "rt_bvh_info")}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for n, s in structs.items():
        lines.append(f'printf("{n} %zu\\n", sizeof({n}));')
        for f, _ in s._fields_:
            lines.append(f'printf("{n}.{f} %zu\\n", offsetof({n}, {f}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "p.c"), os.path.join(d, "p")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-o", exe, src])
        out = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for n, s in structs.items():
        assert int(out[n]) == C.sizeof(s), n
        for f, _ in s._fields_:
            assert int(out[f"{n}.{f}"]) == getattr(s, f).offset, f"{n}.{f}"


def test_no_device_is_a_loud_error(rtlib):
    """Without a B200 every compute entry point returns RT_ERR_NO_DEVICE — there is no CPU path."""
    if rtlib.rt_device_count() > 0:
        return
    desc, keep = scenes.random_scene().to_desc()
    h = C.c_void_p()
    assert rtlib.rt_scene_create(C.byref(desc), 0, C.byref(h)) == abi.RT_ERR_NO_DEVICE
    assert not h.value and b"no CPU path" in rtlib.rt_last_error()
    cam = abi.rt_camera()
    assert rtlib.rt_camera_from_options(C.byref(scenes.camera_options(64, 1)), C.byref(cam)) == 0
    opts = abi.rt_render_opts(1, 0, 0, 1, 0)
    buf = np.zeros((cam.width * cam.height, 3), np.float32)
    rc = rtlib.rt_primary_rays(C.byref(cam), C.byref(opts), 0, 4, buf.ctypes.data_as(C.c_void_p),
                               buf.ctypes.data_as(C.c_void_p))
    assert rc == abi.RT_ERR_NO_DEVICE


def test_invalid_arguments(rtlib):
    h = C.c_void_p()
    assert rtlib.rt_scene_create(None, 0, C.byref(h)) == abi.RT_ERR_INVALID_ARGUMENT
    bad = scenes.random_scene()
    desc, keep = bad.to_desc()
    desc.abi_version = 99
    assert rtlib.rt_scene_create(C.byref(desc), 0, C.byref(h)) == abi.RT_ERR_INVALID_ARGUMENT
    assert b"abi_version" in rtlib.rt_last_error()
    bad.materials["kind"][5] = 17
    desc, keep = bad.to_desc()
    assert rtlib.rt_scene_create(C.byref(desc), 0, C.byref(h)) == abi.RT_ERR_UNSUPPORTED
    assert rtlib.rt_camera_from_options(None, None) == abi.RT_ERR_INVALID_ARGUMENT
    rtlib.rt_scene_destroy(None)  # must be a no-op


def test_camera_from_options_equals_oracle(rtlib, orc):
    """rt_camera_from_options (host, camera.go:128-166) is bit-identical to the oracle's."""
    cases = [scenes.camera_options(w, 7) for w in (400, 1200, 1920, 3840, 37)]
    cases.append(scenes.camera_options(600, 3, look_from=(278, 278, -800), look_at=(278, 278, 0), vfov_deg=40,
                                       defocus_deg=0, aspect=1.0, background=(0, 0, 0)))
    cases.append(scenes.camera_options(400, 3, look_from=(0, 0, 12), defocus_deg=0))
    for o in cases:
        a, b = abi.rt_camera(), orc.camera_from_options(o)
        assert rtlib.rt_camera_from_options(C.byref(o), C.byref(a)) == 0
        assert bytes(a) == bytes(b)


def test_non_finite_geometry_is_rejected(rtlib):
    """Validation happens before any device work, so it is checkable without a GPU."""
    bad = scenes.random_scene()
    bad.spheres["cx"][7] = np.inf
    desc, keep = bad.to_desc()
    h = C.c_void_p()
    assert rtlib.rt_scene_create(C.byref(desc), 0, C.byref(h)) == abi.RT_ERR_INVALID_ARGUMENT
    assert b"non-finite" in rtlib.rt_last_error()
    m = scenes.mixed_scene()
    m.quad_ids = np.array([0, 2, 3, 3], np.uint32)        # not a permutation
    desc, keep = m.to_desc()
    rc = rtlib.rt_scene_create(C.byref(desc), 0, C.byref(h))
    assert rc in (abi.RT_ERR_INVALID_ARGUMENT, abi.RT_ERR_NO_DEVICE)   # the id check needs the device selected first


def test_large_scene_validation_reports_the_first_bad_element(rtlib):
    """Arrays of >= 131 072 elements are validated on several threads; the error is still the FIRST bad element's."""
    n = 300_000
    sph = np.zeros(n, scenes.SPHERE_DT)
    sph["r"] = 0.2
    sph["cx"] = np.arange(n, dtype=np.float32)
    mats = np.zeros(1, scenes.MATERIAL_DT)
    tex = np.zeros(1, scenes.TEXTURE_DT)
    sph["material"][250_001] = 5          # out of range, in the last thread's share
    sph["cy"][140_000] = np.nan           # non-finite, earlier: this is the one to report
    desc, keep = scenes.SceneData(sph, mats, tex).to_desc()
    h = C.c_void_p()
    assert rtlib.rt_scene_create(C.byref(desc), 0, C.byref(h)) == abi.RT_ERR_INVALID_ARGUMENT
    assert rtlib.rt_last_error() == b"sphere 140000: non-finite geometry"
    sph["cy"][140_000] = 0.0
    desc, keep = scenes.SceneData(sph, mats, tex).to_desc()
    assert rtlib.rt_scene_create(C.byref(desc), 0, C.byref(h)) == abi.RT_ERR_INVALID_ARGUMENT
    assert rtlib.rt_last_error() == b"sphere 250001: material 5 out of range"
    many = np.zeros(200_000, scenes.MATERIAL_DT)
    many["kind"][199_999] = 77
    sph["material"][250_001] = 0
    desc, keep = scenes.SceneData(sph, many, tex).to_desc()
    assert rtlib.rt_scene_create(C.byref(desc), 0, C.byref(h)) == abi.RT_ERR_UNSUPPORTED
    assert rtlib.rt_last_error() == b"material 199999: unknown kind 77"


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the
    contract's keys on rank 0, nothing and exit 0 on the other ranks of a torchrun launch."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--width", "64", "--spp", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C2")
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29871")
    out = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_png_writers_round_trip(tmp_path):
    """The PNG output of the host mirrors (the reference's TODO at camera.go:196): api.encode_png and
    rtgo::write_png (stored deflate blocks, no compression library) decode to the RGB8 they were given —
    checked with Pillow, including an image larger than one 64 KiB deflate block."""
    from PIL import Image
    from raytracer_go_b200 import api
    rng = np.random.default_rng(8)
    for w, h in ((1, 1), (7, 5), (300, 200)):
        rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        path = tmp_path / f"py_{w}x{h}.png"
        path.write_bytes(api.encode_png(rgb))
        assert np.array_equal(np.asarray(Image.open(path).convert("RGB")), rgb)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "png_test.cpp"
    src.write_text('#include <fstream>\n#include "%s/raytracer_go_b200/host/rtgo.hpp"\n'
                   'int main(int argc, char **argv) {\n'
                   '    const int w = atoi(argv[1]), h = atoi(argv[2]);\n'
                   '    std::vector<uint8_t> rgb((size_t)w * h * 3);\n'
                   '    for (size_t i = 0; i < rgb.size(); i++) rgb[i] = (uint8_t)((i * 2654435761u) >> 13);\n'
                   '    std::ofstream f(argv[3], std::ios::binary);\n'
                   '    rtgo::write_png(f, rgb.data(), w, h);\n'
                   '    return 0;\n}\n' % root)
    exe = tmp_path / "png_test"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), str(src)])
    for w, h in ((1, 1), (13, 3), (400, 225)):
        out = tmp_path / f"cpp_{w}x{h}.png"
        subprocess.check_call([str(exe), str(w), str(h), str(out)])
        i = np.arange(w * h * 3, dtype=np.uint64)
        want = (((i * 2654435761) & 0xFFFFFFFF) >> 13).astype(np.uint8).reshape(h, w, 3)
        assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), want)
