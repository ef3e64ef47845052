"""TEST INFRASTRUCTURE: ctypes wrapper over oracle/liboracle.so (the CPU restatement of the
reference's hot path, oracle.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.  It shares nothing with the product but
the C struct layouts of include/rt_b200.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from raytracer_go_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

MODE_LINEAR, MODE_REF_BVH = 0, 1        # World.Hit (hittables.go:55) / BVH.Hit (bvh.go:220)
ORDER_RECURSIVE, ORDER_ITERATIVE = 0, 1  # ray.go:32-54 as written / unrolled front to back


class orc_stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("box_tests", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("hits", C.c_uint64), ("seconds", C.c_double),
                ("threads", C.c_int32), ("reserved", C.c_int32)]


def build(force=False):
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_camera_from_options.argtypes = [C.POINTER(abi.rt_camera_options), C.POINTER(abi.rt_camera)]
        L.orc_philox4x32_10.argtypes = [C.c_void_p] * 3
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_philox4x32.restype = None
        L.orc_set_philox_rounds.argtypes = [C.c_int]
        L.orc_rng_floats.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.orc_rng_floats.restype = None
        L.orc_trace.argtypes = [C.POINTER(abi.rt_scene_desc), C.c_int, C.c_uint64, C.c_void_p, C.c_void_p,
                                C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_hit_info.argtypes = [C.POINTER(abi.rt_scene_desc), C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_float, C.POINTER(C.c_int32), C.c_void_p]
        L.orc_aabb_hit.argtypes = [C.c_void_p] * 4 + [C.c_float, C.c_float]
        L.orc_reflect.argtypes = [C.c_void_p] * 3
        L.orc_reflect.restype = None
        L.orc_refract.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_refract.restype = None
        L.orc_reflectance.argtypes = [C.c_float, C.c_float]
        L.orc_reflectance.restype = C.c_float
        L.orc_texture.argtypes = [C.POINTER(abi.rt_scene_desc), C.c_uint32, C.c_float, C.c_float,
                                  C.c_void_p, C.c_void_p]
        L.orc_texture.restype = None
        L.orc_encode_pixel.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_encode_pixel.restype = None
        L.orc_scatter.argtypes = [C.POINTER(abi.rt_scene_desc), C.c_void_p, C.c_void_p, C.c_uint64,
                                  C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_scatter_fed.argtypes = [C.POINTER(abi.rt_scene_desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p]
        L.orc_get_color_fed.argtypes = [C.POINTER(abi.rt_scene_desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_get_ray_fed.argtypes = [C.POINTER(abi.rt_camera), C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p]
        L.orc_primary_rays.argtypes = [C.POINTER(abi.rt_camera), C.c_uint64, C.c_int32, C.c_int32,
                                       C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_render.argtypes = [C.POINTER(abi.rt_scene_desc), C.POINTER(abi.rt_camera), C.c_uint64,
                                 C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                 C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(orc_stats)]
        L.orc_resolve.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.orc_resolve.restype = None
        L.orc_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def hardware_threads():
    return int(lib().orc_hardware_threads())


def camera_from_options(opts):
    cam = abi.rt_camera()
    assert lib().orc_camera_from_options(C.byref(opts), C.byref(cam)) == 0
    return cam


STREAM_ROUNDS = 7  # Philox rounds of the render streams (csrc/rt_rng.h: RT_PHILOX_ROUNDS)


def philox(ctr, key, rounds=10):
    """Philox4x32 with `rounds` rounds (10 = Random123's default; the render streams use STREAM_ROUNDS)."""
    c = np.ascontiguousarray(ctr, np.uint32)
    k = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32(_p(c), _p(k), rounds, _p(out))
    return out


def set_philox_rounds(rounds):
    """Rounds of the oracle's render streams; returns the previous value."""
    return int(lib().orc_set_philox_rounds(rounds))


def rng_floats(seed, pixel, sample, n):
    out = np.zeros(n, np.float32)
    lib().orc_rng_floats(seed, pixel, sample, n, _p(out))
    return out


def trace(scene, origins, dirs, tmin=0.001, tmax=np.inf, mode=MODE_LINEAR, bvh_seed=1, threads=0):
    o, d = _f32(origins).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
    n = len(o)
    ids = np.empty(n, np.int32)
    ts = np.empty(n, np.float32)
    desc, keep = scene.to_desc()
    rc = lib().orc_trace(C.byref(desc), mode, bvh_seed, _p(o), _p(d), n, tmin, tmax, _p(ids), _p(ts),
                         threads or hardware_threads())
    assert rc == 0, "oracle rejected the scene"
    return ids, ts


def hit_info(scene, origin, direction, tmin=0.001, tmax=np.inf):
    """-> None or dict(id, point, normal, t, u, v, front_face) of World.Hit."""
    desc, keep = scene.to_desc()
    out = np.zeros(10, np.float32)
    oid = C.c_int32(-1)
    o, d = _f32(origin), _f32(direction)
    assert lib().orc_hit_info(C.byref(desc), _p(o), _p(d), tmin, tmax, C.byref(oid), _p(out)) == 0
    if oid.value < 0:
        return None
    return dict(id=oid.value, point=out[0:3].copy(), normal=out[3:6].copy(), t=out[6], u=out[7],
                v=out[8], front_face=bool(out[9]))


def aabb_hit(bmin, bmax, origin, direction, tmin, tmax):
    a, b, o, d = _f32(bmin), _f32(bmax), _f32(origin), _f32(direction)
    return bool(lib().orc_aabb_hit(_p(a), _p(b), _p(o), _p(d), tmin, tmax))


def reflect(v, n):
    out = np.zeros(3, np.float32)
    a, b = _f32(v), _f32(n)
    lib().orc_reflect(_p(a), _p(b), _p(out))
    return out


def refract(uv, n, eta):
    out = np.zeros(3, np.float32)
    a, b = _f32(uv), _f32(n)
    lib().orc_refract(_p(a), _p(b), eta, _p(out))
    return out


def reflectance(cos_theta, eta):
    return float(lib().orc_reflectance(cos_theta, eta))


def texture(scene, tex, u, v, p):
    desc, keep = scene.to_desc()
    out = np.zeros(3, np.float32)
    pp = _f32(p)
    lib().orc_texture(C.byref(desc), tex, u, v, _p(pp), _p(out))
    return out


def encode_pixel(mean):
    out = np.zeros(3, np.uint8)
    m = _f32(mean)
    lib().orc_encode_pixel(_p(m), _p(out))
    return out


def scatter(scene, origin, direction, seed=1, pixel=0, sample=0):
    """Material.Scatter at the World.Hit of the ray -> None (miss) or dict."""
    desc, keep = scene.to_desc()
    out = np.zeros(10, np.float32)
    o, d = _f32(origin), _f32(direction)
    rc = lib().orc_scatter(C.byref(desc), _p(o), _p(d), seed, pixel, sample, _p(out))
    if rc != 0:
        return None
    return dict(scattered=bool(out[0]), origin=out[1:4].copy(), dir=out[4:7].copy(),
                attenuation=out[7:10].copy())


def scatter_fed(scene, origin, direction, feed):
    """Material.Scatter + Emit at the World.Hit of the ray, uniforms from `feed` (four per block) -> None or dict."""
    desc, keep = scene.to_desc()
    out, emit = np.zeros(10, np.float32), np.zeros(3, np.float32)
    o, d, f = _f32(origin), _f32(direction), _f32(feed)
    rc = lib().orc_scatter_fed(C.byref(desc), _p(o), _p(d), _p(f), len(f), _p(out), _p(emit))
    if rc != 0:
        return None
    return dict(scattered=bool(out[0]), origin=out[1:4].copy(), dir=out[4:7].copy(), attenuation=out[7:10].copy(),
                emitted=emit)


def get_color_fed(scene, origin, direction, background, max_depth, feed, order=ORDER_RECURSIVE):
    """Ray.GetColor (ray.go:32-54) for an explicit ray, uniforms from `feed`."""
    desc, keep = scene.to_desc()
    out = np.zeros(3, np.float32)
    o, d, b, f = _f32(origin), _f32(direction), _f32(background), _f32(feed)
    assert lib().orc_get_color_fed(C.byref(desc), _p(o), _p(d), _p(b), max_depth, order, _p(f), len(f), _p(out)) == 0
    return out


def get_ray_fed(cam, i, j, feed):
    o, d, f = np.zeros(3, np.float32), np.zeros(3, np.float32), _f32(feed)
    assert lib().orc_get_ray_fed(C.byref(cam), i, j, _p(f), len(f), _p(o), _p(d)) == 0
    return o, d


def primary_rays(cam, seed, pixel_begin, n_pixels, sample_offset=0, sample_count=1):
    n = n_pixels * sample_count
    o = np.empty((n, 3), np.float32)
    d = np.empty((n, 3), np.float32)
    assert lib().orc_primary_rays(C.byref(cam), seed, sample_offset, sample_count, pixel_begin,
                                  n_pixels, _p(o), _p(d)) == 0
    return o, d


def render(scene, cam, seed, sample_offset=0, sample_count=0, total_spp=0, mode=MODE_LINEAR,
           order=ORDER_RECURSIVE, bvh_seed=1, threads=0, rows=None):
    """-> (rgb uint8 (H,W,3), accum float32 (H,W,3), orc_stats)."""
    desc, keep = scene.to_desc()
    H, W = cam.height, cam.width
    rgb = np.zeros((H, W, 3), np.uint8)
    acc = np.zeros((H, W, 3), np.float32)
    st = orc_stats()
    r0, r1 = rows if rows else (0, 0)
    rc = lib().orc_render(C.byref(desc), C.byref(cam), seed, sample_offset, sample_count, total_spp,
                          mode, order, bvh_seed, threads or hardware_threads(), r0, r1, _p(rgb), _p(acc),
                          C.byref(st))
    assert rc == 0, "oracle rejected the scene"
    return rgb, acc, st


def resolve(accum, total_spp):
    a = _f32(accum).reshape(-1, 3)
    out = np.zeros((len(a), 3), np.uint8)
    lib().orc_resolve(_p(a), len(a), total_spp, _p(out))
    return out.reshape(np.asarray(accum).shape)
