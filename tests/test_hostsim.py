"""Pre-GPU logic checks: the per-ray device headers (rt_trace.h, rt_shade.h, rt_rng.h) and the host
BVH builder compiled for the host (tests/hostsim, TEST-ONLY) against the oracle.  These catch
traversal / shading / flattening bugs in this GPU-less container; the real parity tests are the
`-m gpu` ones, which run the CUDA kernels through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from raytracer_go_b200 import abi, scenes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hs():
    d = os.path.join(HERE, "hostsim")
    subprocess.check_call(["make", "-s", "-C", d, "libhostsim.so"])
    return C.CDLL(os.path.join(d, "libhostsim.so"))


def hs_trace(hs, scene, o, d, max_leaf=4, radius=0.0, tmin=0.001, tmax=np.inf):
    desc, keep = scene.to_desc()
    o, d = np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32)
    ids, ts = np.empty(len(o), np.int32), np.empty(len(o), np.float32)
    bt, st = C.c_uint64(), C.c_uint64()
    hs.hs_trace(C.byref(desc), max_leaf, C.c_float(radius), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                C.c_int64(len(o)), C.c_float(tmin), C.c_float(tmax), ids.ctypes.data_as(C.c_void_p),
                ts.ctypes.data_as(C.c_void_p), C.byref(bt), C.byref(st))
    return ids, ts, bt.value, st.value


@pytest.mark.parametrize("max_leaf", [1, 2, 4, 8])
def test_flat_bvh_traversal_equals_world_hit(hs, orc, random_scene, max_leaf):
    cam = orc.camera_from_options(scenes.camera_options(240, 2))
    ro, rd = orc.primary_rays(cam, 5, 0, cam.width * cam.height, 0, 2)
    ids, ts, bt, st = hs_trace(hs, random_scene, ro, rd, max_leaf)
    rids, rts = orc.trace(random_scene, ro, rd)
    assert np.array_equal(ids, rids)
    assert np.array_equal(ts[rids >= 0].view(np.uint32), rts[rids >= 0].view(np.uint32))
    assert bt / len(ro) < 40 and st / len(ro) < 8  # the tree actually culls


def test_traversal_on_secondary_like_rays(hs, orc, random_scene):
    rng = np.random.default_rng(11)
    n = 200_000
    sp = random_scene.spheres
    pick = rng.integers(1, len(sp), n)
    c = np.stack([sp["cx"][pick], sp["cy"][pick], sp["cz"][pick]], -1)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    o = (c + v * (sp["r"][pick] * rng.choice([1.0, 1.0, 2.0, 0.5], n))[:, None]).astype(np.float32)
    d = (rng.normal(size=(n, 3)) * rng.choice([0.2, 1, 9], n)[:, None]).astype(np.float32)
    ids, ts, _, _ = hs_trace(hs, random_scene, o, d)
    rids, rts = orc.trace(random_scene, o, d)
    assert np.array_equal(ids, rids)
    assert np.array_equal(ts[rids >= 0].view(np.uint32), rts[rids >= 0].view(np.uint32))


def test_axis_parallel_and_degenerate_rays(hs, orc, random_scene):
    """Zero direction components must not break the fused slab test (cull_rcp)."""
    o = np.array([[0, 5, 0], [4, 5, 0], [-4, 1, 9], [0.3, 0.2, 20], [0, 1, 0], [13, 2, 3], [0, 5, 0]], np.float32)
    d = np.array([[0, -1, 0], [0, -3, 0], [0, 0, -1], [0, 0, -2], [1, 0, 0], [0, 0, 0], [0, -0.0, -0.0]], np.float32)
    ids, ts, _, _ = hs_trace(hs, random_scene, o, d)
    rids, rts = orc.trace(random_scene, o, d)
    assert np.array_equal(ids, rids)
    assert ids[0] >= 0 and ids[1] >= 0 and ids[2] >= 0 and ids[5] == -1
    assert np.array_equal(ts[rids >= 0], rts[rids >= 0])


def test_ties_and_small_scenes(hs, orc):
    tex, mat = np.zeros(1, scenes.TEXTURE_DT), np.zeros(1, scenes.MATERIAL_DT)
    for n in (1, 2, 3, 9, 17):
        sph = np.zeros(n, scenes.SPHERE_DT)
        for k in range(n):
            sph[k] = (0, 0, -1, 0.5, 0)  # n coincident spheres: object 0 must win
        s = scenes.SceneData(sph, mat, tex)
        ids, ts, _, _ = hs_trace(hs, s, [(0, 0, 0), (0, 0, -1)], [(0, 0, -1), (0, 1, 0)])
        assert ids.tolist() == [0, 0] and ts.tolist() == [0.5, 0.5]
    empty = scenes.SceneData(np.zeros(0, scenes.SPHERE_DT), mat, tex)
    ids, _, _, _ = hs_trace(hs, empty, [(0, 0, 0)], [(0, 0, -1)])
    assert ids[0] == -1


def test_padding_grows_with_origin_radius(hs, random_scene):
    desc, keep = random_scene.to_desc()
    out = []
    for radius in (0.0, 30.0, 300.0):
        nn, ns, md, p0, p1 = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_float(), C.c_float()
        hs.hs_bvh_stats(C.byref(desc), 4, C.c_float(radius), C.byref(nn), C.byref(ns), C.byref(md), C.byref(p0),
                        C.byref(p1))
        out.append((nn.value, ns.value, md.value, p0.value, p1.value))
    assert out[0][1] == len(random_scene.spheres) and out[0][2] < 32
    assert out[2][4] > out[1][4] > 0 and out[1][3] > 0


def test_megakernel_path_logic_equals_oracle(hs, orc, random_scene):
    """generate_ray + trace + shade_hit + resolve in the kernel's order == oracle (iterative)."""
    cam = orc.camera_from_options(scenes.camera_options(96, 3))
    desc, keep = random_scene.to_desc()
    rgb = np.zeros((cam.height, cam.width, 3), np.uint8)
    acc = np.zeros((cam.height, cam.width, 3), np.float32)
    hs.hs_render(C.byref(desc), C.byref(cam), C.c_uint64(42), 2, 3, 3, 4, rgb.ctypes.data_as(C.c_void_p),
                 acc.ctypes.data_as(C.c_void_p))
    rrgb, racc, _ = orc.render(random_scene, cam, 42, sample_offset=2, sample_count=3, order=orc.ORDER_ITERATIVE)
    assert np.array_equal(acc.view(np.uint32), racc.view(np.uint32))
    assert np.array_equal(rgb, rrgb)


def test_image_texture_path_equals_oracle(hs, orc):
    s = scenes.earth_scene()
    cam = orc.camera_from_options(scenes.camera_options(96, 2, look_from=(0, 0, -12), defocus_deg=0.0))
    desc, keep = s.to_desc()
    rgb = np.zeros((cam.height, cam.width, 3), np.uint8)
    acc = np.zeros((cam.height, cam.width, 3), np.float32)
    hs.hs_render(C.byref(desc), C.byref(cam), C.c_uint64(7), 0, 2, 2, 4, rgb.ctypes.data_as(C.c_void_p),
                 acc.ctypes.data_as(C.c_void_p))
    rrgb, racc, _ = orc.render(s, cam, 7, order=orc.ORDER_ITERATIVE)
    assert np.array_equal(acc.view(np.uint32), racc.view(np.uint32))
    # the out-of-bounds colour (0, 0.529, 0) shows up on ~5/24 of the longitudes
    green = (racc[..., 1] > 0.3) & (racc[..., 0] == 0) & (racc[..., 2] == 0)
    assert green.mean() > 0.005


def _hs_render(hs, scene, cam, seed, spp):
    desc, keep = scene.to_desc()
    rgb = np.zeros((cam.height, cam.width, 3), np.uint8)
    acc = np.zeros((cam.height, cam.width, 3), np.float32)
    hs.hs_render(C.byref(desc), C.byref(cam), C.c_uint64(seed), 0, spp, spp, 4, rgb.ctypes.data_as(C.c_void_p),
                 acc.ctypes.data_as(C.c_void_p))
    return rgb, acc


@pytest.mark.parametrize("name", ["cornell", "quads", "mixed"])
def test_quad_scenes_equal_oracle(hs, orc, name):
    """Quad / Box / DiffuseLight scenes (main.go:132-160, 194-225) and a mixed sphere+quad world:
    closest hits (object IDs across both kinds) and the rendered sums equal the oracle's."""
    if name == "cornell":
        s, o = scenes.cornell_box_scene(), scenes.cornell_camera_options(64, 8)
    elif name == "quads":
        s, o = scenes.quad_demo_scene(), scenes.quad_demo_camera_options(96, 4)
    else:
        s, o = scenes.mixed_scene(), scenes.cornell_camera_options(64, 8)
    cam = orc.camera_from_options(o)
    ro, rd = orc.primary_rays(cam, 5, 0, cam.width * cam.height, 0, 1)
    rng = np.random.default_rng(2)
    so = rng.uniform(5, 550, size=(20000, 3)).astype(np.float32) if name != "quads" else rng.uniform(-3, 5, size=(20000, 3)).astype(np.float32)
    sd = rng.normal(size=(20000, 3)).astype(np.float32)
    o_all, d_all = np.concatenate([ro, so]), np.concatenate([rd, sd])
    ids, ts, _, _ = hs_trace(hs, s, o_all, d_all)
    rids, rts = orc.trace(s, o_all, d_all)
    assert np.array_equal(ids, rids)
    assert np.array_equal(ts[rids >= 0].view(np.uint32), rts[rids >= 0].view(np.uint32))
    assert (rids >= 0).mean() > 0.15
    rgb, acc = _hs_render(hs, s, cam, 11, cam.spp)
    rrgb, racc, _ = orc.render(s, cam, 11, order=orc.ORDER_ITERATIVE)
    assert np.array_equal(acc.view(np.uint32), racc.view(np.uint32))
    assert np.array_equal(rgb, rrgb)
    if name != "quads":
        assert acc.max() > 5  # the light is seen


def test_quad_edge_cases(hs, orc):
    """Edges are inside (alpha, beta in [0,1]), parallel rays miss, coincident quads: lower ID wins."""
    tex, mat = np.zeros(1, scenes.TEXTURE_DT), np.zeros(1, scenes.MATERIAL_DT)
    q = np.zeros(3, scenes.QUAD_DT)
    q[0] = ((-1, -1, -2), (2, 0, 0), (0, 2, 0), 0)
    q[1] = ((-1, -1, -2), (2, 0, 0), (0, 2, 0), 0)   # coincident twin
    q[2] = ((-1, -1, -5), (2, 0, 0), (0, 2, 0), 0)
    s = scenes.SceneData(np.zeros(0, scenes.SPHERE_DT), mat, tex, quads=q)
    o = np.zeros((6, 3), np.float32)
    d = np.array([[0, 0, -1], [1, 1, -2], [1.001, 0, -2], [1, 0, 0], [0, 0, 1], [2.4, 2.4, -5]], np.float32)
    ids, ts, _, _ = hs_trace(hs, s, o, d)
    rids, rts = orc.trace(s, o, d)
    assert np.array_equal(ids, rids) and ids.tolist() == [0, 0, -1, -1, -1, 0]
    assert ts[0] == 2.0 and ts[1] == 1.0 and np.array_equal(ts[rids >= 0], rts[rids >= 0])


@pytest.mark.parametrize("name", ["perlin", "simple-light"])
def test_noise_texture_scenes_equal_oracle(hs, orc, name):
    """Perlin / NoiseTexture (materials.go:195-295; main.go:106-130, 162-192)."""
    if name == "perlin":
        s, o = scenes.perlin_demo_scene(), scenes.perlin_camera_options(96, 4)
    else:
        s, o = scenes.simple_light_scene(), scenes.simple_light_camera_options(96, 8)
    cam = orc.camera_from_options(o)
    rgb, acc = _hs_render(hs, s, cam, 21, cam.spp)
    rrgb, racc, _ = orc.render(s, cam, 21, order=orc.ORDER_ITERATIVE)
    assert np.array_equal(acc.view(np.uint32), racc.view(np.uint32))
    assert np.array_equal(rgb, rrgb)
    assert len(np.unique(rgb.reshape(-1, 3), axis=0)) > 50   # marble, not a flat colour


def _hollow_glass_scene():
    """The book's hollow glass sphere: a dielectric sphere with NEGATIVE radius inside a positive one
    (hittables.go:119: the normal is (p - c) * r, so a negative radius flips it)."""
    tex = np.zeros(2, scenes.TEXTURE_DT)
    tex[0]["kind"], tex[0]["a"], tex[0]["b"], tex[0]["scale"] = abi.RT_TEX_CHECKER, (.2, .3, .1), (.9, .9, .9), 0.32
    tex[1]["a"] = (0.1, 0.2, 0.5)
    mat = np.zeros(3, scenes.MATERIAL_DT)
    mat[0]["kind"], mat[0]["texture"] = abi.RT_MAT_LAMBERTIAN, 0
    mat[1]["kind"], mat[1]["ior"] = abi.RT_MAT_DIELECTRIC, 1.5
    mat[2]["kind"], mat[2]["texture"] = abi.RT_MAT_LAMBERTIAN, 1
    sph = np.zeros(5, scenes.SPHERE_DT)
    sph[0] = (0, -100.5, -1, 100, 0)
    sph[1] = (-1, 0, -1, 0.5, 1)
    sph[2] = (-1, 0, -1, -0.4, 1)     # negative radius: the inner surface of the shell
    sph[3] = (0, 0, -1, 0.5, 2)
    sph[4] = (1, 0, -1, 0.0, 2)       # zero radius: never hit
    return scenes.SceneData(sph, mat, tex, name="hollow-glass")


def test_negative_and_zero_radius_spheres(hs, orc):
    s = _hollow_glass_scene()
    cam = orc.camera_from_options(scenes.camera_options(96, 6, look_from=(-2, 2, 1), look_at=(0, 0, -1), vfov_deg=40,
                                                        defocus_deg=0.0, focus_dist=1.0))
    ro, rd = orc.primary_rays(cam, 3, 0, cam.width * cam.height, 0, 2)
    ids, ts, _, _ = hs_trace(hs, s, ro, rd)
    rids, rts = orc.trace(s, ro, rd)
    assert np.array_equal(ids, rids) and np.array_equal(ts[rids >= 0], rts[rids >= 0])
    assert (rids == 1).any() and not (rids == 4).any()
    inside_o = np.tile(np.array([[-1, 0, -1]], np.float32), (64, 1))          # from inside the shell
    inside_d = np.random.default_rng(1).normal(size=(64, 3)).astype(np.float32)
    i2, t2, _, _ = hs_trace(hs, s, inside_o, inside_d)
    r2, rt2 = orc.trace(s, inside_o, inside_d)
    assert np.array_equal(i2, r2) and (r2 == 2).all() and np.array_equal(t2, rt2)
    rgb, acc = _hs_render(hs, s, cam, 9, cam.spp)
    rrgb, racc, _ = orc.render(s, cam, 9, order=orc.ORDER_ITERATIVE)
    assert np.array_equal(acc.view(np.uint32), racc.view(np.uint32)) and np.array_equal(rgb, rrgb)


def bvh_hash(hs, scene, radius=0.0, refit=0.0, threads=None, max_leaf=4):
    hs.hs_bvh_hash.restype = C.c_uint64
    desc, keep = scene.to_desc()
    old = os.environ.pop("RT_B200_BVH_THREADS", None)
    if threads:
        os.environ["RT_B200_BVH_THREADS"] = str(threads)
    try:
        return hs.hs_bvh_hash(C.byref(desc), max_leaf, C.c_float(radius), C.c_float(refit))
    finally:
        os.environ.pop("RT_B200_BVH_THREADS", None)
        if old is not None:
            os.environ["RT_B200_BVH_THREADS"] = old


def test_parallel_bvh_build_is_deterministic(hs):
    """The host build splits and emits subtrees on several threads (bvh_build.cpp); every array of the
    flattened tree must be byte-identical for any thread count, after a refit too.  160 K spheres is above
    both parallel thresholds (PAR_RANGE, PAR_SUBTREE)."""
    scene = scenes.stress_scene(200)
    ref = bvh_hash(hs, scene, 150.0, threads=1)
    for t in (2, 3, 8, None):
        assert bvh_hash(hs, scene, 150.0, threads=t) == ref
    ref_refit = bvh_hash(hs, scene, 150.0, refit=300.0, threads=1)
    assert ref_refit != ref
    assert bvh_hash(hs, scene, 150.0, refit=300.0, threads=8) == ref_refit


def test_bvh_layout_is_the_depth_first_build(hs, random_scene):
    """Pins the flattened tree of the C2 scene (nodes, device nodes, slots, metadata): the hash was taken
    from the single-threaded recursive builder that wrote the arrays in depth-first order as it went."""
    assert bvh_hash(hs, random_scene, threads=1) == 0x145DF021C2BFCBD8
    assert bvh_hash(hs, random_scene) == 0x145DF021C2BFCBD8
    assert bvh_hash(hs, scenes.cornell_box_scene()) == 0xAFF672732185BF18


def geometric_scene(n=240, ratio=1.3):
    """Spheres whose positions and radii grow geometrically: binned SAH peels one of them off per level."""
    tex = np.zeros(1, scenes.TEXTURE_DT)
    mat = np.zeros(1, scenes.MATERIAL_DT)
    sph = np.zeros(n, scenes.SPHERE_DT)
    x = ratio ** np.arange(n, dtype=np.float64)
    sph["cx"], sph["cy"], sph["cz"] = x, 0.1 * x, -0.05 * x
    sph["r"] = 0.1 * x
    return scenes.SceneData(sph, mat, tex, ray_origin_radius=1.0, name="geometric")


def test_bvh_depth_stays_within_the_traversal_stack(hs, orc):
    """A scene that would make a pure SAH tree one level deep per primitive: below inner depth 30 the builder
    halves ranges by count, so the chain of inner nodes fits the kernels' 64-entry traversal stack, and the
    traversal still returns World.Hit's answer."""
    s = geometric_scene()
    desc, keep = s.to_desc()
    nn, ns, md, p0, p1 = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_float(), C.c_float()
    hs.hs_bvh_stats(C.byref(desc), 1, C.c_float(1.0), C.byref(nn), C.byref(ns), C.byref(md), C.byref(p0), C.byref(p1))
    assert ns.value == len(s.spheres)
    assert 30 < md.value <= 62, md.value
    rng = np.random.default_rng(3)
    n = 20000
    pick = rng.integers(0, len(s.spheres), n)
    c = np.stack([s.spheres["cx"][pick], s.spheres["cy"][pick], s.spheres["cz"][pick]], -1).astype(np.float64)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    o = (c + v * (s.spheres["r"][pick] * 1.0000001)[:, None]).astype(np.float32)  # on the surfaces
    d = rng.normal(size=(n, 3)).astype(np.float32)
    ids, ts, _, _ = hs_trace(hs, s, o, d, max_leaf=1, radius=1.0)
    rids, rts = orc.trace(s, o, d)
    near = (rids < 0) | (rts < 1e3)  # inside the envelope the padding is exact for
    assert np.array_equal(ids[near], rids[near])


def test_fused_range_mappings_are_exact():
    """rt_rng.h maps a Philox word to [-1, 1) / [-0.5, 0.5) with ONE fused multiply-add where the reference
    multiplies and then adds (math.go:30-32, camera.go:290): every product involved is exact (k * 2^-24 with
    k < 2^24, times 2), so both round the same real number once.  All 2^24 values of k, bit for bit."""
    k = np.arange(1 << 24, dtype=np.uint32)
    r = k.astype(np.float32) * np.float32(2.0 ** -24)                      # Float32()
    ref = np.float32(-1.0) + r * (np.float32(1.0) - np.float32(-1.0))     # lo + r * (hi - lo), unfused float32
    fused = (k.astype(np.float64) * 2.0 ** -23 - 1.0).astype(np.float32)  # fmaf: exact product, one rounding
    assert np.array_equal(ref.view(np.uint32), fused.view(np.uint32))
    ref = np.float32(-0.5) + r
    fused = (k.astype(np.float64) * 2.0 ** -24 - 0.5).astype(np.float32)
    assert np.array_equal(ref.view(np.uint32), fused.view(np.uint32))


@pytest.mark.parametrize("seed", range(8))
def test_random_scenes_traversal_equals_world_hit(hs, orc, seed):
    """Fuzz of the builder + flattened traversal against World.Hit on scenes the generators never make
    (tests/fuzz_scenes.py), every leaf size."""
    from tests.fuzz_scenes import fuzz_scene_and_rays
    s, o, d, radius = fuzz_scene_and_rays(seed)
    rids, rts = orc.trace(s, o, d)
    for max_leaf in (1, 4, 8):
        ids, ts, _, _ = hs_trace(hs, s, o, d, max_leaf=max_leaf, radius=radius)
        assert np.array_equal(ids, rids), (seed, max_leaf, int((ids != rids).sum()))
        hit = rids >= 0
        assert np.array_equal(ts[hit].view(np.uint32), rts[hit].view(np.uint32))


def hs_trace_leaf_start(hs, scene, o, d, start_slots, max_leaf=4, radius=0.0, tmin=0.001, tmax=np.inf):
    desc, keep = scene.to_desc()
    o, d = np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32)
    st_ = np.ascontiguousarray(start_slots, np.uint32)
    ids, ts = np.empty(len(o), np.int32), np.empty(len(o), np.float32)
    bt, st = C.c_uint64(), C.c_uint64()
    hs.hs_trace_leaf_start(C.byref(desc), max_leaf, C.c_float(radius), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                           st_.ctypes.data_as(C.c_void_p), C.c_int64(len(o)), C.c_float(tmin), C.c_float(tmax),
                           ids.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p), C.byref(bt), C.byref(st))
    return ids, ts, bt.value, st.value


@pytest.mark.parametrize("max_leaf", [1, 4, 8])
def test_leaf_start_from_any_leaf_equals_world_hit(hs, orc, random_scene, max_leaf):
    """Leaf start (rt_trace.h): the chain of a leaf plus the leaf cover the whole tree, so the closest hit does not
    depend on WHICH leaf a ray starts from — random start slots must reproduce World.Hit bit for bit."""
    rng = np.random.default_rng(21)
    n = 100_000
    sp = random_scene.spheres
    pick = rng.integers(0, len(sp), n)
    c = np.stack([sp["cx"][pick], sp["cy"][pick], sp["cz"][pick]], -1)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    o = (c + v * (np.minimum(sp["r"][pick], 3.0) * rng.choice([1.0, 1.0, 2.0, 0.5], n))[:, None]).astype(np.float32)
    d = (rng.normal(size=(n, 3)) * rng.choice([0.2, 1, 9], n)[:, None]).astype(np.float32)
    rids, rts = orc.trace(random_scene, o, d)
    ids, ts, bt_any, _ = hs_trace_leaf_start(hs, random_scene, o, d, rng.integers(0, 1 << 30, n), max_leaf)
    assert np.array_equal(ids, rids)
    assert np.array_equal(ts[rids >= 0].view(np.uint32), rts[rids >= 0].view(np.uint32))


def test_leaf_start_on_quads_and_fuzz_scenes(hs, orc):
    from tests.fuzz_scenes import fuzz_scene_and_rays
    rng = np.random.default_rng(5)
    for seed in range(6):
        s, o, d, radius = fuzz_scene_and_rays(seed, m=3000)
        ids, ts, _, _ = hs_trace_leaf_start(hs, s, o, d, rng.integers(0, 1 << 30, len(o)), 4, radius)
        rids, rts = orc.trace(s, o, d)
        assert np.array_equal(ids, rids), seed
        assert np.array_equal(ts[rids >= 0].view(np.uint32), rts[rids >= 0].view(np.uint32))
    s = scenes.mixed_scene()
    n = 20_000
    o = rng.uniform(-4, 4, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    ids, ts, _, _ = hs_trace_leaf_start(hs, s, o, d, rng.integers(0, 1 << 30, n))
    rids, rts = orc.trace(s, o, d)
    assert np.array_equal(ids, rids)
    assert np.array_equal(ts[rids >= 0].view(np.uint32), rts[rids >= 0].view(np.uint32))


def hs_pixel_candidates(hs, scene, opts, n_pixels, stride, spp, begin=0, cap=15):
    from oracle import pyoracle as orc
    cam = orc.camera_from_options(opts)  # (from the oracle: this test needs no GPU library)
    desc, keep = scene.to_desc()
    out = (C.c_double * 6)()
    rc = hs.hs_pixel_candidates(C.byref(desc), C.byref(cam), C.c_uint64(123), C.c_int64(begin), C.c_int64(n_pixels),
                                C.c_int64(stride), spp, 4, cap, out)
    assert rc == 0
    return dict(pixels=out[0], mean=out[1] / max(1.0, out[0]), longest=out[2], missing=out[3], differ=out[4], overflow=out[5])


@pytest.mark.parametrize("name", ["random", "random_wide", "cornell", "mixed", "perlin", "earth", "no_defocus"])
def test_pixel_candidate_lists_are_supersets(hs, name):
    """Per-pixel candidate lists of the primary stage (pixel_beam + beam_candidates, the functions
    pixel_candidates_kernel calls): for every camera ray drawn for a pixel the closest hit is in the pixel's list, and
    trace_candidates over the list returns the tree traversal's answer bit for bit.  Also pins the list statistics the
    design relies on (a handful of candidates per pixel, overflows rare)."""
    opts, n, stride, spp = scenes.camera_options(400, 1), 400 * 225 // 7, 7, 8
    if name == "random":
        s = scenes.random_scene()
    elif name == "random_wide":     # 90 degrees: direction intervals straddle zero on two axes in the image centre
        s, opts = scenes.random_scene(), scenes.camera_options(400, 1, vfov_deg=90.0, look_from=(0.5, 0.3, 0.5), look_at=(0.5, 0.3, -5))
    elif name == "no_defocus":
        s, opts = scenes.random_scene(), scenes.camera_options(400, 1, defocus_deg=0.0)
    elif name == "cornell":
        s, opts, n, stride = scenes.cornell_box_scene(), scenes.cornell_camera_options(200, 1), 200 * 200 // 3, 3
    elif name == "mixed":
        s = scenes.mixed_scene()
    elif name == "perlin":
        s, opts = scenes.perlin_demo_scene(), scenes.perlin_camera_options(400, 1)
    else:
        s, opts = scenes.earth_scene(), scenes.camera_options(400, 1, look_from=(0, 0, -12), defocus_deg=0.0)
    r = hs_pixel_candidates(hs, s, opts, n, stride, spp)
    assert r["pixels"] == n and r["missing"] == 0 and r["differ"] == 0
    if name == "random":
        assert 3.5 < r["mean"] < 6.0 and r["longest"] <= 24 and r["overflow"] < 0.01 * n


def test_beam_test_is_conservative_for_single_boxes(hs):
    """The property the candidate lists rest on, one box at a time: if the kernels' slab test accepts ANY ray drawn from a
    beam, the beam test accepts the box.  800 000 random (box, beam) pairs — flat boxes, the huge ground box, beams as
    narrow as a pixel's and wide ones, direction intervals that straddle zero — 24 rays each, corners included."""
    hs.hs_beam_box_property.restype = C.c_int64
    total = 0
    for seed in range(4):
        acc = C.c_int64()
        assert hs.hs_beam_box_property(C.c_uint64(seed), C.c_int64(200_000), 24, C.byref(acc)) == 0
        total += acc.value
    assert total > 200_000          # the rays do hit their boxes in a third of the pairs: the property is exercised
