"""Known-answer tests that pin the CPU oracle.

The reference ships no tests or golden vectors, so each expected value here is derived by hand
from the reference source (citation beside it).  These are what make the oracle trustworthy as
the checker of the CUDA path; the same cases run on the device in tests/test_gpu_parity.py.
"""
import math

import numpy as np
import pytest

from raytracer_go_b200 import abi, philox, scenes

F = np.float32


def one_sphere(c=(0, 0, -1), r=0.5, n=1):
    tex = np.zeros(1, scenes.TEXTURE_DT)
    mat = np.zeros(1, scenes.MATERIAL_DT)
    sph = np.zeros(n, scenes.SPHERE_DT)
    for k in range(n):
        sph[k] = (*c, r, 0)
    return scenes.SceneData(sph, mat, tex)


def test_philox_random123_vectors(orc):
    """Random123 kat_vectors for philox4x32 with 10 rounds (its default; the scene generators use it) and with 7
    rounds (the render streams, csrc/rt_rng.h): zero, all-ones and pi-digit counters / keys."""
    zero, ones = ([0, 0, 0, 0], [0, 0]), ([0xffffffff] * 4, [0xffffffff] * 2)
    pi = ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])
    kats = [
        (10, *zero, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        (10, *ones, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        (10, *pi, [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
        (7, *zero, [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]),
        (7, *ones, [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]),
        (7, *pi, [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]),
    ]
    for rounds, ctr, key, want in kats:
        assert orc.philox(ctr, key, rounds).tolist() == want
        assert philox.philox4x32_10(np.array(ctr, np.uint32), key, rounds).tolist() == want
    assert orc.STREAM_ROUNDS == 7


def test_rng_stream_layout(orc):
    """Stream = blocks philox(ctr=(pixel, sample, b, 0)), floats = (w >> 8) * 2^-24 in [0,1)."""
    seed = 0x0123456789ABCDEF
    f = orc.rng_floats(seed, 77, 5, 10)
    want = []
    for blk in range(3):
        w = orc.philox([77, 5, blk, 0], [seed & 0xFFFFFFFF, seed >> 32], orc.STREAM_ROUNDS)
        want += [F(int(x) >> 8) * F(2.0 ** -24) for x in w]
    assert f.tolist() == [float(x) for x in want[:10]]
    assert (f >= 0).all() and (f < 1).all()


def test_seven_round_streams_pass_basic_statistics(orc):
    """Not a substitute for BigCrush (Random123 reports Philox4x32-7 as Crush-resistant) — a guard against a gross
    mistake in how the 7-round streams are keyed: uniforms of neighbouring pixels / samples / blocks must look
    independent.  4-sigma bounds on the mean, a 256-bin chi-square, and correlations along each counter axis."""
    seed = 0xC0FFEE
    pix = np.stack([orc.rng_floats(seed, p, 0, 64) for p in range(4096)])          # consecutive pixels, sample 0
    smp = np.stack([orc.rng_floats(seed, 1234, k, 64) for k in range(4096)])       # one pixel, consecutive samples
    for u in (pix, smp):
        n = u.size
        assert abs(float(u.mean()) - 0.5) < 4 * np.sqrt(1 / 12 / n)
        hist = np.bincount((u.ravel() * 256).astype(int), minlength=256)
        chi2 = float(((hist - n / 256) ** 2 / (n / 256)).sum())
        assert 255 - 5 * np.sqrt(2 * 255) < chi2 < 255 + 5 * np.sqrt(2 * 255)
        c = u - 0.5
        bound = 5 / 12 / np.sqrt(c[1:].size)                                        # 5 sigma of a product of two uniforms
        assert abs(float((c[1:] * c[:-1]).mean())) < bound                          # along the pixel / sample axis
        assert abs(float((c[:, 1:] * c[:, :-1]).mean())) < bound                    # along the word / block axis
        assert abs(float((c[:, 4:] * c[:, :-4]).mean())) < bound                    # same word of consecutive blocks
    assert abs(float(((pix - 0.5) * (smp - 0.5)).mean())) < 5 / 12 / np.sqrt(pix.size)


def test_sphere_front_hit(orc):
    """hittables.go:97-126: a=1, halfB=-1, c=0.75, disc=0.25, t=0.5; u=11/24, v=0.5."""
    h = orc.hit_info(one_sphere(), (0, 0, 0), (0, 0, -1))
    assert h["id"] == 0 and h["t"] == 0.5 and h["front_face"]
    assert h["point"].tolist() == [0, 0, -0.5] and h["normal"].tolist() == [0, 0, 1]
    assert h["u"] == pytest.approx(11 / 24, abs=1e-6) and h["v"] == pytest.approx(0.5, abs=1e-6)


def test_sphere_from_inside(orc):
    """Near root -0.5 fails the strict interval, far root 0.5 is taken; normal flipped."""
    h = orc.hit_info(one_sphere(), (0, 0, -1), (0, 0, -1))
    assert h["t"] == 0.5 and not h["front_face"]
    assert h["normal"].tolist() == [0, 0, 1]  # outward (0,0,-1) flipped against the ray


def test_strict_interval(orc):
    """bvh.go:18-20: min < v && v < max — a root equal to either end is rejected."""
    s = one_sphere()
    assert orc.hit_info(s, (0, 0, 0), (0, 0, -1), 0.001, 0.5) is None
    assert orc.hit_info(s, (0, 0, 0), (0, 0, -1), 0.001, np.nextafter(F(0.5), F(1)))["t"] == 0.5
    assert orc.hit_info(s, (0, 0, 0), (0, 0, -1), 0.5, np.inf)["t"] == 1.5
    assert orc.hit_info(s, (0, 0, 0), (0, 0, -1), 1.5, np.inf) is None


def test_tie_break_first_object_wins(orc):
    """hittables.go:59-69: a later object with an equal t fails the strict `<`."""
    s = one_sphere(n=3)
    ids, ts = orc.trace(s, [(0, 0, 0)], [(0, 0, -1)])
    assert ids[0] == 0 and ts[0] == 0.5


def test_aabb_axis_parallel(orc):
    """bvh.go:84-101 with dir.x = 0: invD = +Inf."""
    lo, hi = (-1, -1, -1), (1, 1, 1)
    assert orc.aabb_hit(lo, hi, (0, 0, 5), (0, 0, -1), 0.001, np.inf)        # inside the x,y slabs
    assert not orc.aabb_hit(lo, hi, (2, 0, 5), (0, 0, -1), 0.001, np.inf)    # outside the x slab
    # on the boundary: (min - o) * inf = NaN compares false -> interval unchanged on that side
    assert orc.aabb_hit(lo, hi, (-1, 0, 5), (0, 0, -1), 0.001, np.inf)
    assert not orc.aabb_hit(lo, hi, (0, 0, 5), (0, 0, 1), 0.001, np.inf)     # pointing away


def test_reflect_refract(orc):
    r2 = 1 / math.sqrt(2)
    np.testing.assert_allclose(orc.reflect((r2, -r2, 0), (0, 1, 0)), (r2, r2, 0), atol=1e-7)
    np.testing.assert_allclose(orc.refract((0, -1, 0), (0, 1, 0), 1 / 1.5), (0, -1, 0), atol=1e-7)
    # 45 degrees into glass: sin t = sin 45 / 1.5
    out = orc.refract((r2, -r2, 0), (0, 1, 0), 1 / 1.5)
    assert out[0] == pytest.approx(r2 / 1.5, abs=1e-6)
    assert np.linalg.norm(out) == pytest.approx(1.0, abs=1e-6)


def test_schlick(orc):
    """materials.go:115-119."""
    assert orc.reflectance(1.0, 1 / 1.5) == pytest.approx(0.04, abs=1e-7)
    assert orc.reflectance(0.0, 1 / 1.5) == pytest.approx(1.0, abs=1e-7)
    assert orc.reflectance(1.0, 1.5) == pytest.approx(0.04, abs=1e-7)


def test_checker(orc):
    """materials.go:127-137; Go's % keeps the sign so odd negative sums are 'odd'."""
    s = scenes.random_scene()
    even, odd = [F(0.2), F(0.3), F(0.1)], [F(0.9)] * 3
    assert orc.texture(s, 0, 0, 0, (0.1, 0.1, 0.1)).tolist() == even      # 0+0+0
    assert orc.texture(s, 0, 0, 0, (-0.1, 0.1, 0.1)).tolist() == odd      # -1+0+0
    assert orc.texture(s, 0, 0, 0, (-0.1, -0.1, 0.1)).tolist() == even    # -2
    assert orc.texture(s, 0, 0, 0, (0.33, 0.1, 0.1)).tolist() == odd      # 1


def test_pixel_encode(orc):
    """vec3.go:162-166 sqrt, 145-152 clamp and *255.999, 141-143 int()."""
    assert orc.encode_pixel((0.25, 1.0, 0.0)).tolist() == [127, 255, 0]
    assert orc.encode_pixel((0.7, 0.8, 1.0)).tolist() == [214, 228, 255]
    assert orc.encode_pixel((4.0, 1e-12, 0.5)).tolist() == [255, 0, 181]


def test_image_texture_quirks(orc):
    """materials.go:175-193: v flipped, nearest texel, NO clamp to W-1/H-1 -> out-of-bounds colour."""
    img = np.zeros((2, 4, 3), np.uint16)
    img[0, 0] = (65535, 0, 0)       # top-left texel
    img[1, 3] = (0, 0, 65535)       # bottom-right texel
    tex = np.zeros(1, scenes.TEXTURE_DT)
    tex[0]["kind"], tex[0]["image"], tex[0]["oob"] = abi.RT_TEX_IMAGE, 0, (0.25, 0.5, 0.75)
    s = scenes.SceneData(np.zeros(0, scenes.SPHERE_DT), np.zeros(0, scenes.MATERIAL_DT), tex, images=[img])
    p = (0, 0, 0)
    assert orc.texture(s, 0, 0.0, 0.99, p).tolist() == [1, 0, 0]     # v=0.99 -> j = int(0.01*2) = 0
    assert orc.texture(s, 0, 0.99, 0.25, p).tolist() == [0, 0, 1]    # j = int(0.75*2) = 1, i = 3
    assert orc.texture(s, 0, 1.0, 0.5, p).tolist() == [0.25, 0.5, 0.75]   # i = W: out of bounds
    assert orc.texture(s, 0, 1.2, 0.5, p).tolist() == [0.25, 0.5, 0.75]   # u clamps to 1 first
    assert orc.texture(s, 0, 0.5, 0.0, p).tolist() == [0.25, 0.5, 0.75]   # v = 0 -> j = H


def test_camera_init(orc):
    """camera.go:128-166 on the random-scene camera (main.go:228-239)."""
    for width, height in [(400, 225), (1200, 675), (1920, 1080), (3840, 2160)]:
        cam = orc.camera_from_options(scenes.camera_options(width, 10))
        assert (cam.width, cam.height) == (width, height)
    cam = orc.camera_from_options(scenes.camera_options(400, 10))
    assert list(cam.center) == [13, 2, 3]
    # the pixel grid is centred on the look-at point at the focus distance (10 from the eye)
    du, dv, p00 = (np.array(list(x), np.float64) for x in (cam.pixel_du, cam.pixel_dv, cam.pixel00))
    centre = p00 + du * 199.5 + dv * 112
    eye = np.array([13, 2, 3.0])
    w = eye / np.linalg.norm(eye)
    np.testing.assert_allclose(centre, eye - 10 * w, atol=2e-4)
    assert abs(np.dot(du, dv)) < 1e-9 and abs(np.dot(du, w)) < 1e-7
    # viewport height = 2 tan(10 deg) * 10 over 225 rows
    assert np.linalg.norm(dv) == pytest.approx(2 * math.tan(math.radians(10)) * 10 / 225, rel=1e-5)
    # defocus disk radius = 10 tan(0.3 deg)
    assert np.linalg.norm(list(cam.defocus_u)) == pytest.approx(10 * math.tan(math.radians(0.3)), rel=1e-5)


def test_get_ray_draw_order(orc):
    """camera.go:265-299 with the block layout of rt_rng.h: block 0 = (dx, dy, disk.x, disk.y), a
    further block (two candidate pairs) per rejected disk pair; direction = pixel sample - origin."""
    cam = orc.camera_from_options(scenes.camera_options(400, 1))
    seed = 99
    c, du, dv = (np.array(list(x), F) for x in (cam.center, cam.pixel_du, cam.pixel_dv))
    p00, ku, kv = (np.array(list(x), F) for x in (cam.pixel00, cam.defocus_u, cam.defocus_v))
    extra_blocks = 0
    for pix, k in [(400 * 100 + 37, 3), (5, 0), (400 * 224 + 399, 7), (12345, 11), (777, 2), (31337, 5)]:
        o, d = orc.primary_rays(cam, seed, pix, 1, k, 1)
        f = orc.rng_floats(seed, pix, k, 40)
        dx, dy = F(-0.5) + f[0], F(-0.5) + f[1]
        q = 2
        while True:
            sx, sy = F(-1) + f[q] * F(2), F(-1) + f[q + 1] * F(2)
            q += 2
            if sx * sx + sy * sy < 1:
                break
        extra_blocks += (q - 1) // 4
        i, j = pix % 400, pix // 400
        pc = p00 + du * F(i)
        pc = pc + dv * F(j)
        pc = pc + (du * dx + dv * dy)
        origin = c + (ku * sx + kv * sy)
        assert o[0].tolist() == origin.tolist()
        assert d[0].tolist() == (pc - origin).tolist()
    assert extra_blocks >= 1  # at least one of the cases exercised the rejection path


def test_materials_scatter(orc):
    """One Scatter per material kind on a unit sphere hit at (0,0,1) by a ray along -z."""
    def scene_with(kind, **kw):
        tex = np.zeros(1, scenes.TEXTURE_DT)
        tex[0]["a"] = (0.1, 0.2, 0.3)
        mat = np.zeros(1, scenes.MATERIAL_DT)
        mat[0]["kind"] = kind
        for k, v in kw.items():
            mat[0][k] = v
        sph = np.zeros(1, scenes.SPHERE_DT)
        sph[0] = (0, 0, 0, 1, 0)
        return scenes.SceneData(sph, mat, tex)

    o, d = (0, 0, 3), (0, 0, -2)
    lam = orc.scatter(scene_with(abi.RT_MAT_LAMBERTIAN), o, d, seed=5)
    assert lam["scattered"] and lam["origin"].tolist() == [0, 0, 1]
    assert lam["attenuation"].tolist() == [F(0.1), F(0.2), F(0.3)]
    # direction = normal + unit vector: within the unit ball around the normal tip
    assert np.linalg.norm(lam["dir"] - np.array([0, 0, 1])) == pytest.approx(1.0, abs=1e-6)

    met = orc.scatter(scene_with(abi.RT_MAT_METAL, albedo=(0.7, 0.6, 0.5), fuzz=0.0), o, d, seed=5)
    assert met["scattered"] and met["dir"].tolist() == [0, 0, 1]   # mirror of unit(-z) about +z
    assert met["attenuation"].tolist() == [F(0.7), F(0.6), F(0.5)]

    die = orc.scatter(scene_with(abi.RT_MAT_DIELECTRIC, ior=1.5), o, d, seed=5)
    assert die["scattered"] and die["attenuation"].tolist() == [1, 1, 1]
    # normal incidence: reflectance 0.04, so the stream's first float decides; both are along z
    assert abs(die["dir"][2]) == pytest.approx(1.0, abs=1e-6) and die["dir"][0] == 0 and die["dir"][1] == 0

    light = orc.scatter(scene_with(abi.RT_MAT_DIFFUSE_LIGHT), o, d, seed=5)
    assert not light["scattered"]


def test_empty_world_render(orc):
    """ray.go:53: no hittables -> every pixel is the encoded background."""
    s = scenes.SceneData(np.zeros(0, scenes.SPHERE_DT), np.zeros(0, scenes.MATERIAL_DT),
                         np.zeros(0, scenes.TEXTURE_DT))
    cam = orc.camera_from_options(scenes.camera_options(32, 2))
    for mode in (orc.MODE_LINEAR, orc.MODE_REF_BVH):
        rgb, acc, st = orc.render(s, cam, 1, mode=mode)
        assert (rgb == np.array([214, 228, 255], np.uint8)).all()
        assert st.rays == st.samples == 32 * 18 * 2


def test_depth_limit(orc):
    """ray.go:33-35: depth 0 is black; depth 1 on a closed diffuse scene is black where it hits."""
    s = scenes.random_scene()
    o = scenes.camera_options(48, 1, max_depth=0)
    cam = orc.camera_from_options(o)
    rgb, acc, st = orc.render(s, cam, 1)
    assert (acc == 0).all() and st.rays == 0
    cam1 = orc.camera_from_options(scenes.camera_options(48, 1, max_depth=1))
    for order in (orc.ORDER_RECURSIVE, orc.ORDER_ITERATIVE):
        rgb, acc, st = orc.render(s, cam1, 1, order=order)
        ro, rd = orc.primary_rays(cam1, 1, 0, 48 * 27, 0, 1)
        ids, _ = orc.trace(s, ro, rd)
        hit = (ids >= 0).reshape(27, 48)
        assert (acc[hit] == 0).all() and (acc[~hit] == np.array([0.7, 0.8, 1.0], F)).all()


def test_recursive_and_iterative_orders_agree(orc):
    """The device accumulates front to back; the reference recurses (ray.go:48-50).  Same paths,
    products associated differently: equal to a few ulp."""
    s = scenes.random_scene()
    cam = orc.camera_from_options(scenes.camera_options(120, 4))
    r0, a0, s0 = orc.render(s, cam, 11, order=orc.ORDER_RECURSIVE)
    r1, a1, s1 = orc.render(s, cam, 11, order=orc.ORDER_ITERATIVE)
    assert s0.rays == s1.rays and s0.hits == s1.hits
    np.testing.assert_allclose(a0, a1, rtol=2e-6, atol=1e-7)
    assert (np.abs(r0.astype(int) - r1.astype(int)) <= 1).all()


def test_reference_bvh_equals_linear_list(orc):
    """BVH.Hit over a reference-style tree (bvh.go:142-249) == World.Hit on primary rays, for
    several axis-choice seeds (the reference's tree is random per run, bvh.go:147)."""
    s = scenes.random_scene()
    cam = orc.camera_from_options(scenes.camera_options(200, 1))
    ro, rd = orc.primary_rays(cam, 3, 0, cam.width * cam.height, 0, 1)
    ids, ts = orc.trace(s, ro, rd, mode=orc.MODE_LINEAR)
    for bvh_seed in (1, 2, 3):
        i2, t2 = orc.trace(s, ro, rd, mode=orc.MODE_REF_BVH, bvh_seed=bvh_seed)
        assert (ids != i2).mean() < 1e-4
        same = ids == i2
        assert np.array_equal(ts[same], t2[same])


def test_quad_hit(orc):
    """hittables.go:149-190: Q=(-1,-1,-2), u=(2,0,0), v=(0,2,0); o=0, d=(0,0,-1) -> t=2, alpha=beta=0.5."""
    tex, mat = np.zeros(1, scenes.TEXTURE_DT), np.zeros(1, scenes.MATERIAL_DT)
    q = np.zeros(1, scenes.QUAD_DT)
    q[0] = ((-1, -1, -2), (2, 0, 0), (0, 2, 0), 0)
    s = scenes.SceneData(np.zeros(0, scenes.SPHERE_DT), mat, tex, quads=q)
    h = orc.hit_info(s, (0, 0, 0), (0, 0, -1))
    assert h["id"] == 0 and h["t"] == 2.0 and h["u"] == 0.5 and h["v"] == 0.5
    assert h["normal"].tolist() == [0, 0, 1] and h["front_face"]      # n = u x v = +z, ray along -z
    assert orc.hit_info(s, (0, 0, 0), (1.01, 0, -2)) is None           # alpha > 1: outside (InPlane)
    assert orc.hit_info(s, (0, 0, 0), (1.0, 1.0, -2))["u"] == 1.0      # the edge is inside
    assert orc.hit_info(s, (0, 0, 0), (1, 0, 0)) is None               # parallel: |denom| < 1e-8
    back = orc.hit_info(s, (0, 0, -4), (0, 0, 1))
    assert back["t"] == 2.0 and not back["front_face"] and back["normal"].tolist() == [0, 0, -1]


def test_mixed_world_object_ids(orc):
    """Object ID = position in World.hittables across spheres and quads; first object wins ties."""
    m = scenes.mixed_scene()
    o = [(100, 300, 400), (190, 400, 190), (278, 278, -800), (278, 100, 278)]
    d = [(0, -1, 0), (0, -1, 0), (0, 0, 1), (0, 1, 0)]
    ids, ts = orc.trace(m, o, d)
    assert ids.tolist() == [0, 1, 3, 2]        # floor quad, glass sphere from above, back wall, light
    assert ts.tolist() == [300.0, 220.0, 1355.0, 454.0]
    i2, _ = orc.trace(m, o, d, mode=orc.MODE_REF_BVH)
    assert np.array_equal(ids, i2)


def test_cornell_box_renders(orc):
    """main.go:194-225: emissive quad lights a closed box; background is black."""
    s = scenes.cornell_box_scene()
    cam = orc.camera_from_options(scenes.cornell_camera_options(48, 32))
    assert (cam.width, cam.height) == (48, 48)
    rgb, acc, st = orc.render(s, cam, 3)
    mean = acc / 32
    assert mean.max() == 15.0                                   # the light seen directly (15,15,15)
    assert 0.3 < (acc > 0).any(-1).mean() < 0.98                # noisy but lit; black background never adds
    left, right = mean[:, :8].reshape(-1, 3).mean(0), mean[:, -8:].reshape(-1, 3).mean(0)
    assert left[1] > left[0] and right[0] > right[1]            # green wall at image-left, red at image-right
    r2, a2, _ = orc.render(s, cam, 3, mode=orc.MODE_REF_BVH)
    assert (np.abs(a2 - acc) > 1e-3).any(-1).mean() < 0.01


def test_perlin_noise_kat(orc):
    """materials.go:223-288: at lattice points every corner weight's offset vector is 0, so Noise and
    Turb vanish and the marble is 0.5*(1 + sin(scale*z)); off-lattice the value stays in [0, 1]."""
    s = scenes.perlin_demo_scene()
    v = orc.texture(s, 0, 0, 0, (1, 2, 0.75))            # scale 4 -> (4, 8, 3): all octaves on the lattice
    want = np.float32(0.5) * (np.float32(1) + np.float32(math.sin(3.0)))
    assert v.tolist() == [float(want)] * 3
    rng = np.random.default_rng(0)
    vals = np.array([orc.texture(s, 0, 0, 0, p)[0] for p in rng.uniform(-5, 5, size=(300, 3))])
    assert vals.min() >= 0 and vals.max() <= 1 and vals.std() > 0.2
    # negative coordinates index the tables through `& 255` (Go ints are two's complement)
    assert 0 <= orc.texture(s, 0, 0, 0, (-0.3, -7.2, -100.6))[0] <= 1
    # a different table seed gives a different field
    s2 = scenes.perlin_demo_scene(seed=99)
    assert orc.texture(s2, 0, 0, 0, (0.3, 0.4, 0.5))[0] != orc.texture(s, 0, 0, 0, (0.3, 0.4, 0.5))[0]


def test_host_resolve_equals_oracle_resolve(orc):
    """api.resolve_host (numpy; used by the checkpointed render of the Python mirror) against the oracle's
    restatement of camera.go:261 + vec3.go:141-166 on sums that cover negatives, values above 1, huge values,
    zeros, NaN and the pixel-encode KATs of SURVEY section 4."""
    from raytracer_go_b200 import api
    rng = np.random.default_rng(12)
    spp = 37
    acc = (rng.uniform(-0.2, 1.4, size=(64, 97, 3)) * spp).astype(np.float32)
    acc[0, 0] = (0.25 * spp, 1.0 * spp, 0.0)
    acc[0, 1] = (0.7 * spp, 0.8 * spp, 1.0 * spp)
    acc[0, 2] = (np.nan, np.inf, -np.inf)
    acc[0, 3] = (1e30, 1e-30, -0.0)
    assert np.array_equal(api.resolve_host(acc, spp), orc.resolve(acc, spp))
