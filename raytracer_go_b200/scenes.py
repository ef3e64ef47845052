"""Deterministic synthetic scenes following the reference's own recipes (main.go:227-289, 80-104).

The reference seeds its scene RNG from the clock (main.go:246), so its scenes differ run to run;
these builders draw the same quantities in the same order from a counter-based Philox stream per
grid cell, with the float32 arithmetic of the Go expressions, so that the oracle, the CUDA
library and (elsewhere) the Go side can all be handed identical bytes.

Scene data are numpy structured arrays laid out exactly like the C structs of include/rt_b200.h.
"""
import ctypes as C
import hashlib
import math

import numpy as np

from . import abi
from .philox import stream_floats

F = np.float32

SPHERE_DT = np.dtype([("cx", F), ("cy", F), ("cz", F), ("r", F), ("material", np.uint32)])
MATERIAL_DT = np.dtype([("kind", np.uint32), ("albedo", F, 3), ("fuzz", F), ("ior", F),
                        ("texture", np.uint32)])
QUAD_DT = np.dtype([("q", F, 3), ("u", F, 3), ("v", F, 3), ("material", np.uint32)])
TEXTURE_DT = np.dtype([("kind", np.uint32), ("a", F, 3), ("b", F, 3), ("scale", F),
                       ("image", np.uint32), ("oob", F, 3)])
assert SPHERE_DT.itemsize == C.sizeof(abi.rt_sphere)
assert MATERIAL_DT.itemsize == C.sizeof(abi.rt_material)
assert TEXTURE_DT.itemsize == C.sizeof(abi.rt_texture)
PERLIN_DT = np.dtype([("vec", F, (256, 3)), ("perm_x", np.uint8, 256), ("perm_y", np.uint8, 256),
                      ("perm_z", np.uint8, 256)])
assert QUAD_DT.itemsize == C.sizeof(abi.rt_quad)
assert PERLIN_DT.itemsize == C.sizeof(abi.rt_perlin)

SCENE_SEED_RANDOM = 0x5EED0001
TEXTURE_SEED_EARTH = 0x5EED0003
SCENE_SEED_STRESS = 0x5EED0004
RENDER_SEED = 0xC0FFEE


class SceneData:
    """Flat, pointer-free scene (what the Go bridge would fill by walking World.hittables)."""

    def __init__(self, spheres, materials, textures, images=(), ray_origin_radius=0.0, name="", quads=None,
                 sphere_ids=None, quad_ids=None, perlins=None):
        self.spheres = np.ascontiguousarray(spheres, dtype=SPHERE_DT)
        self.quads = np.ascontiguousarray(quads if quads is not None else np.zeros(0, QUAD_DT), dtype=QUAD_DT)
        # object IDs = position in World.hittables (hittables.go:48-53); None = spheres first, then quads
        self.sphere_ids = None if sphere_ids is None else np.ascontiguousarray(sphere_ids, np.uint32)
        self.quad_ids = None if quad_ids is None else np.ascontiguousarray(quad_ids, np.uint32)
        self.perlins = np.ascontiguousarray(perlins if perlins is not None else np.zeros(0, PERLIN_DT), dtype=PERLIN_DT)
        self.materials = np.ascontiguousarray(materials, dtype=MATERIAL_DT)
        self.textures = np.ascontiguousarray(textures, dtype=TEXTURE_DT)
        self.images = [np.ascontiguousarray(im, dtype=np.uint16) for im in images]  # (h, w, 3)
        self.ray_origin_radius = float(ray_origin_radius)
        self.name = name

    def to_desc(self):
        """-> (rt_scene_desc, keepalive).  Pointers alias this object's arrays."""
        d = abi.rt_scene_desc()
        d.abi_version = abi.RT_B200_ABI_VERSION
        d.spheres = self.spheres.ctypes.data_as(C.POINTER(abi.rt_sphere))
        d.n_spheres = len(self.spheres)
        d.materials = self.materials.ctypes.data_as(C.POINTER(abi.rt_material))
        d.n_materials = len(self.materials)
        d.textures = self.textures.ctypes.data_as(C.POINTER(abi.rt_texture))
        d.n_textures = len(self.textures)
        imgs = (abi.rt_image * max(1, len(self.images)))()
        for k, im in enumerate(self.images):
            imgs[k].h, imgs[k].w = im.shape[0], im.shape[1]
            imgs[k].rgb16 = im.ctypes.data_as(C.POINTER(C.c_uint16))
        d.images = imgs
        d.n_images = len(self.images)
        d.ray_origin_radius = self.ray_origin_radius
        d.perlins = self.perlins.ctypes.data_as(C.POINTER(abi.rt_perlin))
        d.n_perlins = len(self.perlins)
        d.quads = self.quads.ctypes.data_as(C.POINTER(abi.rt_quad))
        d.n_quads = len(self.quads)
        if self.sphere_ids is not None or self.quad_ids is not None:
            sid = self.sphere_ids if self.sphere_ids is not None else np.zeros(0, np.uint32)
            qid = self.quad_ids if self.quad_ids is not None else np.zeros(0, np.uint32)
            d.sphere_ids = sid.ctypes.data_as(C.POINTER(C.c_uint32))
            d.quad_ids = qid.ctypes.data_as(C.POINTER(C.c_uint32))
            return d, (imgs, self, sid, qid)
        return d, (imgs, self)

    def n_objects(self):
        return len(self.spheres) + len(self.quads)

    def nbytes(self):
        return (self.spheres.nbytes + self.quads.nbytes + self.materials.nbytes + self.textures.nbytes + self.perlins.nbytes
                + sum(im.nbytes for im in self.images))

    def sha256(self):
        h = hashlib.sha256()
        for a in (self.spheres, self.quads, self.materials, self.textures, self.perlins, *self.images):
            h.update(a.tobytes())
        return h.hexdigest()


def _tex(kind=abi.RT_TEX_SOLID, a=(0, 0, 0), b=(0, 0, 0), scale=0.0, image=0, oob=(0, 0, 0)):
    t = np.zeros((), TEXTURE_DT)
    t["kind"], t["a"], t["b"], t["scale"], t["image"], t["oob"] = kind, a, b, scale, image, oob
    return t


def _mat(kind, albedo=(0, 0, 0), fuzz=0.0, ior=0.0, texture=0):
    m = np.zeros((), MATERIAL_DT)
    m["kind"], m["albedo"], m["fuzz"], m["ior"], m["texture"] = kind, albedo, fuzz, ior, texture
    return m


def _sph(c, r, material):
    s = np.zeros((), SPHERE_DT)
    s["cx"], s["cy"], s["cz"], s["r"], s["material"] = c[0], c[1], c[2], r, material
    return s


def _quad(q, u, v, material):
    r = np.zeros((), QUAD_DT)
    r["q"], r["u"], r["v"], r["material"] = q, u, v, material
    return r


def box_quads(a, b, material):
    """hittables.go:200-216 `Box(a, b, mat)`: the six quads, in the reference's order."""
    a, b = np.asarray(a, F), np.asarray(b, F)
    mn, mx = np.minimum(a, b), np.maximum(a, b)
    dx, dy, dz = (F(mx[0] - mn[0]), 0, 0), (0, F(mx[1] - mn[1]), 0), (0, 0, F(mx[2] - mn[2]))
    neg = lambda v: tuple(F(-1) * F(c) for c in v)  # noqa: E731  (Scale(v, -1))
    return [
        _quad((mn[0], mn[1], mx[2]), dx, dy, material),
        _quad((mx[0], mn[1], mx[2]), neg(dz), dy, material),
        _quad((mx[0], mn[1], mn[2]), neg(dx), dy, material),
        _quad((mn[0], mn[1], mn[2]), dz, dy, material),
        _quad((mn[0], mx[1], mx[2]), dx, neg(dz), material),
        _quad((mn[0], mn[1], mn[2]), dx, dz, material),
    ]


def new_perlin(seed):
    """NewPerlin (materials.go:202-216): 256 vectors NewVec3RandRange32(-1, 1) and three permutations
    made by Permute (materials.go:272-278: `target := rand.Intn(i)`, i from 255 down to 1).  The
    reference draws them from clock-seeded generators; here from a seeded numpy Philox."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    p = np.zeros((), PERLIN_DT)
    r = rng.random((256, 3), dtype=np.float32)
    p["vec"] = F(-1) + r * F(2)
    for name in ("perm_x", "perm_y", "perm_z"):
        a = np.arange(256, dtype=np.uint8)
        for i in range(255, 0, -1):
            t = int(rng.integers(0, i))
            a[i], a[t] = a[t], a[i]
        p[name] = a
    return p


def perlin_demo_scene(seed=0x5EED0002):
    """main.go:106-130: two spheres with the marble NoiseTexture(scale 4)."""
    tex = [_tex(abi.RT_TEX_NOISE, scale=4.0, image=0)]
    mats = [_mat(abi.RT_MAT_LAMBERTIAN, texture=0)]
    sph = [_sph((0, -1000, 0), 1000, 0), _sph((0, 2, 0), 2, 0)]
    return SceneData(np.array(sph, SPHERE_DT), np.array(mats, MATERIAL_DT), np.array(tex, TEXTURE_DT),
                     perlins=np.array([new_perlin(seed)], PERLIN_DT), name="perlin")


def simple_light_scene(seed=0x5EED0002):
    """main.go:162-192: marble ground and sphere, a red sphere, and a spherical DiffuseLight(4,4,4)."""
    tex = [_tex(abi.RT_TEX_NOISE, scale=4.0, image=0), _tex(a=(1, 0, 0)), _tex(a=(4, 4, 4))]
    mats = [_mat(abi.RT_MAT_LAMBERTIAN, texture=0), _mat(abi.RT_MAT_LAMBERTIAN, texture=1),
            _mat(abi.RT_MAT_DIFFUSE_LIGHT, texture=2)]
    sph = [_sph((0, -1000, 0), 1000, 0), _sph((0, 2, 0), 2, 0), _sph((-4, 2, 4), 2, 1), _sph((0, 7, 0), 2, 2)]
    return SceneData(np.array(sph, SPHERE_DT), np.array(mats, MATERIAL_DT), np.array(tex, TEXTURE_DT),
                     perlins=np.array([new_perlin(seed)], PERLIN_DT), name="simple-light")


def simple_light_camera_options(width=400, spp=500):
    """main.go:163-173."""
    return camera_options(width, spp, look_from=(26, 3, 6), look_at=(0, 2, 0), vfov_deg=20.0, defocus_deg=0.0,
                          background=(0, 0, 0))


def perlin_camera_options(width=400, spp=100):
    """main.go:107-117."""
    return camera_options(width, spp, look_from=(13, 2, 3), look_at=(0, 0, 0), vfov_deg=20.0, defocus_deg=0.0)


def cornell_box_scene():
    """main.go:194-225 — the scene main.go renders as checked in (main.go:55)."""
    tex = [_tex(a=(.65, .05, .05)), _tex(a=(.73, .73, .73)), _tex(a=(.12, .45, .15)), _tex(a=(15, 15, 15))]
    red, white, green, light = 0, 1, 2, 3
    mats = [_mat(abi.RT_MAT_LAMBERTIAN, texture=0), _mat(abi.RT_MAT_LAMBERTIAN, texture=1),
            _mat(abi.RT_MAT_LAMBERTIAN, texture=2), _mat(abi.RT_MAT_DIFFUSE_LIGHT, texture=3)]
    quads = [
        _quad((555, 0, 0), (0, 555, 0), (0, 0, 555), green),
        _quad((0, 0, 0), (0, 555, 0), (0, 0, 555), red),
        _quad((343, 554, 332), (-130, 0, 0), (0, 0, -105), light),
        _quad((0, 0, 0), (555, 0, 0), (0, 0, 555), white),
        _quad((555, 555, 555), (-555, 0, 0), (0, 0, -555), white),
        _quad((0, 0, 555), (555, 0, 0), (0, 555, 0), white),
    ]
    quads += box_quads((130, 0, 65), (295, 165, 230), white)
    quads += box_quads((265, 0, 295), (430, 330, 460), white)
    return SceneData(np.zeros(0, SPHERE_DT), np.array(mats, MATERIAL_DT), np.array(tex, TEXTURE_DT),
                     quads=np.array(quads, QUAD_DT), name="cornell")


def cornell_camera_options(width=600, spp=200, max_depth=50):
    """main.go:195-205."""
    return camera_options(width, spp, max_depth, look_from=(278, 278, -800), look_at=(278, 278, 0), vfov_deg=40.0,
                          defocus_deg=0.0, background=(0, 0, 0), aspect=1.0)


def quad_demo_scene():
    """main.go:132-160: five coloured quads."""
    cols = [(1, .2, .2), (.2, 1, .2), (.2, .2, 1), (1, .5, 0), (.2, .8, .8)]
    tex = [_tex(a=c) for c in cols]
    mats = [_mat(abi.RT_MAT_LAMBERTIAN, texture=i) for i in range(5)]
    quads = [
        _quad((-3, -2, 5), (0, 0, -4), (0, 4, 0), 0),
        _quad((-2, -2, 0), (4, 0, 0), (0, 4, 0), 1),
        _quad((3, -2, 1), (0, 0, 4), (0, 4, 0), 2),
        _quad((-2, 3, 1), (4, 0, 0), (0, 0, 4), 3),
        _quad((-2, -3, 5), (4, 0, 0), (0, 0, -4), 4),
    ]
    return SceneData(np.zeros(0, SPHERE_DT), np.array(mats, MATERIAL_DT), np.array(tex, TEXTURE_DT),
                     quads=np.array(quads, QUAD_DT), name="quads")


def quad_demo_camera_options(width=400, spp=100):
    """main.go:133-143."""
    return camera_options(width, spp, look_from=(0, 0, 9), look_at=(0, 0, 0), vfov_deg=80.0, defocus_deg=0.0)


def mixed_scene():
    """Spheres and quads interleaved in World order (exercises object IDs across both kinds): a lit
    room of quads with a glass, a metal and a diffuse sphere in it."""
    tex = [_tex(a=(.73, .73, .73)), _tex(a=(7, 7, 7)), _tex(abi.RT_TEX_CHECKER, a=(.2, .3, .1), b=(.9, .9, .9), scale=40.0),
           _tex(a=(.65, .05, .05))]
    mats = [_mat(abi.RT_MAT_LAMBERTIAN, texture=0), _mat(abi.RT_MAT_DIFFUSE_LIGHT, texture=1),
            _mat(abi.RT_MAT_LAMBERTIAN, texture=2), _mat(abi.RT_MAT_DIELECTRIC, ior=1.5),
            _mat(abi.RT_MAT_METAL, albedo=(.8, .85, .88), fuzz=0.05), _mat(abi.RT_MAT_LAMBERTIAN, texture=3)]
    quads = [
        _quad((0, 0, 0), (555, 0, 0), (0, 0, 555), 2),        # floor (checker)        id 0
        _quad((213, 554, 227), (130, 0, 0), (0, 0, 105), 1),  # light                  id 2
        _quad((0, 0, 555), (555, 0, 0), (0, 555, 0), 0),      # back wall              id 3
        _quad((0, 555, 0), (555, 0, 0), (0, 0, 555), 0),      # ceiling                id 5
    ]
    spheres = [_sph((190, 90, 190), 90, 3), _sph((400, 120, 300), 120, 4), _sph((300, 60, 120), 60, 5)]  # ids 1, 4, 6
    return SceneData(np.array(spheres, SPHERE_DT), np.array(mats, MATERIAL_DT), np.array(tex, TEXTURE_DT),
                     quads=np.array(quads, QUAD_DT), sphere_ids=[1, 4, 6], quad_ids=[0, 2, 3, 5], name="mixed")


def random_scene(half=11, seed=SCENE_SEED_RANDOM, exclude=(), name="random"):
    """main.go:240-286 on cells i,j in [-half, half).  `exclude`: extra (x, z, radius) discs in
    which small spheres are dropped (used to make room for the earth sphere of config C3)."""
    n = 2 * half
    ii, jj = np.meshgrid(np.arange(-half, half, dtype=np.int64), np.arange(-half, half, dtype=np.int64),
                         indexing="ij")  # i outer, j inner: main.go:249-250
    ii, jj = ii.ravel(), jj.ravel()
    cell = np.arange(n * n, dtype=np.uint64)
    # draw order per cell: matPer, cx, cz, then the material's draws (main.go:251-266)
    r = stream_floats(seed, (cell & 0xFFFFFFFF).astype(np.uint32), (cell >> 32).astype(np.uint32), 9,
                      tag=0x5CE0E)
    mat_per = r[:, 0]
    cx = ii.astype(F) + F(0.9) * r[:, 1]
    cy = np.full(cx.shape, F(0.2))
    cz = jj.astype(F) + F(0.9) * r[:, 2]
    # main.go:254-256: keep if |centre - (4, 0.2, 0)| > 0.9
    dx, dy, dz = cx - F(4), cy - F(0.2), cz - F(0)
    ln = np.sqrt((dx * dx + dy * dy + dz * dz).astype(np.float64)).astype(F)
    keep = ln > F(0.9)
    for (ex, ez, er) in exclude:
        ddx, ddz = cx - F(ex), cz - F(ez)
        keep &= (ddx * ddx + ddz * ddz) > F(er) * F(er)

    textures = [_tex(abi.RT_TEX_CHECKER, a=(0.2, 0.3, 0.1), b=(0.9, 0.9, 0.9), scale=0.32)]  # main.go:242
    materials = [_mat(abi.RT_MAT_LAMBERTIAN, texture=0)]                                      # main.go:243
    spheres = [_sph((0, -1000, 0), 1000, 0)]                                                  # main.go:244

    idx = np.nonzero(keep)[0]
    k = len(idx)
    mp = mat_per[idx]
    is_lam = mp < F(0.8)
    is_met = (~is_lam) & (mp < F(0.95))
    tex_arr = np.zeros(k, TEXTURE_DT)
    mat_arr = np.zeros(k, MATERIAL_DT)
    sph_arr = np.zeros(k, SPHERE_DT)
    # Lambertian: Mul(NewVec3Rand32, NewVec3Rand32), main.go:259
    lam_col = r[idx, 3:6] * r[idx, 6:9]
    # Metal: NewVec3RandRange32(0.5, 1) and RandF32N(0, 0.5), main.go:264-265 / math.go:30-32
    met_alb = F(0.5) + r[idx, 3:6] * (F(1) - F(0.5))
    met_fuzz = F(0) + r[idx, 6] * (F(0.5) - F(0))
    n_tex = 1
    tex_index = np.zeros(k, np.uint32)
    lam_rows = np.nonzero(is_lam)[0]
    tex_index[lam_rows] = n_tex + np.arange(len(lam_rows), dtype=np.uint32)
    tex_sel = np.zeros(len(lam_rows), TEXTURE_DT)
    tex_sel["kind"] = abi.RT_TEX_SOLID
    tex_sel["a"] = lam_col[lam_rows]
    mat_arr["kind"] = np.where(is_lam, abi.RT_MAT_LAMBERTIAN,
                               np.where(is_met, abi.RT_MAT_METAL, abi.RT_MAT_DIELECTRIC))
    mat_arr["texture"] = tex_index
    mat_arr["albedo"] = np.where(is_met[:, None], met_alb, F(0))
    mat_arr["fuzz"] = np.where(is_met, met_fuzz, F(0))
    mat_arr["ior"] = np.where(~is_lam & ~is_met, F(1.5), F(0))                               # main.go:269
    sph_arr["cx"], sph_arr["cy"], sph_arr["cz"] = cx[idx], cy[idx], cz[idx]
    sph_arr["r"] = F(0.2)                                                                      # main.go:272
    sph_arr["material"] = 1 + np.arange(k, dtype=np.uint32)
    del tex_arr

    textures = np.concatenate([np.array(textures, TEXTURE_DT), tex_sel])
    materials = np.concatenate([np.array(materials, MATERIAL_DT), mat_arr])
    spheres = np.concatenate([np.array(spheres, SPHERE_DT), sph_arr])

    # main.go:278-285: the three big spheres
    t_brown = len(textures)
    textures = np.concatenate([textures, np.array([_tex(abi.RT_TEX_SOLID, a=(0.4, 0.2, 0.1))], TEXTURE_DT)])
    m0 = len(materials)
    materials = np.concatenate([materials, np.array([
        _mat(abi.RT_MAT_DIELECTRIC, ior=1.5),
        _mat(abi.RT_MAT_LAMBERTIAN, texture=t_brown),
        _mat(abi.RT_MAT_METAL, albedo=(0.7, 0.6, 0.5), fuzz=0.0)], MATERIAL_DT)])
    spheres = np.concatenate([spheres, np.array([
        _sph((0, 1, 0), 1, m0), _sph((-4, 1, 0), 1, m0 + 1), _sph((4, 1, 0), 1, m0 + 2)], SPHERE_DT)])
    return SceneData(spheres, materials, textures, name=name)


def procedural_earth_map(w=2048, h=1024, seed=TEXTURE_SEED_EARTH):
    """Stand-in for textures/earthmap.jpg (missing from the checkout, .MISSING_LARGE_BLOBS:2): a
    seeded multi-octave value-noise land/sea map, RGB8 widened x257 to the ABI's RGB16."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    acc = np.zeros((h, w), np.float64)
    amp, tot = 1.0, 0.0
    for octave in range(6):
        gh, gw = 4 << octave, 8 << octave
        g = rng.random((gh, gw))
        g = np.concatenate([g, g[:, :1]], axis=1)   # wrap in longitude
        g = np.concatenate([g, g[-1:, :]], axis=0)
        y = np.linspace(0, gh, h, endpoint=False)
        x = np.linspace(0, gw, w, endpoint=False)
        y0, x0 = y.astype(int), x.astype(int)
        fy, fx = (y - y0)[:, None], (x - x0)[None, :]
        fy, fx = fy * fy * (3 - 2 * fy), fx * fx * (3 - 2 * fx)
        a = g[y0][:, x0] * (1 - fx) + g[y0][:, x0 + 1] * fx
        b = g[y0 + 1][:, x0] * (1 - fx) + g[y0 + 1][:, x0 + 1] * fx
        acc += amp * (a * (1 - fy) + b * fy)
        tot += amp
        amp *= 0.5
    v = acc / tot
    lat = np.abs(np.linspace(-1, 1, h))[:, None]
    land = v > 0.52
    ice = lat > 0.86 - 0.1 * v
    rgb = np.zeros((h, w, 3), np.float64)
    sea = np.stack([0.02 + 0.1 * v, 0.12 + 0.3 * v, 0.35 + 0.5 * v], -1)
    gnd = np.stack([0.15 + 0.6 * (v - 0.5), 0.35 + 0.3 * (1 - lat) * np.ones_like(v), 0.1 + 0.2 * v], -1)
    rgb[:] = sea
    rgb[land] = gnd[land]
    rgb[np.broadcast_to(ice, land.shape)] = 0.92
    rgb8 = np.clip(rgb * 255.0, 0, 255).astype(np.uint16)
    return (rgb8 * np.uint16(257)).astype(np.uint16)


# Go's zero colour for an out-of-bounds *image.YCbCr pixel, RGBA() -> (0, 34678, 0)
JPEG_OOB = (F(0) * F(1.0 / 65535.0), F(34678) * F(1.0 / 65535.0), F(0) * F(1.0 / 65535.0))
EARTH_CENTER = (-4.0, 2.0, -4.5)


def earth_random_scene(seed=SCENE_SEED_RANDOM, tex_w=2048, tex_h=1024):
    """Config C3: the random scene plus the earth sphere of main.go:98-100 (radius 2,
    Lambertian(ImageTexture)), moved to EARTH_CENTER with the small spheres under it dropped."""
    base = random_scene(11, seed, exclude=[(EARTH_CENTER[0], EARTH_CENTER[2], 1.6)], name="earth+random")
    img = procedural_earth_map(tex_w, tex_h)
    t = len(base.textures)
    textures = np.concatenate([base.textures, np.array(
        [_tex(abi.RT_TEX_IMAGE, image=0, oob=JPEG_OOB)], TEXTURE_DT)])
    m = len(base.materials)
    materials = np.concatenate([base.materials, np.array([_mat(abi.RT_MAT_LAMBERTIAN, texture=t)], MATERIAL_DT)])
    spheres = np.concatenate([base.spheres, np.array([_sph(EARTH_CENTER, 2, m)], SPHERE_DT)])
    return SceneData(spheres, materials, textures, images=[img], name="earth+random")


def earth_scene(tex_w=256, tex_h=128):
    """main.go:80-104: one textured sphere (small map for tests)."""
    img = procedural_earth_map(tex_w, tex_h)
    textures = np.array([_tex(abi.RT_TEX_IMAGE, image=0, oob=JPEG_OOB)], TEXTURE_DT)
    materials = np.array([_mat(abi.RT_MAT_LAMBERTIAN, texture=0)], MATERIAL_DT)
    spheres = np.array([_sph((0, 0, 0), 2, 0)], SPHERE_DT)
    return SceneData(spheres, materials, textures, images=[img], name="earth")


def stress_scene(half=500, seed=SCENE_SEED_STRESS):
    """Config C4: the same cell recipe on a (2*half)^2 grid (~1e6 spheres at half=500)."""
    s = random_scene(half, seed, name=f"stress{half}")
    s.ray_origin_radius = 150.0
    return s


def camera_options(width, spp, max_depth=50, look_from=(13, 2, 3), look_at=(0, 0, 0), vfov_deg=20.0,
                   defocus_deg=0.6, focus_dist=10.0, background=(0.7, 0.8, 1.0), aspect=16.0 / 9.0):
    """main.go:228-239 as an rt_camera_options (ToRadians = deg * float32(pi/180), math.go:46-52)."""
    o = abi.rt_camera_options()
    o.aspect_ratio = aspect
    o.image_width = int(width)
    o.spp = int(spp)
    o.max_depth = int(max_depth)
    rad_ratio = F(math.pi / 180.0)
    o.fov_radians = float(F(vfov_deg) * rad_ratio)
    o.defocus_angle_radians = float(F(defocus_deg) * rad_ratio)
    o.focus_dist = focus_dist
    o.look_from = (C.c_float * 3)(*look_from)
    o.look_at = (C.c_float * 3)(*look_at)
    o.vup = (C.c_float * 3)(0, 1, 0)
    o.background = (C.c_float * 3)(*background)
    return o


# BASELINE.json configs
CONFIGS = {
    "C1": dict(scene="random", width=1200, spp=10),
    "C2": dict(scene="random", width=1200, spp=500),
    "C3": dict(scene="earth+random", width=1920, spp=256),
    "C4": dict(scene="stress", width=1920, spp=64),
    "C5": dict(scene="random", width=3840, spp=4096),
    # not a BASELINE.json config: the scene main.go renders as checked in (main.go:55, 194-225)
    "CB": dict(scene="cornell", width=600, spp=200),
}


def build_config(name, width=None, spp=None, stress_half=500):
    cfg = dict(CONFIGS[name])
    if width:
        cfg["width"] = width
    if spp:
        cfg["spp"] = spp
    if cfg["scene"] == "random":
        scene = random_scene()
        cam = camera_options(cfg["width"], cfg["spp"])
    elif cfg["scene"] == "cornell":
        scene = cornell_box_scene()
        cam = cornell_camera_options(cfg["width"], cfg["spp"])
    elif cfg["scene"] == "earth+random":
        scene = earth_random_scene()
        cam = camera_options(cfg["width"], cfg["spp"])
    else:
        scene = stress_scene(stress_half)
        # raised and pulled back along the reference's view direction so a wide field is visible
        cam = camera_options(cfg["width"], cfg["spp"], look_from=(13 * 4, 2 * 12, 3 * 4))
    return scene, cam
