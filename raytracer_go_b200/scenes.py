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
TEXTURE_DT = np.dtype([("kind", np.uint32), ("a", F, 3), ("b", F, 3), ("scale", F),
                       ("image", np.uint32), ("oob", F, 3)])
assert SPHERE_DT.itemsize == C.sizeof(abi.rt_sphere)
assert MATERIAL_DT.itemsize == C.sizeof(abi.rt_material)
assert TEXTURE_DT.itemsize == C.sizeof(abi.rt_texture)

SCENE_SEED_RANDOM = 0x5EED0001
TEXTURE_SEED_EARTH = 0x5EED0003
SCENE_SEED_STRESS = 0x5EED0004
RENDER_SEED = 0xC0FFEE


class SceneData:
    """Flat, pointer-free scene (what the Go bridge would fill by walking World.hittables)."""

    def __init__(self, spheres, materials, textures, images=(), ray_origin_radius=0.0, name=""):
        self.spheres = np.ascontiguousarray(spheres, dtype=SPHERE_DT)
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
        return d, (imgs, self)

    def nbytes(self):
        return (self.spheres.nbytes + self.materials.nbytes + self.textures.nbytes
                + sum(im.nbytes for im in self.images))

    def sha256(self):
        h = hashlib.sha256()
        for a in (self.spheres, self.materials, self.textures, *self.images):
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
    elif cfg["scene"] == "earth+random":
        scene = earth_random_scene()
        cam = camera_options(cfg["width"], cfg["spp"])
    else:
        scene = stress_scene(stress_half)
        # raised and pulled back along the reference's view direction so a wide field is visible
        cam = camera_options(cfg["width"], cfg["spp"], look_from=(13 * 4, 2 * 12, 3 * 4))
    return scene, cam
