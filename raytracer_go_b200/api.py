"""Host-side mirror of the reference's scene-description API over librt_b200.so.

The names follow raytracer-go's package `internal` (camera.go, hittables.go, materials.go) so
that a scene written against the reference reads the same here:

    cam = NewCamera(16.0 / 9.0, 400, WithSamplesPerPixel(500), WithMaxRayDepth(50), ...)
    world = NewWorld(); world.Add(NewSphere(NewVec3(0, -1000, 0), 1000, NewLambertian(tex)))
    tree = NewBVHFromWorld(world)
    cam.Render(tree, f)            # writes the same "P3" PPM text (camera.go:180-252)

`Camera.Render` is the drop-in boundary: it flattens the world into the pointer-free arrays of
include/rt_b200.h (what the cgo bridge in INTEGRATION.md does on the Go side), makes ONE call into
the CUDA library and writes the PPM.  Nothing here computes colours on the CPU; without the
library and a B200 every compute call raises.
"""
import ctypes as C
import math

import numpy as np

from . import abi, lib, scenes

F = np.float32


# ------------------------------------------------------------------------------ low-level handle
class Scene:
    """Owns an rt_scene* (device BVH + materials + texels, uploaded once)."""

    def __init__(self, data, device=0):
        self._lib = lib.load()
        self.data = data
        desc, keep = data.to_desc()
        h = C.c_void_p()
        lib.check(self._lib.rt_scene_create(C.byref(desc), device, C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rt_scene_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream):
        """Run on the caller's stream (e.g. torch.cuda.current_stream().cuda_stream); 0 restores."""
        lib.check(self._lib.rt_scene_set_stream(self._h, C.c_void_p(cuda_stream)))

    def render(self, cam, seed=scenes.RENDER_SEED, sample_offset=0, sample_count=0, want_accum=False, flags=0,
               rows=None):
        """rt_render -> (rgb (H,W,3) uint8, accum (H,W,3) float32 or None, rt_stats).  `rows` =
        (row_begin, row_count, row_step) renders only those scanlines; the outputs then have row_count rows."""
        if rows and rows[1] == 0:  # in the C struct row_count 0 means "the whole image"
            raise ValueError("empty row set")
        H, W = (rows[1] if rows else cam.height), cam.width
        rgb = np.empty((H, W, 3), np.uint8)
        acc = np.empty((H, W, 3), np.float32) if want_accum else None
        opts = abi.rt_render_opts(seed, self.device, sample_offset, sample_count, flags, *(rows or (0, 0, 0)))
        st = abi.rt_stats()
        lib.check(self._lib.rt_render(self._h, C.byref(cam), C.byref(opts), rgb.ctypes.data_as(C.c_void_p),
                                      acc.ctypes.data_as(C.c_void_p) if want_accum else None, C.byref(st)))
        return rgb, acc, st

    def render_accum_device(self, cam, d_accum_ptr, seed=scenes.RENDER_SEED, sample_offset=0, sample_count=0,
                            flags=0, rows=None):
        """rt_render_accum_device into a device buffer of rows*W*3 float32 (e.g. tensor.data_ptr())."""
        if rows and rows[1] == 0:
            raise ValueError("empty row set")
        opts = abi.rt_render_opts(seed, self.device, sample_offset, sample_count, flags, *(rows or (0, 0, 0)))
        st = abi.rt_stats()
        lib.check(self._lib.rt_render_accum_device(self._h, C.byref(cam), C.byref(opts), C.c_void_p(d_accum_ptr),
                                                   C.byref(st)))
        return st

    def trace(self, origins, dirs, tmin=0.001, tmax=np.inf):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        ids = np.empty(len(o), np.int32)
        ts = np.empty(len(o), np.float32)
        lib.check(self._lib.rt_trace(self._h, o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), len(o),
                                     tmin, tmax, ids.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p)))
        return ids, ts

    def bvh_info(self):
        info = abi.rt_bvh_info()
        lib.check(self._lib.rt_scene_bvh_info(self._h, C.byref(info)))
        return info

    def bvh_copy(self):
        info = self.bvh_info()
        nodes = np.zeros((info.n_nodes, 8), np.uint32)
        ids = np.zeros(info.n_slots, np.int32)
        lib.check(self._lib.rt_scene_bvh_copy(self._h, nodes.ctypes.data_as(C.c_void_p),
                                              ids.ctypes.data_as(C.c_void_p)))
        return nodes, ids, info


def render_multi(data, cam, devices, seed=scenes.RENDER_SEED, sample_offset=0, sample_count=0, want_accum=False,
                 tile_split=False):
    """rt_render_multi: one call, several GPUs (sample-split inside the library, or interleaved
    scanlines with tile_split=True: bit-identical to the single-GPU render)."""
    desc, keep = data.to_desc()
    H, W = cam.height, cam.width
    rgb = np.empty((H, W, 3), np.uint8)
    acc = np.empty((H, W, 3), np.float32) if want_accum else None
    devs = (C.c_int32 * len(devices))(*devices)
    opts = abi.rt_render_opts(seed, devices[0], sample_offset, sample_count,
                              abi.RT_FLAG_TILE_SPLIT if tile_split else 0)
    st = abi.rt_stats()
    lib.check(lib.load().rt_render_multi(C.byref(desc), C.byref(cam), C.byref(opts), devs, len(devices),
                                         rgb.ctypes.data_as(C.c_void_p),
                                         acc.ctypes.data_as(C.c_void_p) if want_accum else None, C.byref(st)))
    return rgb, acc, st


def resolve_device(d_accum_ptr, width, height, total_spp, device=0, cuda_stream=0):
    rgb = np.empty((height, width, 3), np.uint8)
    lib.check(lib.load().rt_resolve_device(C.c_void_p(d_accum_ptr), width, height, total_spp, device,
                                           C.c_void_p(cuda_stream), rgb.ctypes.data_as(C.c_void_p)))
    return rgb


def primary_rays(cam, seed, pixel_begin, n_pixels, sample_offset=0, sample_count=1, device=0):
    n = n_pixels * sample_count
    o = np.empty((n, 3), np.float32)
    d = np.empty((n, 3), np.float32)
    opts = abi.rt_render_opts(seed, device, sample_offset, sample_count, 0)
    lib.check(lib.load().rt_primary_rays(C.byref(cam), C.byref(opts), pixel_begin, n_pixels,
                                         o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p)))
    return o, d


def camera_from_options(opts):
    cam = abi.rt_camera()
    lib.check(lib.load().rt_camera_from_options(C.byref(opts), C.byref(cam)))
    return cam


def device_count():
    return int(lib.load().rt_device_count())


# ------------------------------------------------------- mirror of package `internal` (reference)
def NewVec3(x, y, z):  # vec3.go:15
    return (F(x), F(y), F(z))


def NewVec3Zero():  # vec3.go:23
    return NewVec3(0, 0, 0)


class SolidColor:  # materials.go:151-163
    def __init__(self, albedo):
        self.albedo = albedo


def NewSolidColor(x, y, z):
    return SolidColor(NewVec3(x, y, z))


class Checkered:  # materials.go:121-145
    def __init__(self, scale, even, odd):
        self.scale, self.even, self.odd = F(scale), even, odd


def NewCheckered(scale, even, odd):
    return Checkered(scale, even, odd)


class ImageTexture:  # materials.go:165-173
    """img: (h, w, 3) uint16 array = img.At(i,j).RGBA() r,g,b; oob = the image model's zero colour."""

    def __init__(self, img, oob=scenes.JPEG_OOB):
        self.img, self.oob = np.ascontiguousarray(img, np.uint16), oob


def NewImageTexture(img, oob=scenes.JPEG_OOB):
    return ImageTexture(img, oob)


class NoiseTexture:  # materials.go:280-295
    """perlin: a scenes.PERLIN_DT record (NewPerlin's tables, materials.go:202-216)."""

    def __init__(self, perlin, scale):
        self.perlin, self.scale = perlin, F(scale)


def NewNoiseTexture(seed, scale):
    """The reference takes a *rand.Rand (clock-seeded, main.go:120-123); here a seed."""
    return NoiseTexture(scenes.new_perlin(seed), scale)


class Lambertian:  # materials.go:19-31
    def __init__(self, albedo):
        self.albedo = albedo


def NewLambertian(albedo):
    return Lambertian(albedo)


class Metal:  # materials.go:44-58
    def __init__(self, albedo, fuzz):
        self.albedo, self.fuzz = albedo, F(fuzz)


def NewMetal(albedo, fuzz):
    return Metal(albedo, fuzz)


class Dielectric:  # materials.go:77-89
    def __init__(self, ir):
        self.refractiveIndex = F(ir)


def NewDielectric(ir):
    return Dielectric(ir)


class DiffuseLight:  # materials.go:297-309
    def __init__(self, emit):
        self.emit = emit


def NewDiffuseLight(emit):
    return DiffuseLight(emit)


class Sphere:  # hittables.go:78-94
    def __init__(self, center, radius, mat):
        self.Center, self.Radius, self.Material = center, F(radius), mat


def NewSphere(center, radius, mat):
    return Sphere(center, radius, mat)


class Quad:  # hittables.go:138-165
    def __init__(self, Q, u, v, material):
        self.Q, self.u, self.v, self.material = Q, u, v, material


def NewQuad(Q, u, v, material):
    return Quad(Q, u, v, material)


def Box(a, b, mat):
    """hittables.go:200-216: the six quads of an axis-aligned box, in the reference's order."""
    mn = tuple(min(F(x), F(y)) for x, y in zip(a, b))
    mx = tuple(max(F(x), F(y)) for x, y in zip(a, b))
    dx, dy, dz = NewVec3(mx[0] - mn[0], 0, 0), NewVec3(0, mx[1] - mn[1], 0), NewVec3(0, 0, mx[2] - mn[2])
    neg = lambda v: tuple(F(-1) * c for c in v)  # noqa: E731  Scale(v, -1)
    return [
        NewQuad(NewVec3(mn[0], mn[1], mx[2]), dx, dy, mat),
        NewQuad(NewVec3(mx[0], mn[1], mx[2]), neg(dz), dy, mat),
        NewQuad(NewVec3(mx[0], mn[1], mn[2]), neg(dx), dy, mat),
        NewQuad(NewVec3(mn[0], mn[1], mn[2]), dz, dy, mat),
        NewQuad(NewVec3(mn[0], mx[1], mx[2]), dx, neg(dz), mat),
        NewQuad(NewVec3(mn[0], mn[1], mn[2]), dx, dz, mat),
    ]


class World:  # hittables.go:39-53
    def __init__(self):
        self.hittables = []

    def Add(self, *hittables):
        for h in hittables:  # world.Add(internal.Box(...)...) spreads a slice in Go (main.go:220)
            self.hittables.extend(h if isinstance(h, (list, tuple)) else [h])


def NewWorld():
    return World()


class BVH:
    """NewBVHFromWorld (bvh.go:138-140).  The reference builds a random-axis pointer tree here;
    the device BVH is built inside rt_scene_create, so this only keeps the insertion-ordered list
    (object ID = index in World.hittables, hittables.go:48-53)."""

    def __init__(self, world):
        self.hittables = list(world.hittables)


def NewBVHFromWorld(world):
    return BVH(world)


def flatten_world(world):
    """World/BVH -> scenes.SceneData: materials and textures de-duplicated by identity, spheres in
    insertion order.  This is the walk the cgo bridge performs."""
    tex_index, mat_index = {}, {}
    textures, materials, images, perlins = [], [], [], []
    spheres, quads, sphere_ids, quad_ids = [], [], [], []

    def tex_id(t):
        if id(t) in tex_index:
            return tex_index[id(t)]
        rec = np.zeros((), scenes.TEXTURE_DT)
        if isinstance(t, Checkered):
            rec["kind"], rec["a"], rec["b"], rec["scale"] = abi.RT_TEX_CHECKER, t.even, t.odd, t.scale
        elif isinstance(t, ImageTexture):
            rec["kind"], rec["image"], rec["oob"] = abi.RT_TEX_IMAGE, len(images), t.oob
            images.append(t.img)
        elif isinstance(t, SolidColor):
            rec["kind"], rec["a"] = abi.RT_TEX_SOLID, t.albedo
        elif isinstance(t, NoiseTexture):
            rec["kind"], rec["scale"], rec["image"] = abi.RT_TEX_NOISE, t.scale, len(perlins)
            perlins.append(t.perlin)
        else:
            raise TypeError(f"texture {type(t).__name__} is outside the accelerated path")
        textures.append(rec)
        tex_index[id(t)] = len(textures) - 1
        return tex_index[id(t)]

    def mat_id(m):
        if id(m) in mat_index:
            return mat_index[id(m)]
        rec = np.zeros((), scenes.MATERIAL_DT)
        if isinstance(m, Lambertian):
            rec["kind"], rec["texture"] = abi.RT_MAT_LAMBERTIAN, tex_id(m.albedo)
        elif isinstance(m, Metal):
            rec["kind"], rec["albedo"], rec["fuzz"] = abi.RT_MAT_METAL, m.albedo, m.fuzz
        elif isinstance(m, Dielectric):
            rec["kind"], rec["ior"] = abi.RT_MAT_DIELECTRIC, m.refractiveIndex
        elif isinstance(m, DiffuseLight):
            rec["kind"], rec["texture"] = abi.RT_MAT_DIFFUSE_LIGHT, tex_id(m.emit)
        else:
            raise TypeError(f"material {type(m).__name__} is outside the accelerated path")
        materials.append(rec)
        mat_index[id(m)] = len(materials) - 1
        return mat_index[id(m)]

    for k, h in enumerate(world.hittables):  # k = object ID (hittables.go:48-53)
        if isinstance(h, Sphere):
            spheres.append((h.Center[0], h.Center[1], h.Center[2], h.Radius, mat_id(h.Material)))
            sphere_ids.append(k)
        elif isinstance(h, Quad):
            quads.append((h.Q, h.u, h.v, mat_id(h.material)))
            quad_ids.append(k)
        else:
            raise TypeError(f"hittable {type(h).__name__} is not a Sphere or a Quad")
    return scenes.SceneData(np.array(spheres, scenes.SPHERE_DT).reshape(-1),
                            np.array(materials, scenes.MATERIAL_DT).reshape(-1),
                            np.array(textures, scenes.TEXTURE_DT).reshape(-1), images,
                            quads=np.array(quads, scenes.QUAD_DT).reshape(-1),
                            sphere_ids=sphere_ids, quad_ids=quad_ids,
                            perlins=np.array(perlins, scenes.PERLIN_DT).reshape(-1))


# CameraOpt functional options, camera.go:54-102
def WithSamplesPerPixel(samples):
    return lambda o: setattr(o, "spp", int(samples))


def WithMaxRayDepth(depth):
    return lambda o: setattr(o, "max_depth", int(depth))


def ToRadians(degrees):  # math.go:46-52
    return float(F(degrees) * F(math.pi / 180.0))


def WithFOVDegrees(fov):
    return lambda o: setattr(o, "fov_radians", ToRadians(fov))


def WithLookAt(v):
    return lambda o: setattr(o, "look_at", (C.c_float * 3)(*v))


def WithLookFrom(v):
    return lambda o: setattr(o, "look_from", (C.c_float * 3)(*v))


def WithDefocusAngleDegrees(deg):
    return lambda o: setattr(o, "defocus_angle_radians", ToRadians(deg))


def WithFocusDist(dist):
    return lambda o: setattr(o, "focus_dist", float(dist))


def WithBackgroundColor(color):
    return lambda o: setattr(o, "background", (C.c_float * 3)(*color))


class Camera:
    """camera.go:23-52.  `seed`/`device` have no counterpart in the reference (its RNG is
    clock-seeded, camera.go:170) and default to fixed values."""

    def __init__(self, options, seed=scenes.RENDER_SEED, device=0):
        self.options, self.seed, self.device = options, seed, device
        self.c = camera_from_options(options)  # Camera.init, camera.go:128-166
        self.last_stats = None

    def Render(self, world, writer):
        """camera.go:180-231: header, then one "R G B" line per pixel, row-major."""
        w, h = self.c.width, self.c.height
        writer.write("P3\n%d %d\n255\n" % (w, h))  # camera.go:183-188
        with Scene(flatten_world(world), self.device) as sc:
            rgb, _, self.last_stats = sc.render(self.c, self.seed)
        px = rgb.reshape(-1, 3)
        # chunked like stage.Agg(5000) + StartChunkRenderer (camera.go:225, 237-252)
        for b in range(0, len(px), 5000):
            writer.write("\n".join("%d %d %d" % (p[0], p[1], p[2]) for p in px[b:b + 5000]) + "\n")
        return None

    def RenderP6(self, world, writer):
        """Binary PPM (the reference's TODO at camera.go:196: its P3 text is ~12 bytes per pixel and,
        at 4K, costs more host time than the GPU render).  `writer` takes bytes."""
        w, h = self.c.width, self.c.height
        with Scene(flatten_world(world), self.device) as sc:
            rgb, _, self.last_stats = sc.render(self.c, self.seed)
        writer.write(b"P6\n%d %d\n255\n" % (w, h))
        writer.write(rgb.tobytes())
        return None

    def RenderPNG(self, world, writer):
        """PNG (the other half of the reference's TODO at camera.go:196).  `writer` takes bytes."""
        with Scene(flatten_world(world), self.device) as sc:
            rgb, _, self.last_stats = sc.render(self.c, self.seed)
        writer.write(encode_png(rgb))
        return None


def resolve_host(accum, total_spp):
    """camera.go:261 + vec3.go:141-166 on float32 sums that are already on the host: scale by 1/spp, sqrt,
    clamp to [0, 1], * 255.999, truncate — the resolve kernel's arithmetic (every step is one correctly rounded
    float32 operation, so numpy gives the same bytes)."""
    mean = np.asarray(accum, np.float32) * np.float32(1.0 / total_spp)
    with np.errstate(invalid="ignore"):
        g = np.sqrt(mean)
        g = np.where(g < 0, np.float32(0), np.where(g > 1, np.float32(1), g)).astype(np.float32)
        g = g * np.float32(255.999)
    return np.where(np.isnan(g), 0, g).astype(np.int32).astype(np.uint8)


def render_checkpointed(data, cam, path, seed=scenes.RENDER_SEED, chunk_spp=64, device=0, stop_after=None):
    """Resumable render (the reference has no checkpointing; SURVEY section 5 lists it as optional).  The
    cam.spp samples of every pixel are rendered in chunks of `chunk_spp` global sample indices; after each chunk
    the running FP32 sums and the sample cursor are written to `path` (.npz, replaced atomically).  A call that
    finds a checkpoint of the same frame continues at its cursor, so an interrupted render followed by a
    resumed one adds the same chunk sums in the same order as an uninterrupted one: bit-identical images.
    "The same frame" = same image shape, seed, spp, chunk size, camera bytes, scene bytes (SceneData.sha256 over
    spheres, quads, materials, textures, Perlin tables and texels) and library ABI version; anything else in the
    file is ignored and the render starts over.
    `stop_after` = number of chunks to render in this call (to interrupt deliberately).
    Returns (rgb or None while unfinished, sums, samples done)."""
    import os
    shape = (cam.height, cam.width, 3)
    done, acc = 0, np.zeros(shape, np.float32)
    scene_sha, abi_version = data.sha256(), int(abi.RT_B200_ABI_VERSION)
    if os.path.exists(path):
        with np.load(path) as z:
            if (tuple(z["shape"]) == shape and int(z["seed"]) == seed and int(z["spp"]) == cam.spp
                    and int(z["chunk_spp"]) == chunk_spp and bytes(z["camera"].tobytes()) == bytes(cam)
                    and "scene_sha" in z.files and str(z["scene_sha"]) == scene_sha
                    and int(z["abi_version"]) == abi_version):
                done, acc = int(z["done"]), z["acc"].astype(np.float32)
    chunks = 0
    with Scene(data, device) as sc:
        while done < cam.spp and (stop_after is None or chunks < stop_after):
            n = min(chunk_spp, cam.spp - done)
            _, part, _ = sc.render(cam, seed, done, n, want_accum=True)
            acc = acc + part
            done += n
            chunks += 1
            tmp = path + ".tmp.npz"
            np.savez(tmp, acc=acc, done=done, shape=np.array(shape), seed=seed, spp=cam.spp, chunk_spp=chunk_spp,
                     camera=np.frombuffer(bytes(cam), np.uint8), scene_sha=scene_sha, abi_version=abi_version)
            os.replace(tmp, path)
    return (resolve_host(acc, cam.spp) if done >= cam.spp else None), acc, done


def encode_png(rgb):
    """RGB8 (h, w, 3) -> PNG bytes: 8-bit truecolour, filter 0 on every scanline, one zlib stream."""
    import struct
    import zlib
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def NewCamera(aspectRatio, imageWidth, *opts, seed=scenes.RENDER_SEED, device=0):
    """camera.go:104-126 with its defaults."""
    o = abi.rt_camera_options()
    o.aspect_ratio = aspectRatio
    o.image_width = int(imageWidth)
    o.fov_radians = float(F(math.pi / 2))
    o.spp, o.max_depth = 100, 50
    o.focus_dist, o.defocus_angle_radians = 10.0, 0.0
    o.look_at = (C.c_float * 3)(0, 0, 0)
    o.look_from = (C.c_float * 3)(0, 0, -1)
    o.vup = (C.c_float * 3)(0, 1, 0)
    o.background = (C.c_float * 3)(0, 0, 0)
    for fn in opts:
        fn(o)
    return Camera(o, seed, device)
