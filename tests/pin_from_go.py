"""TEST INFRASTRUCTURE: the reference pin.

The reference holds no golden vectors and Go is not installed where this repository is built, so the oracle is pinned
by a fixture that the REAL Go code writes wherever Go exists:

  tests/golden/pin_inputs.json        seeded inputs, written by make_inputs() below (committed)
  raytracer_go_b200/go/parity_dump_test.go   package internal: evaluates the inputs with the reference's own functions
  tests/golden/from_go/pin_outputs.json      what it writes (scripts/pin_from_go.sh); commit it once produced
  tests/test_pin_from_go.py           compares the oracle — and the CUDA library where the C ABI exposes the function —
                                      with that file, bit for bit; skipped (with the reason) while the file is absent

Every float32 is stored as its bit pattern.  oracle_outputs() evaluates the same inputs with the oracle in the same
schema; tests/golden/pin_expected_by_oracle.json is that evaluation, committed so that a maintainer running the Go side
can diff the two files directly, and so that the comparison code itself is exercised without Go.

    python -m tests.pin_from_go            # rewrites pin_inputs.json and pin_expected_by_oracle.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from raytracer_go_b200 import abi, scenes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUTS = os.path.join(GOLDEN, "pin_inputs.json")
EXPECTED = os.path.join(GOLDEN, "pin_expected_by_oracle.json")
FROM_GO = os.path.join(GOLDEN, "from_go", "pin_outputs.json")
VERSION = 1
F = np.float32
INF_BITS = 0x7F800000
PIN_SEED = 0xC0FFEE  # Philox seed of the GetRay cases


def bits(x):
    """float32 value(s) -> JSON-able bit pattern(s)."""
    a = np.asarray(x, dtype=np.float32)
    b = a.view(np.uint32)
    return int(b) if a.ndim == 0 else [int(v) for v in b.ravel()]


def unbits(b):
    a = np.asarray(b, dtype=np.uint32)
    return a.view(np.float32)


def _k24(u):
    """uniforms in [0,1) with 24 random bits -> the integers k = u * 2^24 the Go feedSource takes."""
    return [int(v) for v in np.round(np.asarray(u, np.float64) * 16777216.0).astype(np.int64)]


# ---- inputs ------------------------------------------------------------------------------------------------------------
def make_inputs():
    from oracle import pyoracle as orc
    rng = np.random.default_rng(20261019)
    inp = {"version": VERSION}

    # textures / materials shared by the scatter, texture and get_color sections
    img = scenes.procedural_earth_map(64, 32).astype(np.uint16)          # (h, w, 3) RGB16
    per = scenes.new_perlin(0x5EED0002)
    inp["images"] = [{"w": 64, "h": 32, "rgb16": [int(v) for v in img.ravel()]}]
    inp["perlins"] = [{"vec": [bits(v) for v in per["vec"].reshape(256, 3)],
                       "perm_x": [int(v) for v in per["perm_x"].ravel()], "perm_y": [int(v) for v in per["perm_y"].ravel()],
                       "perm_z": [int(v) for v in per["perm_z"].ravel()]}]
    z3 = bits(np.zeros(3, F))
    inp["textures"] = [
        {"kind": "solid", "a": bits(F([0.4, 0.2, 0.1])), "b": z3, "scale": 0, "image": 0, "perlin": 0},
        {"kind": "checker", "a": bits(F([0.2, 0.3, 0.1])), "b": bits(F([0.9, 0.9, 0.9])), "scale": bits(F(0.32)), "image": 0, "perlin": 0},
        {"kind": "image", "a": z3, "b": z3, "scale": 0, "image": 0, "perlin": 0},
        {"kind": "noise", "a": z3, "b": z3, "scale": bits(F(4.0)), "image": 0, "perlin": 0},
        {"kind": "solid", "a": bits(F([4.0, 4.0, 4.0])), "b": z3, "scale": 0, "image": 0, "perlin": 0},
    ]
    mat = lambda kind, albedo=(0, 0, 0), fuzz=0.0, ior=0.0, texture=0: {  # noqa: E731
        "kind": kind, "albedo": bits(F(albedo)), "fuzz": bits(F(fuzz)), "ior": bits(F(ior)), "texture": texture}
    inp["materials"] = [
        mat("lambertian", texture=0), mat("lambertian", texture=1), mat("lambertian", texture=2), mat("lambertian", texture=3),
        mat("metal", (0.7, 0.6, 0.5), 0.0), mat("metal", (0.8, 0.8, 0.9), 0.3), mat("metal", (0.5, 0.9, 0.6), 1.0),
        mat("dielectric", ior=1.5), mat("dielectric", ior=1.0 / 1.5), mat("light", texture=4),
    ]

    def ray_case(o, d, tmin=0.001, tmax=np.inf, **kw):
        c = {"o": bits(F(o)), "d": bits(F(d)), "tmin": bits(F(tmin)), "tmax": bits(F(tmax))}
        c.update(kw)
        return c

    # Sphere.Hit: rays aimed near the sphere (hit / graze / miss), from outside, the surface and the inside
    cases = []
    for k in range(240):
        c = rng.normal(scale=3.0, size=3)
        r = 10.0 ** rng.uniform(-1.5, 1.5) * (1 if k % 7 else -1)
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        o = c + v * abs(r) * rng.choice([0.0, 0.6, 1.0, 1.5, 5.0, 40.0])
        target = c + rng.normal(size=3) * abs(r) * rng.choice([0.3, 0.9, 1.0, 1.2])
        d = (target - o) * rng.choice([0.1, 1.0, 7.0])
        if not np.any(d):
            d = v
        tmax = np.inf if k % 3 else float(np.linalg.norm(target - o) / max(np.linalg.norm(d), 1e-9))
        cases.append(ray_case(o, d, 0.001, tmax, c=bits(F(c)), r=bits(F(r))))
    inp["sphere_hit"] = cases
    # Quad.Hit
    cases = []
    for k in range(120):
        q = rng.normal(scale=2.0, size=3)
        u, v = rng.normal(size=3) * rng.uniform(0.2, 3), rng.normal(size=3) * rng.uniform(0.2, 3)
        if k % 5 == 0:  # axis-aligned, as Box() makes them
            u, v = np.array([rng.uniform(0.5, 2), 0, 0]), np.array([0, rng.uniform(0.5, 2), 0])
        target = q + u * rng.uniform(-0.2, 1.2) + v * rng.uniform(-0.2, 1.2)
        o = target + rng.normal(size=3) * rng.uniform(0.5, 6)
        cases.append(ray_case(o, (target - o) * rng.choice([0.2, 1.0, 3.0]), q=bits(F(q)), u=bits(F(u)), v=bits(F(v))))
    inp["quad_hit"] = cases
    # Aabb.Hit incl. axis-parallel rays and origins on a face
    cases = []
    for k in range(240):
        a, b = rng.normal(scale=2.0, size=3), rng.normal(scale=2.0, size=3)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        o = rng.normal(scale=4.0, size=3)
        target = lo + (hi - lo) * rng.uniform(-0.3, 1.3, size=3)
        d = target - o
        if k % 4 == 0:
            d[rng.integers(0, 3)] = 0.0
        if k % 6 == 0:
            ax = rng.integers(0, 3)
            o[ax] = F(lo[ax]) if k % 12 else F(hi[ax])
        tmax = np.inf if k % 3 else rng.uniform(0.2, 2.0)
        cases.append(ray_case(o, d, 0.001, tmax, min=bits(F(lo)), max=bits(F(hi))))
    inp["aabb_hit"] = cases
    # World.Hit / BVH.Hit over the random scene: primary rays + scattered rays
    s = scenes.random_scene()
    cam = orc.camera_from_options(scenes.camera_options(64, 1))
    ro, rd = orc.primary_rays(cam, 2024, 0, 64 * 36, 0, 1)
    pick = np.sort(rng.choice(len(ro), 900, replace=False))
    so = rng.uniform([-11, 0.01, -11], [11, 2.2, 11], size=(600, 3)).astype(F)
    sd = rng.normal(size=(600, 3)).astype(F)
    rays = np.concatenate([np.concatenate([ro[pick], rd[pick]], 1), np.concatenate([so, sd], 1)]).astype(F)
    sph = np.stack([s.spheres["cx"], s.spheres["cy"], s.spheres["cz"], s.spheres["r"]], 1).astype(F)
    inp["world"] = {"spheres": [bits(v) for v in sph], "rays": [bits(v) for v in rays], "tmin": bits(F(0.001)), "tmax": INF_BITS}
    # Scatter / Emit at the hit of one sphere
    cases = []
    for k in range(150):
        m = k % len(inp["materials"])
        r = F(1.0) if k % 4 else F(-0.8)     # a negative radius flips the normal (hollow shell)
        c = rng.normal(scale=0.5, size=3)
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        o = c + v * abs(r) * (rng.choice([0.3, 3.0]) if inp["materials"][m]["kind"] == "dielectric" else 3.0)
        target = c + rng.normal(size=3) * 0.55 * abs(r)
        feed = rng.integers(0, 1 << 24, size=96)
        cases.append({"material": m, "sphere": bits(F([c[0], c[1], c[2], r])), "o": bits(F(o)), "d": bits(F(target - o)),
                      "feed": [int(x) for x in feed], "seed": 1000 + k})
    inp["scatter"] = cases
    # GetTexture
    cases = []
    for k in range(200):
        u, v = rng.uniform(-0.1, 1.25), rng.uniform(-0.1, 1.1)
        if k % 10 == 0:
            u, v = rng.choice([0.0, 1.0]), rng.choice([0.0, 1.0])
        cases.append({"texture": k % 4, "u": bits(F(u)), "v": bits(F(v)), "p": bits(F(rng.normal(scale=3.0, size=3)))})
    inp["texture"] = cases
    # resolve
    vals = rng.uniform(0, 1.3, size=(60, 3)).astype(F) ** F(2.0)
    vals = np.concatenate([vals, F([[0, 1, 0.25], [0.7, 0.8, 1.0], [1e-12, 0.9999999, 1.0000001], [4.0, 0.5, 0.0625]])])
    inp["resolve"] = [bits(v) for v in vals]
    # cameras (main.go:228-239, 194-205, 106-118 and variations)
    def cam_case(width, fov, defocus, focus, frm, at, bg=(0.7, 0.8, 1.0), aspect=16.0 / 9.0, spp=10, depth=50):
        return {"aspect": bits(F(aspect)), "width": width, "fov_deg": bits(F(fov)), "defocus_deg": bits(F(defocus)),
                "focus_dist": bits(F(focus)), "look_from": bits(F(frm)), "look_at": bits(F(at)), "background": bits(F(bg)),
                "spp": spp, "depth": depth}
    inp["camera"] = [cam_case(400, 20, 0.6, 10, (13, 2, 3), (0, 0, 0)), cam_case(1200, 20, 0.6, 10, (13, 2, 3), (0, 0, 0)),
                     cam_case(600, 40, 0, 10, (278, 278, -800), (278, 278, 0), (0, 0, 0), 1.0),
                     cam_case(400, 80, 0, 10, (0, 0, 9), (0, 0, 0)), cam_case(1, 20, 0.6, 10, (13, 2, 3), (0, 0, 0)),
                     cam_case(3840, 20, 2.5, 3.7, (-2, 5, 1), (0.5, 0.25, -1))]
    # GetRay: the uniforms are the Philox stream of (seed, pixel, sample), so the device's rt_primary_rays is pinned too
    cases = []
    for k in range(96):
        ci = [0, 1, 5][k % 3]
        w = inp["camera"][ci]["width"]
        h = int(np.floor(w) / float(unbits(inp["camera"][ci]["aspect"])))
        i, j, smp = int(rng.integers(0, w)), int(rng.integers(0, h)), int(rng.integers(0, 4096))
        u = orc.rng_floats(PIN_SEED, j * w + i, smp, 64)
        cases.append({"camera": ci, "i": i, "j": j, "sample": smp, "feed": _k24(u)})
    inp["get_ray"] = cases
    # reflect / refract / reflectance
    def unit(v):
        return v / np.linalg.norm(v)
    inp["reflect"] = [{"a": bits(F(unit(rng.normal(size=3)))), "b": bits(F(unit(rng.normal(size=3)))), "eta": 0, "cos": 0}
                      for _ in range(64)]
    cases = []
    for _ in range(64):
        n = unit(rng.normal(size=3))
        uv = unit(-n + rng.normal(size=3) * rng.uniform(0, 1.5))
        cases.append({"a": bits(F(uv)), "b": bits(F(n)), "eta": bits(F(rng.choice([1.5, 1 / 1.5, 1.0, 2.4]))), "cos": 0})
    inp["refract"] = cases
    inp["reflectance"] = [{"a": z3, "b": z3, "eta": bits(F(rng.choice([1.5, 1 / 1.5, 1.33]))),
                           "cos": bits(F(rng.choice([0.0, 1.0, rng.uniform(0, 1)])))} for _ in range(64)]
    # Ray.GetColor over spheres + quads, no dielectric, constant feed (u = 0.25: the cube sample (-0.5,-0.5,-0.5))
    gs = F([[0, -100.5, -1, 100], [0, 0, -1.2, 0.5], [-1.0, 0, -1.0, 0.5], [1.0, 0, -1.0, 0.5], [0, 1.4, -1.0, 0.3]])
    gq = F([[-3, -0.5, -3, 6, 0, 0, 0, 4, 0], [-0.5, 2.5, -1.5, 1, 0, 0, 0, 0, 1]])
    cam = orc.camera_from_options(scenes.camera_options(16, 1, look_from=(0, 0.4, 2.0), look_at=(0, 0, -1), vfov_deg=70.0,
                                                        defocus_deg=0.0))
    go, gd = orc.primary_rays(cam, 7, 0, 16 * 9, 0, 1)
    inp["get_color"] = {"spheres": [bits(v) for v in gs], "sphere_materials": [1, 0, 5, 4, 9],
                        "quads": [bits(v) for v in gq], "quad_materials": [3, 9], "background": bits(F([0.7, 0.8, 1.0])),
                        "depth": 12, "feed": [1 << 22], "rays": [bits(v) for v in np.concatenate([go, gd], 1)]}
    return inp


# ---- the oracle's answers in the output schema -----------------------------------------------------------------------
TEX_KIND = {"solid": abi.RT_TEX_SOLID, "checker": abi.RT_TEX_CHECKER, "image": abi.RT_TEX_IMAGE, "noise": abi.RT_TEX_NOISE}
MAT_KIND = {"lambertian": abi.RT_MAT_LAMBERTIAN, "metal": abi.RT_MAT_METAL, "dielectric": abi.RT_MAT_DIELECTRIC,
            "light": abi.RT_MAT_DIFFUSE_LIGHT}


def _tables(inp):
    tex = np.zeros(len(inp["textures"]), scenes.TEXTURE_DT)
    for i, t in enumerate(inp["textures"]):
        tex[i]["kind"], tex[i]["a"], tex[i]["b"] = TEX_KIND[t["kind"]], unbits(t["a"]), unbits(t["b"])
        tex[i]["scale"] = unbits(t["scale"])
        tex[i]["image"] = t["perlin"] if t["kind"] == "noise" else t["image"]
        tex[i]["oob"] = (0, 0, 0)          # image.RGBA64 outside its bounds: the zero colour
    mats = np.zeros(len(inp["materials"]), scenes.MATERIAL_DT)
    for i, m in enumerate(inp["materials"]):
        mats[i]["kind"], mats[i]["albedo"] = MAT_KIND[m["kind"]], unbits(m["albedo"])
        mats[i]["fuzz"], mats[i]["ior"], mats[i]["texture"] = unbits(m["fuzz"]), unbits(m["ior"]), m["texture"]
    images = [np.asarray(im["rgb16"], np.uint16).reshape(im["h"], im["w"], 3) for im in inp["images"]]
    per = np.zeros(len(inp["perlins"]), scenes.PERLIN_DT)
    for i, p in enumerate(inp["perlins"]):
        per[i]["vec"] = unbits(p["vec"]).reshape(256, 3)
        per[i]["perm_x"], per[i]["perm_y"], per[i]["perm_z"] = p["perm_x"], p["perm_y"], p["perm_z"]
    return tex, mats, images, per


def _scene(inp, spheres, sphere_mats, quads=None, quad_mats=None):
    tex, mats, images, per = _tables(inp)
    sp = np.zeros(len(spheres), scenes.SPHERE_DT)
    for i, s in enumerate(spheres):
        v = unbits(s)
        sp[i]["cx"], sp[i]["cy"], sp[i]["cz"], sp[i]["r"], sp[i]["material"] = v[0], v[1], v[2], v[3], sphere_mats[i]
    qd = None
    if quads:
        qd = np.zeros(len(quads), scenes.QUAD_DT)
        for i, q in enumerate(quads):
            v = unbits(q)
            qd[i]["q"], qd[i]["u"], qd[i]["v"], qd[i]["material"] = v[0:3], v[3:6], v[6:9], quad_mats[i]
    return scenes.SceneData(sp, mats, tex, images=images, quads=qd, perlins=per, name="pin")


def _hit_out(h):
    if h is None:
        return {"hit": False, "t": 0, "point": [0, 0, 0], "normal": [0, 0, 0], "front": False, "u": 0, "v": 0}
    return {"hit": True, "t": bits(h["t"]), "point": bits(h["point"]), "normal": bits(h["normal"]),
            "front": bool(h["front_face"]), "u": bits(h["u"]), "v": bits(h["v"])}


def _blocks3(feed_k):
    """Go draws x, y, z of a trial one after the other; the oracle takes u[0..2] of a 4-value block per trial."""
    u = np.asarray(feed_k, np.float64) / 16777216.0
    n = len(u) // 3
    out = np.zeros((n, 4), np.float32)
    out[:, :3] = u[:3 * n].reshape(n, 3)
    return out.ravel()


def camera_of(case):
    """pin camera case -> rt_camera_options with the reference's degree conversion (math.go:46-52)."""
    o = scenes.camera_options(case["width"], case["spp"], max_depth=case["depth"], look_from=tuple(unbits(case["look_from"])),
                              look_at=tuple(unbits(case["look_at"])), vfov_deg=float(unbits(case["fov_deg"])),
                              defocus_deg=float(unbits(case["defocus_deg"])), focus_dist=float(unbits(case["focus_dist"])),
                              background=tuple(unbits(case["background"])), aspect=float(unbits(case["aspect"])))
    return o


def oracle_outputs(inp, dielectric_uniforms=None):
    """The oracle's answers.  dielectric_uniforms: per scatter case, the uniform the Go side's global source drew
    (from its fixture); None = 0.5 everywhere (the committed oracle-only file)."""
    from oracle import pyoracle as orc
    out = {"producer": "oracle", "version": inp["version"]}
    dummy_mat = [0]
    res = []
    for c in inp["sphere_hit"]:
        sc = _scene(inp, [list(c["c"]) + [c["r"]]], dummy_mat)
        res.append(_hit_out(orc.hit_info(sc, unbits(c["o"]), unbits(c["d"]), float(unbits(c["tmin"])), float(unbits(c["tmax"])))))
    out["sphere_hit"] = res
    res = []
    for c in inp["quad_hit"]:
        sc = _scene(inp, [], [], [list(c["q"]) + list(c["u"]) + list(c["v"])], dummy_mat)
        res.append(_hit_out(orc.hit_info(sc, unbits(c["o"]), unbits(c["d"]), float(unbits(c["tmin"])), float(unbits(c["tmax"])))))
    out["quad_hit"] = res
    out["aabb_hit"] = [bool(orc.aabb_hit(unbits(c["min"]), unbits(c["max"]), unbits(c["o"]), unbits(c["d"]),
                                         float(unbits(c["tmin"])), float(unbits(c["tmax"])))) for c in inp["aabb_hit"]]
    w = inp["world"]
    sc = _scene(inp, w["spheres"], [0] * len(w["spheres"]))
    rays = unbits(w["rays"]).reshape(-1, 6)
    ids, ts = orc.trace(sc, rays[:, :3], rays[:, 3:], float(unbits(w["tmin"])), float(unbits(w["tmax"])), mode=orc.MODE_LINEAR)
    out["world_hit"] = {"ids": [int(i) for i in ids], "t": [bits(t) if i >= 0 else 0 for i, t in zip(ids, ts)]}
    ids, ts = orc.trace(sc, rays[:, :3], rays[:, 3:], float(unbits(w["tmin"])), float(unbits(w["tmax"])), mode=orc.MODE_REF_BVH)
    out["bvh_hit"] = {"ids": [int(i) for i in ids], "t": [bits(t) if i >= 0 else 0 for i, t in zip(ids, ts)]}
    res = []
    for k, c in enumerate(inp["scatter"]):
        sc = _scene(inp, [c["sphere"]], [c["material"]])
        kind = inp["materials"][c["material"]]["kind"]
        uni = F(0.5) if dielectric_uniforms is None else unbits(dielectric_uniforms[k])
        feed = np.full(4, uni, F) if kind == "dielectric" else _blocks3(c["feed"])
        r = orc.scatter_fed(sc, unbits(c["o"]), unbits(c["d"]), feed)
        z = [0, 0, 0]
        if r is None:
            res.append({"hit": False, "scattered": False, "origin": z, "dir": z, "attenuation": z, "emitted": z, "uniform": bits(uni)})
        else:
            s = r["scattered"]
            res.append({"hit": True, "scattered": s, "origin": bits(r["origin"]) if s else z, "dir": bits(r["dir"]) if s else z,
                        "attenuation": bits(r["attenuation"]) if s else z, "emitted": bits(r["emitted"]), "uniform": bits(uni)})
    out["scatter"] = res
    sc = _scene(inp, [bits(F([0, 0, 0, 1]))], [0])
    out["texture"] = [bits(orc.texture(sc, c["texture"], float(unbits(c["u"])), float(unbits(c["v"])), unbits(c["p"])))
                      for c in inp["texture"]]
    res = []
    for c in inp["resolve"]:
        rgb = orc.encode_pixel(unbits(c))
        res.append({"rgb": [int(v) for v in rgb], "text": "%d %d %d" % tuple(int(v) for v in rgb)})
    out["resolve"] = res
    cams, res = [], []
    for c in inp["camera"]:
        o = camera_of(c)
        cam = orc.camera_from_options(o)
        cams.append(cam)
        res.append({"center": bits(F(list(cam.center))), "pixel00": bits(F(list(cam.pixel00))), "du": bits(F(list(cam.pixel_du))),
                    "dv": bits(F(list(cam.pixel_dv))), "disk_u": bits(F(list(cam.defocus_u))), "disk_v": bits(F(list(cam.defocus_v))),
                    "height": int(cam.height), "fov_radians": bits(F(o.fov_radians)), "defocus_angle": bits(F(o.defocus_angle_radians))})
    out["camera"] = res
    res = []
    for c in inp["get_ray"]:
        o, d = orc.get_ray_fed(cams[c["camera"]], c["i"], c["j"], np.asarray(c["feed"], np.float64) / 16777216.0)
        res.append({"origin": bits(o), "dir": bits(d)})
    out["get_ray"] = res
    out["reflect"] = [bits(orc.reflect(unbits(c["a"]), unbits(c["b"]))) for c in inp["reflect"]]
    out["refract"] = [bits(orc.refract(unbits(c["a"]), unbits(c["b"]), float(unbits(c["eta"])))) for c in inp["refract"]]
    out["reflectance"] = [bits(F(orc.reflectance(float(unbits(c["cos"])), float(unbits(c["eta"]))))) for c in inp["reflectance"]]
    g = inp["get_color"]
    sc = _scene(inp, g["spheres"], g["sphere_materials"], g["quads"], g["quad_materials"])
    feed = np.asarray(g["feed"], np.float64) / 16777216.0
    rays = unbits(g["rays"]).reshape(-1, 6)
    out["get_color"] = [bits(orc.get_color_fed(sc, r[:3], r[3:], unbits(g["background"]), g["depth"], feed)) for r in rays]
    return out


# ---- the device's answers where the C ABI exposes the function --------------------------------------------------------
def device_outputs(inp):
    """rt_trace (World.Hit), rt_camera_from_options (Camera.init), rt_primary_rays (Camera.GetRay) on the GPU."""
    from raytracer_go_b200 import api
    out = {"producer": "device", "version": inp["version"]}
    w = inp["world"]
    sc = _scene(inp, w["spheres"], [0] * len(w["spheres"]))
    rays = unbits(w["rays"]).reshape(-1, 6)
    with api.Scene(sc) as s:
        ids, ts = s.trace(np.ascontiguousarray(rays[:, :3]), np.ascontiguousarray(rays[:, 3:]), float(unbits(w["tmin"])),
                          float(unbits(w["tmax"])))
    out["world_hit"] = {"ids": [int(i) for i in ids], "t": [bits(t) if i >= 0 else 0 for i, t in zip(ids, ts)]}
    cams, res = [], []
    for c in inp["camera"]:
        o = camera_of(c)
        cam = api.camera_from_options(o)
        cams.append(cam)
        res.append({"center": bits(F(list(cam.center))), "pixel00": bits(F(list(cam.pixel00))), "du": bits(F(list(cam.pixel_du))),
                    "dv": bits(F(list(cam.pixel_dv))), "disk_u": bits(F(list(cam.defocus_u))), "disk_v": bits(F(list(cam.defocus_v))),
                    "height": int(cam.height), "fov_radians": bits(F(o.fov_radians)), "defocus_angle": bits(F(o.defocus_angle_radians))})
    out["camera"] = res
    res = []
    for c in inp["get_ray"]:
        cam = cams[c["camera"]]
        o, d = api.primary_rays(cam, PIN_SEED, c["j"] * cam.width + c["i"], 1, c["sample"], 1)
        res.append({"origin": bits(o[0]), "dir": bits(d[0])})
    out["get_ray"] = res
    return out


# ---- comparison --------------------------------------------------------------------------------------------------------
EXACT_SECTIONS = ["sphere_hit", "quad_hit", "aabb_hit", "world_hit", "scatter", "texture", "resolve", "camera", "get_ray",
                  "reflect", "refract", "reflectance", "get_color"]
# float64 libm results (Go's pure-Go Cephes ports against glibc / CUDA) that may differ in the last bit after rounding to
# float32: sphere UV (acos, atan2), noise / marble (sin), Schlick (pow), camera tan.  One ulp is allowed THERE only.
ULP_FIELDS = {("sphere_hit", "u"), ("sphere_hit", "v"), ("reflectance", None), ("camera", "pixel00"), ("camera", "du"),
              ("camera", "dv"), ("camera", "disk_u"), ("camera", "disk_v")}


def _ulp_diff(a, b):
    a, b = np.asarray(a, np.uint32).astype(np.int64), np.asarray(b, np.uint32).astype(np.int64)
    return int(np.max(np.abs(a - b))) if a.size else 0


def compare(ref, got, sections=None):
    """-> list of human-readable mismatches between two output dicts (ref = the Go fixture)."""
    bad = []
    for sec in sections or EXACT_SECTIONS:
        if sec not in ref or sec not in got:
            bad.append(f"{sec}: missing")
            continue
        a, b = ref[sec], got[sec]
        if isinstance(a, dict):  # world_hit
            for k in a:
                if a[k] != b[k]:
                    n = sum(1 for x, y in zip(a[k], b[k]) if x != y)
                    bad.append(f"{sec}.{k}: {n} of {len(a[k])} differ")
            continue
        if len(a) != len(b):
            bad.append(f"{sec}: {len(a)} cases vs {len(b)}")
            continue
        for i, (x, y) in enumerate(zip(a, b)):
            if x == y:
                continue
            if isinstance(x, dict):
                for k in x:
                    if x[k] != y[k]:
                        if (sec, k) in ULP_FIELDS and _ulp_diff(x[k], y[k]) <= 1:
                            continue
                        if sec == "texture" or k == "uniform":
                            continue
                        bad.append(f"{sec}[{i}].{k}: {x[k]} != {y[k]}")
            elif (sec, None) in ULP_FIELDS and _ulp_diff(x, y) <= 1:
                continue
            elif sec == "texture" and _ulp_diff(x, y) <= 1:  # marble: float64 sin; image / checker / solid are exact anyway
                continue
            else:
                bad.append(f"{sec}[{i}]: {x} != {y}")
    return bad


def load(path):
    with open(path) as f:
        return json.load(f)


def main():
    inp = make_inputs()
    with open(INPUTS, "w") as f:
        json.dump(inp, f, separators=(",", ":"))
    exp = oracle_outputs(inp)
    with open(EXPECTED, "w") as f:
        json.dump(exp, f, separators=(",", ":"))
    print(f"wrote {INPUTS} ({os.path.getsize(INPUTS) >> 10} KiB) and {EXPECTED} ({os.path.getsize(EXPECTED) >> 10} KiB)")


if __name__ == "__main__":
    main()
