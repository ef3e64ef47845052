"""Random scenes for the traversal fuzz tests (CPU: host-compiled headers; GPU: the CUDA path)."""
import numpy as np

from raytracer_go_b200 import scenes


def fuzz_scene_and_rays(seed, m=6000):
    """1..400 spheres with radii over four decades, clustered / coincident / nested centres, negative radii
    (hollow shells, as the book's glass sphere); rays from inside, on and outside them.
    Returns (scene, origins, dirs, origin_radius)."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([1, 2, 3, 7, 40, 400]))
    sph = np.zeros(n, scenes.SPHERE_DT)
    centres = rng.normal(scale=rng.choice([0.5, 3.0, 20.0]), size=(n, 3))
    if n > 3:
        centres[rng.integers(0, n, n // 4)] = centres[0]            # coincident centres
        centres[1] = centres[2]                                       # nested pair
    sph["cx"], sph["cy"], sph["cz"] = centres[:, 0], centres[:, 1], centres[:, 2]
    sph["r"] = 10.0 ** rng.uniform(-2, 2, n) * rng.choice([1.0, 1.0, 1.0, -1.0], n)
    pick = rng.integers(0, n, m)
    c = centres[pick]
    v = rng.normal(size=(m, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    o = (c + v * (np.abs(sph["r"][pick]) * rng.choice([0.0, 0.5, 1.0, 1.0, 1.5, 4.0], m))[:, None]).astype(np.float32)
    d = (rng.normal(size=(m, 3)) * rng.choice([0.1, 1.0, 30.0], m)[:, None]).astype(np.float32)
    radius = float(np.abs(o.astype(np.float64)).max() + np.abs(sph["r"]).max() + np.abs(centres).max()) * 2 + 1
    s = scenes.SceneData(sph, np.zeros(1, scenes.MATERIAL_DT), np.zeros(1, scenes.TEXTURE_DT),
                         ray_origin_radius=radius, name=f"fuzz{seed}")
    return s, o, d, radius
