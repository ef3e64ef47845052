"""The synthetic scene builders are deterministic and follow main.go's recipes."""
import numpy as np

from raytracer_go_b200 import abi, scenes


def test_random_scene_recipe():
    s = scenes.random_scene()
    assert s.sha256() == scenes.random_scene().sha256()
    sp = s.spheres
    assert 470 <= len(sp) <= 488                       # 22*22 cells minus those near (4, 0.2, 0), + 4
    assert tuple(sp[0]) [:4] == (0, -1000, 0, 1000) and s.textures[0]["kind"] == abi.RT_TEX_CHECKER  # main.go:242-244
    small = sp[1:-3]
    assert (small["r"] == np.float32(0.2)).all() and (small["cy"] == np.float32(0.2)).all()
    d = np.sqrt((small["cx"] - 4) ** 2 + small["cz"] ** 2)
    assert (d > 0.9 - 1e-6).all()                      # main.go:254-256
    kinds = s.materials["kind"][small["material"]]
    frac = np.bincount(kinds, minlength=3) / len(kinds)
    assert 0.7 < frac[0] < 0.9 and 0.08 < frac[1] < 0.22 and 0.01 < frac[2] < 0.1   # 80 / 15 / 5 %
    met = s.materials[small["material"]][kinds == abi.RT_MAT_METAL]
    assert (met["albedo"] >= 0.5).all() and (met["albedo"] < 1).all() and (met["fuzz"] < 0.5).all()
    big = sp[-3:]
    assert [tuple(b)[:4] for b in big] == [(0, 1, 0, 1), (-4, 1, 0, 1), (4, 1, 0, 1)]   # main.go:278-285
    assert s.materials["kind"][big["material"]].tolist() == [abi.RT_MAT_DIELECTRIC, abi.RT_MAT_LAMBERTIAN,
                                                             abi.RT_MAT_METAL]
    # cell (i, j) = (-11, -11) is the first small sphere; its centre is i + 0.9*U
    assert -11 <= small["cx"][0] < -10.1 + 1e-6 and -11 <= small["cz"][0] < -10.1 + 1e-6


def test_seed_changes_scene():
    assert scenes.random_scene(seed=1).sha256() != scenes.random_scene(seed=2).sha256()


def test_stress_scene_scales():
    s = scenes.random_scene(half=40, seed=scenes.SCENE_SEED_STRESS)
    assert 6300 < len(s.spheres) <= 6404 and len(s.materials) == len(s.spheres)


def test_earth_scene_and_texture():
    s = scenes.earth_random_scene(tex_w=256, tex_h=128)
    assert s.images[0].shape == (128, 256, 3) and s.images[0].dtype == np.uint16
    assert (s.images[0] % 257 == 0).all()              # RGB8 widened x257
    assert s.textures[-1]["kind"] == abi.RT_TEX_IMAGE
    assert abs(float(s.textures[-1]["oob"][1]) - 34678 / 65535) < 1e-6
    c = np.array(scenes.EARTH_CENTER)
    others = s.spheres[1:-1]
    d = np.sqrt((others["cx"] - c[0]) ** 2 + (others["cy"] - c[1]) ** 2 + (others["cz"] - c[2]) ** 2)
    assert (d > 2 + others["r"]).all()                 # nothing overlaps the earth sphere
    img2 = scenes.procedural_earth_map(256, 128)
    assert np.array_equal(img2, s.images[0])


def test_config_table():
    for name, (w, h) in {"C1": (1200, 675), "C2": (1200, 675), "C3": (1920, 1080), "C5": (3840, 2160)}.items():
        scene, o = scenes.build_config(name, stress_half=8)
        assert o.image_width == w and o.spp == scenes.CONFIGS[name]["spp"] and o.max_depth == 50
