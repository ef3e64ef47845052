"""The BVH built on the GPU (csrc/bvh_device.cuh: Morton codes, radix tree, bottom-up fit, leaf collapse) against the
host-built tree and the oracle.  The tree's topology is free (the closest hit is an argmin over all primitives), so
every comparison is on RESULTS: object IDs and the bits of t from rt_trace, FP32 sums from rt_render.
RT_B200_DEVICE_BVH is read when a scene is created: 0 = host builder, 2 = device builder for any quad-free scene."""
import numpy as np
import pytest

from raytracer_go_b200 import api, scenes
from tests import fuzz_scenes

pytestmark = pytest.mark.gpu
SEED = scenes.RENDER_SEED


def _scene(data, mode, monkeypatch):
    monkeypatch.setenv("RT_B200_DEVICE_BVH", str(mode))
    return api.Scene(data)


def _rays(orc, scene_rng_seed, cam, n_secondary):
    ro, rd = orc.primary_rays(cam, SEED, 0, cam.width * cam.height, 0, 1)
    rng = np.random.default_rng(scene_rng_seed)
    so = rng.uniform([-12, 0.01, -12], [12, 2.5, 12], size=(n_secondary, 3)).astype(np.float32)
    sd = rng.normal(size=(n_secondary, 3)).astype(np.float32)
    return np.concatenate([ro, so]), np.concatenate([rd, sd])


def test_device_built_random_scene_equals_oracle_and_host_build(gpu, orc, random_scene, monkeypatch):
    cam = api.camera_from_options(scenes.camera_options(400, 3))
    o, d = _rays(orc, 3, cam, 200_000)
    with _scene(random_scene, 2, monkeypatch) as sc:
        info = sc.bvh_info()
        assert info.built_on_device == 1 and info.n_slots == len(random_scene.spheres) and 0 < info.max_depth < 62
        ids, ts = sc.trace(o, d)
        rgb, acc, st = sc.render(cam, SEED, want_accum=True)
        nodes, slot_ids, _ = sc.bvh_copy()
    assert sorted(slot_ids.tolist()) == list(range(len(random_scene.spheres)))     # every sphere in exactly one slot
    with _scene(random_scene, 0, monkeypatch) as sc:
        assert sc.bvh_info().built_on_device == 0
        hids, hts = sc.trace(o, d)
        hrgb, hacc, hst = sc.render(cam, SEED, want_accum=True)
    assert np.array_equal(ids, hids) and np.array_equal(ts.view(np.uint32), hts.view(np.uint32))
    assert np.array_equal(acc.view(np.uint32), hacc.view(np.uint32)) and np.array_equal(rgb, hrgb) and st.rays == hst.rays
    rids, rts = orc.trace(random_scene, o[:60_000], d[:60_000])
    assert np.array_equal(ids[:60_000], rids)
    hit = rids >= 0
    assert np.array_equal(ts[:60_000][hit].view(np.uint32), rts[hit].view(np.uint32))


@pytest.mark.parametrize("seed", range(8))
def test_device_builder_on_fuzz_scenes(gpu, orc, seed, monkeypatch):
    """1..400 spheres, radii over four decades, coincident and nested centres, negative radii: built on the device where
    the builder accepts the scene (it declines scenes with more than 64 'huge' spheres), the answer is World.Hit's."""
    s, o, d, radius = fuzz_scenes.fuzz_scene_and_rays(seed)
    with _scene(s, 2, monkeypatch) as sc:
        ids, ts = sc.trace(o, d)
    rids, rts = orc.trace(s, o, d)
    assert np.array_equal(ids, rids)
    hit = rids >= 0
    assert np.array_equal(ts[hit].view(np.uint32), rts[hit].view(np.uint32))


def test_device_built_tree_with_sphere_ids_and_far_camera_rebuild(gpu, orc, monkeypatch):
    """Object IDs given as a permutation survive the Morton sort; a camera far outside the radius the boxes were padded
    for triggers the rebuild with a larger radius (ensure_origin_radius): same results as the host-built tree."""
    base = scenes.random_scene()
    rng = np.random.default_rng(11)
    perm = rng.permutation(len(base.spheres)).astype(np.uint32)
    s = scenes.SceneData(base.spheres, base.materials, base.textures, sphere_ids=perm, name="permuted")
    cam = api.camera_from_options(scenes.camera_options(160, 2, look_from=(130, 20, 30)))
    ro, rd = orc.primary_rays(cam, SEED, 0, cam.width * cam.height, 0, 1)
    with _scene(s, 2, monkeypatch) as sc:
        assert sc.bvh_info().built_on_device == 1
        rgb, acc, _ = sc.render(cam, SEED, want_accum=True)
        ids, ts = sc.trace(ro, rd)
    with _scene(s, 0, monkeypatch) as sc:
        hrgb, hacc, _ = sc.render(cam, SEED, want_accum=True)
        hids, hts = sc.trace(ro, rd)
    assert np.array_equal(ids, hids) and np.array_equal(ts.view(np.uint32), hts.view(np.uint32))
    assert np.array_equal(acc.view(np.uint32), hacc.view(np.uint32)) and np.array_equal(rgb, hrgb)
    rids, _ = orc.trace(s, ro, rd)
    assert np.array_equal(ids, rids) and set(np.unique(ids[ids >= 0])) <= set(perm.tolist())


def test_c4_device_build_is_the_default_and_equals_the_host_build(gpu, orc, monkeypatch):
    """Config C4 (1 M spheres) takes the device builder by default; closest hits of 300 000 primary rays and a
    2-spp frame are bit-identical to the host-built (binned SAH) tree's."""
    scene, o = scenes.build_config("C4", spp=2)
    cam = api.camera_from_options(o)
    ro, rd = orc.primary_rays(cam, SEED, 0, cam.width * cam.height, 0, 1)
    pick = np.sort(np.random.default_rng(5).choice(len(ro), 300_000, replace=False))
    monkeypatch.delenv("RT_B200_DEVICE_BVH", raising=False)
    with api.Scene(scene) as sc:
        info = sc.bvh_info()
        assert info.built_on_device == 1 and info.in_shared_memory == 0 and info.max_depth < 62
        ids, ts = sc.trace(ro[pick], rd[pick])
        rgb, acc, st = sc.render(cam, SEED, want_accum=True)
    with _scene(scene, 0, monkeypatch) as sc:
        assert sc.bvh_info().built_on_device == 0
        hids, hts = sc.trace(ro[pick], rd[pick])
        hrgb, hacc, hst = sc.render(cam, SEED, want_accum=True)
    assert np.array_equal(ids, hids) and np.array_equal(ts.view(np.uint32), hts.view(np.uint32))
    # secondary rays can leave the scene's 150-unit envelope: there the two trees may cull
    # different noise-level candidates, so the frames agree almost everywhere rather than everywhere
    same = (acc.view(np.uint32) == hacc.view(np.uint32)).all(-1)
    assert same.mean() > 0.995, f"{(~same).sum()} pixels differ"
    assert abs(int(st.rays) - int(hst.rays)) < 1e-3 * hst.rays


def test_large_scene_through_render_multi(gpu, monkeypatch):
    """rt_render_multi on a scene every device builds for itself (device builder): the raw arrays are uploaded once to
    devices[0] and copied to the others over NVLink (DeviceStaging).  Tile-split is bitwise the single-GPU render,
    sample-split equal up to the FP32 summation order; with one GPU the call degenerates to rt_render."""
    monkeypatch.delenv("RT_B200_DEVICE_BVH", raising=False)
    scene = scenes.stress_scene(130)          # 67 600 cells -> above the device builder's 50 000-sphere threshold
    assert len(scene.spheres) > 60_000
    cam = api.camera_from_options(scenes.camera_options(320, 6, look_from=(52, 24, 12)))
    with api.Scene(scene) as sc:
        assert sc.bvh_info().built_on_device == 1
        rgb, acc, st = sc.render(cam, SEED, want_accum=True)
    devs = list(range(min(gpu, 4)))
    trgb, tacc, tst = api.render_multi(scene, cam, devs, SEED, want_accum=True, tile_split=True)
    assert np.array_equal(trgb, rgb) and np.array_equal(tacc.view(np.uint32), acc.view(np.uint32)) and tst.samples == st.samples
    srgb, sacc, sst = api.render_multi(scene, cam, devs, SEED, want_accum=True)
    assert sst.rays == st.rays and np.allclose(sacc, acc, rtol=2e-5, atol=1e-5)
    assert (np.abs(srgb.astype(int) - rgb.astype(int)) <= 1).all()


def test_horizon_beams_overflow_their_lists_and_fall_back(gpu, monkeypatch):
    """A camera low above a plane of 67 000 spheres: the beams of the pixels near the horizon touch far more leaves than
    a candidate list holds; those pixels keep the tree traversal (the beam walk stops at the overflow).  With and without
    the lists the frame is bit-identical."""
    from tests import parity_report
    scene = scenes.stress_scene(130)
    cam = api.camera_from_options(scenes.camera_options(240, 16, **parity_report.FAR_CAMERA))
    frames = []
    for lists in ("1", "0"):
        monkeypatch.setenv("RT_B200_PIXEL_LISTS", lists)
        with api.Scene(scene) as sc:
            frames.append(sc.render(cam, SEED, want_accum=True))
    (rgb1, acc1, st1), (rgb0, acc0, st0) = frames
    assert np.array_equal(acc1.view(np.uint32), acc0.view(np.uint32)) and np.array_equal(rgb1, rgb0) and st1.rays == st0.rays
    assert st1.kernel_launches == st0.kernel_launches + 1 and st1.ms_render < 20 * st0.ms_render + 50
