"""TEST INFRASTRUCTURE: closest-hit parity of the device against the oracle over a whole primary frame, binned by
hit distance.  Used by tests/test_gpu_configs.py (assertions) and, run as a module on a GPU box, to write the report
committed under profiles/ (VERDICT r1 item 2a):

    python -m tests.parity_report C4 gpurun_out/c4_parity_by_distance.txt

For every primary ray of the config's frame (1 sample per pixel, the render seed) the device's rt_trace answer is
compared with (1) the reference's own BVH.Hit (bvh.go:220-249; oracle MODE_REF_BVH) on ALL rays and (2) the
brute-force World.Hit (hittables.go:55-72; MODE_LINEAR, the ground truth) on a seeded subset.
"""
import sys
import time

import numpy as np

EDGES = np.array([0, 25, 50, 75, 100, 125, 150, 200, 300, 500, 1000, np.inf])


def frame_rays(orc, cam, seed):
    return orc.primary_rays(cam, seed, 0, cam.width * cam.height, 0, 1)


def compare(ids, ts, rids, rts, dnorm):
    """Per distance bin (distance = the oracle's t * |d|; misses of the oracle in their own row):
    rays, ID mismatches, and among equal IDs the hits whose t differs in any bit."""
    hit = rids >= 0
    dist = np.where(hit, rts * dnorm, np.nan)
    rows = []
    for lo, hi in zip(EDGES[:-1], EDGES[1:]):
        m = hit & (dist >= lo) & (dist < hi)
        same = m & (ids == rids)
        rows.append((lo, hi, int(m.sum()), int((m & (ids != rids)).sum()),
                     int((same & (ts.view(np.uint32) != rts.view(np.uint32))).sum())))
    miss = ~hit
    rows.append((np.nan, np.nan, int(miss.sum()), int((miss & (ids != rids)).sum()), 0))
    return rows


def format_rows(rows, title):
    out = [title, f"{'hit distance':>16} {'rays':>10} {'ID mismatches':>14} {'rate':>10} {'t bits differ (same ID)':>24}"]
    for lo, hi, n, bad, tbad in rows:
        name = "oracle: miss" if np.isnan(lo) else f"[{lo:g}, {hi:g})"
        out.append(f"{name:>16} {n:>10} {bad:>14} {(bad / n if n else 0):>10.2e} {tbad:>24}")
    n = sum(r[2] for r in rows)
    bad = sum(r[3] for r in rows)
    out.append(f"{'all':>16} {n:>10} {bad:>14} {(bad / n if n else 0):>10.2e} {sum(r[4] for r in rows):>24}")
    return "\n".join(out)


# A second camera for the 1 M-sphere scene: low above the ground, looking across the grid, so that primary rays hit
# spheres up to ~1000 units away — far beyond the 150-unit envelope inside which the reference's float32
# discriminant still resolves a 0.2-radius sphere.
FAR_CAMERA = dict(look_from=(52.0, 3.0, 12.0), look_at=(-448.0, 0.2, -103.0))


def run(config, n_linear=60_000, spp_seed=None, far=False, full_bvh=True):
    from oracle import pyoracle as orc
    from raytracer_go_b200 import api, scenes
    seed = scenes.RENDER_SEED if spp_seed is None else spp_seed
    scene, o = scenes.build_config(config, spp=1)
    if far:
        o = scenes.camera_options(o.image_width, 1, **FAR_CAMERA)
    cam = api.camera_from_options(o)
    ro, rd = frame_rays(orc, cam, seed)
    dnorm = np.linalg.norm(rd.astype(np.float64), axis=1).astype(np.float32)
    t0 = time.time()
    with api.Scene(scene) as sc:
        ids, ts = sc.trace(ro, rd)
    t_dev = time.time() - t0
    sub = np.sort(np.random.default_rng(1).choice(len(ro), min(n_linear, len(ro)), replace=False))
    t0 = time.time()
    if full_bvh:
        bids, bts = orc.trace(scene, ro, rd, mode=orc.MODE_REF_BVH)
    else:  # only where the list is evaluated too
        bids, bts = np.full(len(ro), -2, np.int32), np.zeros(len(ro), np.float32)
        bids[sub], bts[sub] = orc.trace(scene, ro[sub], rd[sub], mode=orc.MODE_REF_BVH)
    t_bvh = time.time() - t0
    t0 = time.time()
    lids, lts = orc.trace(scene, ro[sub], rd[sub], mode=orc.MODE_LINEAR)
    t_lin = time.time() - t0
    return dict(scene=scene, cam=cam, ro=ro, rd=rd, ids=ids, ts=ts, bids=bids, bts=bts, sub=sub, lids=lids, lts=lts, dnorm=dnorm,
                seconds=(t_dev, t_bvh, t_lin))


def report(config, far):
    r = run(config, far=far, full_bvh=not far)
    cam, sub = r["cam"], r["sub"]
    text = [f"{config}{' — FAR camera ' + repr(FAR_CAMERA) if far else ''}: {len(r['scene'].spheres)} spheres, {cam.width}x{cam.height} primary rays (1 per pixel), "
            f"ray_origin_radius (envelope) {r['scene'].ray_origin_radius:g}",
            f"device rt_trace {r['seconds'][0]:.2f} s (includes scene creation), oracle BVH.Hit {r['seconds'][1]:.1f} s, "
            f"oracle World.Hit on {len(sub)} rays {r['seconds'][2]:.1f} s", ""]
    text.append(format_rows(compare(r["ids"][sub], r["ts"][sub], r["lids"], r["lts"], r["dnorm"][sub]),
                            f"device vs World.Hit (brute-force list, ground truth), {len(sub)} seeded rays"))
    text.append("")
    text.append(format_rows(compare(r["bids"][sub], r["bts"][sub], r["lids"], r["lts"], r["dnorm"][sub]),
                            f"for scale — the reference's own BVH.Hit vs its World.Hit, the same {len(sub)} rays"))
    if not far:
        text.append("")
        text.append(format_rows(compare(r["ids"], r["ts"], r["bids"], r["bts"], r["dnorm"]),
                                f"device vs the reference's BVH.Hit (random-axis median split, bvh.go:142-249), all {len(r['ids'])} rays"))
    return "\n".join(text)


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "C4"
    out = sys.argv[2] if len(sys.argv) > 2 else None
    text = report(config, False)
    if config == "C4":
        text += "\n\n" + report(config, True)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
