"""Multi-GPU host logic on CPU: index arithmetic, and a world_size-2 gloo run of the sample-split
and tile-split plans in which each rank's share is rendered by the oracle (the kernels cannot run
here) and exchanged exactly as bench.py does on GPUs."""
import os
import socket

import numpy as np
import pytest

from raytracer_go_b200 import scenes, sharding


def test_sample_split_ranges_partition():
    for world in (1, 2, 3, 4, 8):
        for total in (1, 7, 500, 4096):
            got = []
            for r in range(world):
                off, cnt, tot = sharding.sample_split_strong(r, world, total)
                assert tot == total
                got += list(range(off, off + cnt))
            assert got == list(range(total))
        offs = [sharding.sample_split_weak(r, world, 500) for r in range(world)]
        assert [o[0] for o in offs] == [500 * r for r in range(world)] and offs[0][2] == 500 * world


def test_tile_split_partition():
    for world in (1, 2, 4, 8, 7):
        for h in (1, 225, 675, 2160):
            rows = []
            for r in range(world):
                b, e = sharding.tile_split(r, world, h)
                rows += list(range(b, e))
            assert rows == list(range(h))


def test_row_split_partition():
    for world in (1, 2, 3, 4, 8, 7):
        for h in (1, 5, 225, 675, 2160):
            rows = []
            for r in range(world):
                b, n, step = sharding.row_split(r, world, h)
                assert step == world and (n == 0 or b + (n - 1) * step < h)
                rows += [b + k * step for k in range(n)]
            assert sorted(rows) == list(range(h))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import pyoracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = scenes.random_scene()
    cam = orc.camera_from_options(scenes.camera_options(64, 6))
    # sample-split: private accumulators, one reduce, resolve on rank 0 with the total spp
    off, cnt, total = sharding.sample_split_strong(rank, world, 6)
    _, acc, _ = orc.render(scene, cam, 9, sample_offset=off, sample_count=cnt, order=orc.ORDER_ITERATIVE, threads=2)
    t = torch.from_numpy(acc.copy())
    sharding.reduce_accumulators(t, dst=0)
    # tile-split: bands gathered on rank 0
    b, e = sharding.tile_split(rank, world, cam.height)
    rgb, _, _ = orc.render(scene, cam, 9, order=orc.ORDER_ITERATIVE, threads=2, rows=(b, e))
    band = torch.from_numpy(rgb[b:e].copy())
    bands = [torch.empty((sharding.tile_split(r, world, cam.height)[1] - sharding.tile_split(r, world, cam.height)[0],
                          cam.width, 3), dtype=torch.uint8) for r in range(world)] if rank == 0 else None
    dist.gather(band, bands, dst=0)
    # interleaved rows (what bench.py --split tile and rt_render_multi's tile mode do); 45 rows on 2 ranks
    # is a ragged split (23 + 22), which exercises gather_rows' padding
    rb, rn, rs = sharding.row_split(rank, world, cam.height)
    whole, _, _ = orc.render(scene, cam, 9, order=orc.ORDER_ITERATIVE, threads=2)  # the oracle has no strided row set
    mine = torch.from_numpy(whole[rb::rs].copy())
    assert mine.shape[0] == rn
    inter = sharding.gather_rows(mine, cam.height, dst=0)
    if rank == 0:
        q.put((t.numpy(), torch.cat(bands).numpy(), total, inter.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sample_and_tile_split(orc):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    acc, tiled, total, inter = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scene = scenes.random_scene()
    cam = orc.camera_from_options(scenes.camera_options(64, 6))
    rgb, full, _ = orc.render(scene, cam, 9, order=orc.ORDER_ITERATIVE)
    # same per-sample radiances, summed in a different order: equal to rounding
    np.testing.assert_allclose(acc, full, rtol=2e-6, atol=1e-6)
    assert (np.abs(orc.resolve(acc, total).astype(int) - rgb.astype(int)) <= 1).all()
    assert np.array_equal(tiled, rgb)  # tile-split is bitwise the single-process image
    assert np.array_equal(inter, rgb)
