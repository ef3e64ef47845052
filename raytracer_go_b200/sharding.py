"""How one render is sharded over the GPUs of a box.  Pure index arithmetic plus the
one exchange step; used by bench.py under torchrun (one process per GPU, NCCL) and covered on CPU
by a world_size-2 gloo test.

Every (pixel, sample) path is independent (camera.go:202-218, 256-260) and its Philox stream is
keyed by the GLOBAL sample index, so the union of the ranks' work is exactly the single-GPU
render's sample set.

  sample-split  rank r renders global samples [offset, offset+count) of every pixel into a private
                FP32 W*H*3 accumulator; ONE reduce(sum) to rank 0, which resolves with the total
                spp.  This is the path's only exchange step.
  tile-split    rank r renders scanlines r, r+N, r+2N, ... at full spp (interleaved: sky rows are
                cheap, ground rows are not, contiguous bands would not balance); no reduction, the
                rows are gathered on rank 0 and interleaved.  Bit-identical to the single-GPU render.
"""


def sample_split_weak(rank, world, spp_per_rank):
    """Fixed work per GPU: rank r takes samples [r*spp, (r+1)*spp); the image has world*spp spp."""
    return rank * spp_per_rank, spp_per_rank, world * spp_per_rank


def sample_split_strong(rank, world, total_spp):
    """Fixed total: total_spp divided as evenly as possible, remainders to the low ranks."""
    base, rem = divmod(total_spp, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count, total_spp


def tile_split(rank, world, height):
    """Contiguous scanline bands [row_begin, row_end)."""
    base, rem = divmod(height, world)
    rows = base + (1 if rank < rem else 0)
    begin = rank * base + min(rank, rem)
    return begin, begin + rows


def row_split(rank, world, height):
    """Interleaved scanlines: (row_begin, row_count, row_step) for rt_render_opts."""
    return rank, max(0, (height - rank + world - 1) // world), world


def gather_rows(local, height, dst=0):
    """The tile-split exchange: every rank holds its row_split() rows as a (row_count, W, C) tensor;
    returns the (height, W, C) image on `dst` (None elsewhere).  One gather, no arithmetic."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank()
    most = (height + world - 1) // world  # ranks may differ by one row: pad to the largest share
    padded = local
    if local.shape[0] < most:
        padded = torch.cat([local, local.new_zeros((most - local.shape[0],) + tuple(local.shape[1:]))])
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded.contiguous(), parts, dst=dst)
    if rank != dst:
        return None
    out = local.new_empty((height,) + tuple(local.shape[1:]))
    for r in range(world):
        _, count, step = row_split(r, world, height)
        out[r::step] = parts[r][:count]
    return out


def reduce_accumulators(accum, dst=0):
    """The sample-split exchange: sum the ranks' accumulators onto `dst` (NCCL over NVLink on GPUs,
    gloo in the CPU test).  `accum` is a torch tensor; returns it (valid on dst only)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
