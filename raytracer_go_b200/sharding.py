"""How one render is sharded over the GPUs of a box.  Pure index arithmetic plus the
one exchange step; used by bench.py under torchrun (one process per GPU, NCCL) and covered on CPU
by a world_size-2 gloo test.

Every (pixel, sample) path is independent (camera.go:202-218, 256-260) and its Philox stream is
keyed by the GLOBAL sample index, so the union of the ranks' work is exactly the single-GPU
render's sample set.

  sample-split  rank r renders global samples [offset, offset+count) of every pixel into a private
                FP32 W*H*3 accumulator; ONE reduce(sum) to rank 0, which resolves with the total
                spp.  This is the path's only exchange step.
  tile-split    rank r renders a contiguous band of scanlines at full spp; no reduction, the bands
                are gathered (concatenated) on rank 0.
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


def reduce_accumulators(accum, dst=0):
    """The sample-split exchange: sum the ranks' accumulators onto `dst` (NCCL over NVLink on GPUs,
    gloo in the CPU test).  `accum` is a torch tensor; returns it (valid on dst only)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
