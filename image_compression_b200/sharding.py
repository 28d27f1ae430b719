"""Row-stripe sharding of the 4x4 codecs across ranks: every 4x4 block depends only on its own
16 texels, output blocks are in raster order, so a stripe of whole block rows is one contiguous byte range of the
output.  No data-path collective is needed to encode; `gather_blocks` reassembles the packed stream on one rank
(NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def stripe_rows(grid_rows, rank, world):
    """Block-row range [r0, r1) of `rank`: contiguous stripes, the first grid_rows % world ranks take one extra."""
    base, extra = divmod(grid_rows, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def stripe_bytes(grid_rows, grid_cols, block_bytes, rank, world):
    r0, r1 = stripe_rows(grid_rows, rank, world)
    return (r1 - r0) * grid_cols * block_bytes


def gather_blocks(local, grid_rows, grid_cols, block_bytes, dst=0, group=None):
    """Gathers every rank's stripe of packed blocks (uint8 tensor) onto `dst`; returns the whole stream there and
    None elsewhere.  Equal stripes use one gather; uneven ones are padded to the largest stripe and trimmed."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [stripe_bytes(grid_rows, grid_cols, block_bytes, r, world) for r in range(world)]
    assert local.numel() == sizes[rank] and local.dtype == torch.uint8
    biggest = max(sizes)
    send = local if sizes[rank] == biggest else torch.cat([local, local.new_zeros(biggest - sizes[rank])])
    parts = [local.new_empty(biggest) for _ in range(world)] if rank == dst else None
    dist.gather(send, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([p[:n] for p, n in zip(parts, sizes)])
