"""Row-stripe sharding of the 4x4 codecs across ranks: every 4x4 block depends only on its own
16 texels, output blocks are in raster order, so a stripe of whole block rows is one contiguous byte range of the
output.  No data-path collective is needed to encode; `gather_blocks` reassembles the packed stream on one rank
(NCCL on GPUs, gloo in the CPU tests).  `PeerStream` is the B200-native form of that gather: the owner's output
buffer is mapped into every rank over NVLink (CUDA IPC) and each rank's encoder stores its blocks straight into it,
so the gather is fused into the encode kernel's epilogue and costs no separate pass."""
import ctypes as C

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


def pvrtc_stripe_row_indices(height, r0, r1):
    """Image rows a rank must hold to encode PVRTC block rows [r0, r1) (8x4 blocks, height/4 block rows): the
    stripe plus one block row of halo above and below, wrapped round the torus, in the order
    icb_pvrtc2_encode_stripe expects them (first row = 4*(r0-1) mod height)."""
    lh = height // 4
    assert 0 <= r0 < r1 <= lh and r1 - r0 + 2 <= lh
    return [(4 * (r0 - 1) + k) % height for k in range(4 * (r1 - r0 + 2))]


def gather_blocks(local, grid_rows, grid_cols, block_bytes, dst=0, group=None, splits=None):
    """Gathers every rank's stripe of packed blocks (uint8 tensor) onto `dst`; returns the whole stream there and
    None elsewhere.  Equal stripes use one gather; uneven ones are padded to the largest stripe and trimmed.
    `splits` (world + 1 block-row boundaries, e.g. from icb_stripe_partition) overrides the stripe_rows partition."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if splits is not None:
        assert len(splits) == world + 1 and splits[0] == 0 and splits[-1] == grid_rows
        sizes = [(splits[r + 1] - splits[r]) * grid_cols * block_bytes for r in range(world)]
    else:
        sizes = [stripe_bytes(grid_rows, grid_cols, block_bytes, r, world) for r in range(world)]
    assert local.numel() == sizes[rank] and local.dtype == torch.uint8
    biggest = max(sizes)
    send = local if sizes[rank] == biggest else torch.cat([local, local.new_zeros(biggest - sizes[rank])])
    parts = [local.new_empty(biggest) for _ in range(world)] if rank == dst else None
    dist.gather(send, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([p[:n] for p, n in zip(parts, sizes)])


class PeerStream:
    """The packed block stream of one sharded image, owned by rank `dst` and store-mapped into every other rank.

        ps = PeerStream(total_bytes, dst=0)            # collective: allocates on dst, ships the IPC handle, maps it
        icb.encode_stripe_device(..., out=ps.stripe_ptr(byte_offset_of_my_stripe))
        ps.complete()                                   # collective: every rank's stores have landed on dst
        whole = ps.tensor()                             # on dst: uint8 view of the whole stream (a copy); None elsewhere
        ps.close()

    One process per GPU (torch.distributed initialised, NCCL backend); the handle travels through a broadcast."""

    def __init__(self, total_bytes, dst=0, group=None):
        from . import binding
        self._lib = binding.lib()
        self._check = binding._check
        self.total_bytes, self.dst, self.group = int(total_bytes), dst, group
        self.rank = dist.get_rank(group)
        self.owner = self.rank == dst
        self._base = C.c_void_p()
        handle = torch.zeros(64, dtype=torch.uint8)
        if self.owner:
            self._check(self._lib.icb_device_alloc(self.total_bytes, C.byref(self._base)))
            buf = (C.c_uint8 * 64)()
            self._check(self._lib.icb_ipc_export(self._base, buf))
            handle = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device())
        wire = handle.to(dev)
        dist.broadcast(wire, src=dst, group=group)
        error = None
        if not self.owner:
            raw = C.create_string_buffer(wire.cpu().numpy().tobytes(), 64)
            try:
                self._check(self._lib.icb_ipc_open(raw, C.byref(self._base)))
            except Exception as e:  # e.g. no peer access between the two GPUs
                error = e
        # every rank learns whether every mapping succeeded, so that nobody is left waiting in a later collective
        ok = torch.tensor([0 if error else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError("PeerStream: mapping rank %d's buffer failed on at least one rank%s" % (dst, (": %s" % error) if error else ""))

    def stripe_ptr(self, byte_offset):
        assert 0 <= byte_offset <= self.total_bytes
        return self._base.value + int(byte_offset)

    def complete(self):
        """Every rank: wait for the local encoder, then meet the others; after this the owner holds the whole stream."""
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def tensor(self):
        if not self.owner:
            return None
        class _Span:  # zero-copy view of the cudaMalloc'ed stream for torch
            __cuda_array_interface__ = {"shape": (self.total_bytes,), "typestr": "|u1", "data": (self._base.value, False),
                                        "version": 3}
        return torch.as_tensor(_Span(), device="cuda").clone()

    def close(self):
        """Collective: every rank unmaps, then the owner frees (nobody may still have the buffer mapped)."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        if not self.owner and self._base.value:
            self._check(self._lib.icb_ipc_close(self._base))
        dist.barrier(group=self.group)
        if self.owner and self._base.value:
            self._check(self._lib.icb_device_free(self._base))
        self._base = C.c_void_p()
