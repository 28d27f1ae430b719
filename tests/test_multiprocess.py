"""N > 1 host logic on CPU: two gloo ranks shard an image by block-row stripes, encode their stripes (with the CPU
oracle standing in for the GPU encoder -- this test is about partitioning and reassembly, not about kernels), gather
the packed streams on rank 0 and compare with the single-rank result.  Uneven splits included."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import checkers as ck
import imagegen


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from image_compression_b200 import sharding
    ok = True
    for (fmt, h, w) in ((ck.RGBA, 36, 24), (ck.RGB, 21, 16), (ck.BGRA, 8, 8), (ck.RGB, 50, 12)):
        nc = ck.ncomp(fmt)
        bb = 8 if nc == 3 else 16
        img = imagegen.make("smooth_noise", h, w, nc, seed=31)
        grid_rows, grid_cols = (h + 3) // 4, (w + 3) // 4
        r0, r1 = sharding.stripe_rows(grid_rows, rank, world)
        y0, y1 = 4 * r0, min(4 * r1, h)
        if r1 > r0:
            stripe = np.ascontiguousarray(img[y0:y1])
            local = ck.oracle_dxt(fmt, stripe.ravel(), y1 - y0, w)
        else:
            local = np.zeros(0, np.uint8)
        whole = sharding.gather_blocks(torch.from_numpy(local), grid_rows, grid_cols, bb, dst=0)
        if rank == 0:
            want = ck.oracle_dxt(fmt, img.ravel(), h, w)
            ok = ok and np.array_equal(whole.numpy(), want)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        results.put(int(flag.item()))
    dist.destroy_process_group()


def test_two_rank_stripes_reassemble_to_single_rank_output():
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert results.get(timeout=5) == 1


def test_stripe_partition_properties():
    from image_compression_b200 import sharding
    for rows in (1, 2, 7, 2048, 2051):
        for world in (1, 2, 3, 8):
            spans = [sharding.stripe_rows(rows, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert sum(sharding.stripe_bytes(rows, 5, 8, r, world) for r in range(world)) == rows * 5 * 8


def test_pvrtc_stripe_rows_wrap_and_cover():
    """Halo stripes: every rank's resident rows start one block row above its stripe (wrapped), cover the stripe in
    order, end one block row below it (wrapped); together the stripes' own rows tile the image exactly once."""
    from image_compression_b200 import sharding
    for h in (32, 64, 256):
        lh = h // 4
        for world in (2, 3, 4, 8):
            if lh // world + 3 > lh:
                continue
            own = []
            for r in range(world):
                r0, r1 = sharding.stripe_rows(lh, r, world)
                rows = sharding.pvrtc_stripe_row_indices(h, r0, r1)
                assert len(rows) == 4 * (r1 - r0 + 2) and len(set(rows)) == len(rows)
                assert rows[0] == (4 * (r0 - 1)) % h and rows[-1] == (4 * (r1 + 1) - 1) % h
                assert rows[4:-4] == list(range(4 * r0, 4 * r1))
                own += rows[4:-4]
            assert own == list(range(h))
