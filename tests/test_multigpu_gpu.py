"""Two-GPU check of the stripe path on real devices (skipped on a one-GPU box): each rank encodes its block-row stripe
(a) into local memory, gathered with NCCL, and (b) straight into rank 0's buffer through the peer mapping
(sharding.PeerStream, the fused gather); both must equal the oracle's encoding of the whole image.  PVRTC: halo stripes, Z-order stores into rank 0's buffer."""
import os
import socket

import numpy as np
import pytest
import torch

import checkers as ck

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    import torch.distributed as dist

    import image_compression_b200 as icb
    from image_compression_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = True
    for codec, fmt, nc, bb, h, w in ((icb.CODEC_DXT1, icb.RGBA, 4, 8, 1024, 512), (icb.CODEC_DXT5, icb.RGBA, 4, 16, 520, 256),
                                     (icb.CODEC_DXT1, icb.RGB, 3, 8, 260, 512), (icb.CODEC_ETC1, icb.RGB, 3, 8, 64, 256)):
        pitch = w * nc
        whole = ck.synthetic(h * pitch, 5)
        grid_rows, grid_cols = h // 4, w // 4
        r0, r1 = sharding.stripe_rows(grid_rows, rank, world)
        mine = torch.from_numpy(whole[r0 * 4 * pitch:r1 * 4 * pitch].copy()).cuda()
        base = mine.data_ptr() - r0 * 4 * pitch
        local = torch.empty((r1 - r0) * grid_cols * bb, dtype=torch.uint8, device="cuda")
        icb.encode_stripe_device(codec, fmt, base, h, w, pitch, h, w, r0, r1, local)
        ps = sharding.PeerStream(grid_rows * grid_cols * bb, dst=0)
        icb.encode_stripe_device(codec, fmt, base, h, w, pitch, h, w, r0, r1, ps.stripe_ptr(r0 * grid_cols * bb))
        ps.complete()
        gathered = sharding.gather_blocks(local, grid_rows, grid_cols, bb, dst=0)
        if rank == 0:
            if codec == icb.CODEC_ETC1:
                want = ck.oracle_etc1(ck.ETC_SMALLER_ERROR, whole, h, w)
            elif nc == 4 and codec == icb.CODEC_DXT1:
                want = ck.oracle_dxt1_rgba(whole, h, w)
            else:
                want = ck.oracle_dxt(ck.RGB if nc == 3 else ck.RGBA, whole, h, w)
            ok = ok and np.array_equal(gathered.cpu().numpy(), want) and np.array_equal(ps.tensor().cpu().numpy(), want)
        ps.close()
    # PVRTC: stripes with a one-block-row halo each side, blocks stored at their Z-order slots of rank 0's buffer
    n = 256
    img = ck.synthetic(n * n * 4, 6)
    lh = n // 4
    r0, r1 = sharding.stripe_rows(lh, rank, world)
    ys = sharding.pvrtc_stripe_row_indices(n, r0, r1)
    rows = torch.from_numpy(np.ascontiguousarray(img.reshape(n, n * 4)[ys]).ravel()).cuda()
    first = torch.from_numpy(img[:4].copy()).cuda()
    ps = sharding.PeerStream(n * n // 4, dst=0)
    icb.pvrtc_encode_stripe_device(rows, first, n, n, r0, r1, ps.stripe_ptr(0))
    ps.complete()
    if rank == 0:
        ok = ok and np.array_equal(ps.tensor().cpu().numpy(), ck.oracle_pvrtc(img, n, n))
    ps.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        results.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_stripes_nccl_gather_and_peer_stores_match_oracle():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert results.get(timeout=5) == 1


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_host_path_spread_over_two_gpus(icb):
    """ICB_HOST_DEVICES=2: one Compress() call uploads alternate chunks to two GPUs, each over its own PCIe link, and
    every GPU writes its blocks to their place in the caller's buffer.  Same bytes as the oracle, from pageable and
    from pinned memory, for sizes with several chunks, a ragged bottom edge, and a padded (CompressAndPad) grid."""
    import ctypes as C
    import os
    old = os.environ.get("ICB_HOST_DEVICES")
    os.environ["ICB_HOST_DEVICES"] = "2"
    try:
        L = icb.lib()
        for codec, fmt, nc, h, w in ((icb.CODEC_DXT1, icb.RGBA, 4, 4096, 4096), (icb.CODEC_DXT5, icb.RGBA, 4, 2050, 4096),
                                     (icb.CODEC_DXT1, icb.RGB, 3, 3001, 2048), (icb.CODEC_ETC1, icb.RGB, 3, 512, 8192)):
            img = ck.synthetic(h * w * nc, 9)
            if codec == icb.CODEC_ETC1:
                want = ck.oracle_etc1(ck.ETC_SMALLER_ERROR, img, h, w)
            elif codec == icb.CODEC_DXT1 and nc == 4:
                want = ck.oracle_dxt1_rgba(img, h, w)
            else:
                want = ck.oracle_dxt(ck.RGB if nc == 3 else ck.RGBA, img, h, w)
            got = icb.compress_host(codec, fmt, img, h, w)
            assert np.array_equal(got, want), ("pageable", codec, h, w)
            pin_in, pin_out = L.icb_host_alloc(img.size), L.icb_host_alloc(want.size)
            a = np.ctypeslib.as_array(C.cast(pin_in, C.POINTER(C.c_uint8)), shape=(img.size,))
            b = np.ctypeslib.as_array(C.cast(pin_out, C.POINTER(C.c_uint8)), shape=(want.size,))
            a[:] = img
            icb.compress_host(codec, fmt, a, h, w, out=b)
            assert np.array_equal(b, want), ("pinned", codec, h, w)
            L.icb_host_free(pin_in)
            L.icb_host_free(pin_out)
        # CompressAndPad below the image stays on one device (row h-1 replication) and must still be right
        h, w = 1030, 4096
        img = ck.synthetic(h * w * 4, 10)
        got = icb.compress_host(icb.CODEC_DXT5, icb.RGBA, img, h, w, padded=(2048, 4096))
        assert np.array_equal(got, ck.oracle_dxt(ck.RGBA, img, h, w, coded_h=2048, coded_w=4096))
        assert torch.cuda.current_device() == 0
    finally:
        if old is None:
            del os.environ["ICB_HOST_DEVICES"]
        else:
            os.environ["ICB_HOST_DEVICES"] = old


def _want(icb, codec, fmt, img, h, w):
    if codec == icb.CODEC_ETC1:
        return ck.oracle_etc1(ck.ETC_SMALLER_ERROR, img, h, w)
    if codec == icb.CODEC_DXT1 and ck.ncomp(fmt) == 4:
        return ck.oracle_dxt1_rgba(img, h, w, swap_rb=1 if fmt == ck.BGRA else 0)
    return ck.oracle_dxt(fmt, img, h, w)


def _sharded_case(icb, ctx, codec, fmt, nc, h, w, share):
    n = len(ctx.devices)
    pitch = w * nc
    img = ck.synthetic(h * pitch, 12)
    splits = icb.stripe_partition(n, (h + 3) // 4, share)
    stripes = []
    for r in range(n):
        y0, y1 = min(4 * splits[r], h), min(4 * splits[r + 1], h)
        stripes.append(torch.from_numpy(img[y0 * pitch:y1 * pitch].copy()).to("cuda:%d" % ctx.devices[r]) if y1 > y0 else None)
    with torch.cuda.device(ctx.devices[0]):
        out = torch.zeros(icb.compressed_size(codec, h, w), dtype=torch.uint8, device="cuda:%d" % ctx.devices[0])
        ctx.encode(codec, fmt, stripes, h, w, splits, out)
        torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), _want(icb, codec, fmt, img, h, w)), (codec, fmt, h, w, share, splits)


def test_shard_context_one_device(icb):
    """icb_ctx_* / icb_encode_sharded with a single device (what a one-GPU box can run): same bytes as the oracle, the
    call is stream-ordered, the caller's device is untouched."""
    with icb.ShardContext([0]) as ctx:
        assert ctx.devices == [0]
        for codec, fmt, nc, h, w in ((icb.CODEC_DXT1, icb.RGBA, 4, 256, 512), (icb.CODEC_DXT5, icb.RGBA, 4, 130, 256),
                                     (icb.CODEC_DXT1, icb.BGR, 3, 64, 256), (icb.CODEC_ETC1, icb.RGB, 3, 32, 256)):
            _sharded_case(icb, ctx, codec, fmt, nc, h, w, -1)
        img = ck.synthetic(512 * 512 * 4, 13)
        assert np.array_equal(ctx.compress_host(icb.CODEC_DXT5, icb.RGBA, img, 512, 512), ck.oracle_dxt(ck.RGBA, img, 512, 512))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_shard_context_single_process_multi_gpu(icb):
    """One process, N GPUs, no IPC: every device encodes its stripe and stores the blocks into the root's buffer through
    cudaDeviceEnablePeerAccess mappings.  Even and root-heavy partitions, ragged heights, every 4x4 codec; then the
    host-buffer form over the same devices, and icb_set_host_devices as the programmatic ICB_HOST_DEVICES."""
    ndev = torch.cuda.device_count()
    for ids in ([0, 1], [1, 0], list(range(min(ndev, 8)))):
        with icb.ShardContext(ids) as ctx:
            assert ctx.peer_stores
            for codec, fmt, nc, h, w in ((icb.CODEC_DXT1, icb.RGBA, 4, 1024, 512), (icb.CODEC_DXT5, icb.RGBA, 4, 522, 256),
                                         (icb.CODEC_DXT1, icb.RGB, 3, 260, 512), (icb.CODEC_ETC1, icb.RGB, 3, 64, 256)):
                for share in (-1, 475, 900):
                    _sharded_case(icb, ctx, codec, fmt, nc, h, w, share)
            img = ck.synthetic(4096 * 4096 * 4, 14)
            want = ck.oracle_dxt1_rgba(img, 4096, 4096)
            assert np.array_equal(ctx.compress_host(icb.CODEC_DXT1, icb.RGBA, img, 4096, 4096), want)
            assert torch.cuda.current_device() == 0
    prev = icb.lib().icb_set_host_devices(2)
    try:
        assert np.array_equal(icb.compress_host(icb.CODEC_DXT1, icb.RGBA, img, 4096, 4096), want)
    finally:
        icb.lib().icb_set_host_devices(-1 if prev == 0x7fffffff else prev)
