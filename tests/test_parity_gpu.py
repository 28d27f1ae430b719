"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the reference-generated
golden fixtures.  Bit-exact (integer/byte work): every comparison is array_equal."""
import numpy as np
import pytest
import torch

import checkers as ck
import imagegen

pytestmark = pytest.mark.gpu

CODEC = {"dxt1": 0, "dxt5": 1, "etc1": 2, "pvrtc": 3}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_encode(icb, codec, fmt, buf, h, w, pitch=None, coded=None, strategy=2, tma=-1):
    prev = icb.set_tma_mode(tma)
    try:
        ch, cw = coded if coded else (h, w)
        out = icb.encode_device(codec, fmt, dev(buf), h, w, pitch=pitch, coded_h=max(h, ch), coded_w=max(w, cw), strategy=strategy)
        torch.cuda.synchronize()
        return out.cpu().numpy()
    finally:
        icb.set_tma_mode(prev)


def golden_codec(meta):
    if meta["codec"] == "dxt":
        return CODEC["dxt1"] if meta["ncomp"] == 3 else CODEC["dxt5"]
    return CODEC["etc1"] if meta["codec"] == "etc" else CODEC["pvrtc"]


def test_golden_device_path(icb, golden):
    """Every reference-generated fixture through the device entry points (generic driver: tiny, odd shapes)."""
    for meta, buf, want in golden:
        h, w = meta["h"], meta["w"]
        pitch = w * meta["ncomp"] + meta["padding"]
        got = gpu_encode(icb, golden_codec(meta), meta["format"], buf, h, w, pitch=pitch, coded=meta["padded"], strategy=meta["strategy"])
        assert np.array_equal(got, want), meta


def test_golden_host_path(icb, golden):
    """Same fixtures through icb_compress_host (host buffers, H2D/D2H inside the call)."""
    for meta, buf, want in golden:
        got = icb.compress_host(golden_codec(meta), meta["format"], np.ascontiguousarray(buf), meta["h"], meta["w"],
                                padded=meta["padded"], padding=meta["padding"], strategy=meta["strategy"])
        assert np.array_equal(got, want), meta


@pytest.mark.parametrize("fmt", [ck.RGB, ck.BGR, ck.RGBA, ck.BGRA])
@pytest.mark.parametrize("tma", [0, 1])
def test_dxt_vs_oracle(icb, fmt, tma):
    nc = ck.ncomp(fmt)
    codec = CODEC["dxt1"] if nc == 3 else CODEC["dxt5"]
    for kind in imagegen.KINDS:
        # widths are multiples of 4 with 16-byte aligned pitch so that tma=1 is legal; heights are ragged
        for (h, w) in ((64, 256), (37, 512), (130, 64), (16, 1028)):
            if tma == 1 and (w * nc) % 16 != 0:
                continue
            img = imagegen.make(kind, h, w, nc, seed=11)
            got = gpu_encode(icb, codec, fmt, img.ravel(), h, w, tma=tma)
            assert np.array_equal(got, ck.oracle_dxt(fmt, img.ravel(), h, w)), (kind, h, w)


@pytest.mark.parametrize("swap", [0, 1])
@pytest.mark.parametrize("tma", [0, 1])
def test_dxt1_from_rgba_extension(icb, swap, tma):
    fmt = ck.BGRA if swap else ck.RGBA
    for kind in imagegen.KINDS:
        for (h, w) in ((64, 256), (37, 516), (5, 8)):
            img = imagegen.make(kind, h, w, 4, seed=12)
            got = gpu_encode(icb, CODEC["dxt1"], fmt, img.ravel(), h, w, tma=tma)
            assert np.array_equal(got, ck.oracle_dxt1_rgba(img.ravel(), h, w, swap_rb=swap)), (kind, h, w)


@pytest.mark.parametrize("strategy", [0, 1, 2, 3])
@pytest.mark.parametrize("tma", [0, 1])
def test_etc1_vs_oracle(icb, strategy, tma):
    for kind in imagegen.KINDS:
        for (h, w) in ((32, 64), (21, 128), (64, 272)):
            img = imagegen.make(kind, h, w, 3, seed=13)
            got = gpu_encode(icb, CODEC["etc1"], ck.RGB, img.ravel(), h, w, strategy=strategy, tma=tma)
            assert np.array_equal(got, ck.oracle_etc1(strategy, img.ravel(), h, w)), (kind, h, w)


def test_ragged_sizes_and_padding(icb):
    """Widths/heights that are not multiples of 4, odd row padding (unaligned pitch), CompressAndPad grids."""
    for fmt in (ck.RGB, ck.BGRA):
        nc = ck.ncomp(fmt)
        codec = CODEC["dxt1"] if nc == 3 else CODEC["dxt5"]
        for (h, w, padding, coded) in ((1, 1, 0, None), (3, 3, 1, None), (9, 13, 7, None), (70, 258, 0, None),
                                       (70, 258, 0, (96, 300)), (8, 8, 0, (8, 40)), (8, 8, 0, (40, 8)), (33, 67, 5, (64, 128))):
            img = imagegen.make("smooth_noise", h, w, nc, seed=14)
            buf, pitch = imagegen.with_row_padding(img, padding)
            got = gpu_encode(icb, codec, fmt, buf, h, w, pitch=pitch, coded=coded)
            ch, cw = coded if coded else (None, None)
            assert np.array_equal(got, ck.oracle_dxt(fmt, buf, h, w, ch, cw, padding)), (fmt, h, w, padding, coded)
    img = imagegen.make("random", 33, 67, 3, seed=15)
    buf, pitch = imagegen.with_row_padding(img, 3)
    got = gpu_encode(icb, CODEC["etc1"], ck.RGB, buf, 33, 67, pitch=pitch, coded=(64, 128))
    assert np.array_equal(got, ck.oracle_etc1(2, buf, 33, 67, 64, 128, 3))


def test_tma_path_ragged_edges_aligned_pitch(icb):
    """TMA path with image edges that cut through blocks (clamp inside the tile) and a padded, aligned pitch."""
    for fmt, nc in ((ck.RGBA, 4), (ck.RGB, 3)):
        codec = CODEC["dxt5"] if nc == 4 else CODEC["dxt1"]
        for (h, w) in ((67, 1000), (258, 260), (19, 2052)):
            if nc == 3 and w % 4:
                continue
            img = imagegen.make("gradient", h, w, nc, seed=16)
            padding = (-w * nc) % 16 + 32
            buf, pitch = imagegen.with_row_padding(img, padding)
            got = gpu_encode(icb, codec, fmt, buf, h, w, pitch=pitch, tma=1)
            assert np.array_equal(got, ck.oracle_dxt(fmt, buf, h, w, None, None, padding)), (fmt, h, w)
    img = imagegen.make("random", 67, 1001, 4, seed=17)  # width not a multiple of 4, RGBA: TMA still legal
    buf, pitch = imagegen.with_row_padding(img, 12)
    got = gpu_encode(icb, CODEC["dxt5"], ck.RGBA, buf, 67, 1001, pitch=pitch, tma=1)
    assert np.array_equal(got, ck.oracle_dxt(ck.RGBA, buf, 67, 1001, None, None, 12))


def test_pvrtc_vs_oracle(icb):
    for s in (8, 16, 64, 256):
        for kind in imagegen.KINDS:
            img = imagegen.make(kind, s, s, 4, seed=s + 1)
            got = icb.pvrtc_encode_device(dev(img.ravel()), s, s).cpu().numpy()
            assert np.array_equal(got, ck.oracle_pvrtc(img.ravel(), s, s)), (s, kind)


def test_pvrtc_rejections(icb):
    img = dev(np.zeros(16 * 16 * 4, np.uint8))
    for (h, w) in ((8, 16), (4, 4), (12, 12)):
        with pytest.raises(icb.IcbError) as e:
            icb.pvrtc_encode_device(img, h, w)
        assert e.value.status == -4
    with pytest.raises(icb.IcbError) as e:
        icb.compress_host(icb.CODEC_PVRTC2, icb.RGBA, np.zeros(8 * 12 * 4, np.uint8), 8, 8, padding=4)
    assert e.value.status == -4


def test_host_path_size_mismatch(icb):
    src = np.zeros(16 * 16 * 4, np.uint8)
    with pytest.raises(icb.IcbError) as e:
        icb.compress_host(icb.CODEC_DXT5, icb.RGBA, src, 16, 16, out=np.zeros(100, np.uint8))
    assert e.value.status == -3


def test_synthetic_fill_matches_oracle_stream(icb):
    for (n, seed, off) in ((4096, 1, 0), (1000, 2, 13), (77, 3, 5), (65536 + 3, 2, 8 * 1000 + 1)):
        buf = torch.empty(n, dtype=torch.uint8, device="cuda")
        icb.fill_synthetic(buf, seed, off)
        assert np.array_equal(buf.cpu().numpy(), ck.synthetic(n, seed, off)), (n, seed, off)


def test_medium_synthetic_all_codecs(icb):
    """512^2 synthetic input (the bench's generator) through every codec, TMA path, against the oracle."""
    n = 512
    rgb = torch.empty(n * n * 3, dtype=torch.uint8, device="cuda")
    rgba = torch.empty(n * n * 4, dtype=torch.uint8, device="cuda")
    icb.fill_synthetic(rgb, 1)
    icb.fill_synthetic(rgba, 2)
    h_rgb, h_rgba = rgb.cpu().numpy(), rgba.cpu().numpy()
    assert np.array_equal(icb.encode_device(0, ck.RGB, rgb, n, n).cpu().numpy(), ck.oracle_dxt(ck.RGB, h_rgb, n, n))
    assert np.array_equal(icb.encode_device(0, ck.RGBA, rgba, n, n).cpu().numpy(), ck.oracle_dxt1_rgba(h_rgba, n, n))
    assert np.array_equal(icb.encode_device(1, ck.RGBA, rgba, n, n).cpu().numpy(), ck.oracle_dxt(ck.RGBA, h_rgba, n, n))
    assert np.array_equal(icb.encode_device(2, ck.RGB, rgb, n, n).cpu().numpy(), ck.oracle_etc1(2, h_rgb, n, n))
    assert np.array_equal(icb.pvrtc_encode_device(rgba, n, n).cpu().numpy(), ck.oracle_pvrtc(h_rgba, n, n))


def test_stripes_equal_whole_image(icb):
    """Row-stripe sharding: concatenated stripe outputs == single launch, for uneven splits."""
    h, w = 200, 512
    for codec, fmt in ((0, ck.RGBA), (1, ck.BGRA), (2, ck.RGB), (0, ck.RGB)):
        nc = ck.ncomp(fmt)
        img = dev(imagegen.make("smooth_noise", h, w, nc, seed=18).ravel())
        whole = icb.encode_device(codec, fmt, img, h, w).cpu().numpy()
        grid_rows = (h + 3) // 4
        bb = 16 if codec == 1 else 8
        for world in (2, 3, 8):
            parts = []
            for rank in range(world):
                r0, r1 = icb.stripe_rows(grid_rows, rank, world)
                out = torch.empty((r1 - r0) * (w // 4) * bb, dtype=torch.uint8, device="cuda")
                icb.encode_stripe_device(codec, fmt, img.data_ptr(), h, w, w * nc, h, w, r0, r1, out)
                parts.append(out.cpu().numpy())
            assert np.array_equal(np.concatenate(parts), whole), (codec, fmt, world)


FULL_SIZE = {  # BASELINE.json configs 1-4 (+ the API-faithful RGB888 DXT1): workload -> (codec, format, ncomp, side, seed)
    "dxt1_rgba8": (0, ck.RGBA, 4, 8192, 2), "dxt1_rgb8": (0, ck.RGB, 3, 8192, 1), "dxt5_rgba8": (1, ck.RGBA, 4, 8192, 2),
    "etc1_rgb8": (2, ck.RGB, 3, 4096, 1), "pvrtc2_rgba8": (3, ck.RGBA, 4, 4096, 2),
}


@pytest.mark.parametrize("workload", sorted(FULL_SIZE))
def test_full_size_byte_compare_with_cpu_reference(icb, workload):
    """Every benchmarked configuration at its FULL size, all output bytes: device path (the exact call bench.py times)
    and host path against the reference's own CPU encoder (oracle/_ref, unmodified, multi-threaded over row stripes;
    the oracle port where _ref was not built).  Input = the bench's synthetic stream S(seed)."""
    codec, fmt, nc, n, seed = FULL_SIZE[workload]
    src = torch.empty(n * n * nc, dtype=torch.uint8, device="cuda")
    icb.fill_synthetic(src, seed)
    host = src.cpu().numpy()
    want, kind = ck.cpu_encode_full(workload, host, n, n)
    if codec == 3:
        got = icb.pvrtc_encode_device(src, n, n)
    else:
        got = icb.encode_device(codec, fmt, src, n, n)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert got.size == want.size
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, "%s: %d bytes differ from the CPU %s, first at byte %d" % (workload, bad.size, kind, bad[0])
    via_host = icb.compress_host(codec, fmt, host, n, n)
    assert np.array_equal(via_host, want), "%s: icb_compress_host differs from the CPU %s" % (workload, kind)


def test_etc1_line_and_dot_forms_on_full_warps(icb):
    """ETC1 on 2048 x 1024 images whose warps (32 neighbouring blocks of a block row) take EVERY split between the line
    form (codewords that cannot clamp, a prefix decided by the warp's smallest margin to black and white) and the dot
    form: horizontal bands whose mean level sweeps the whole range, smooth structure across them, noise of growing
    amplitude, and single saturated pixels that pull one warp's margin down (per warp 0 .. 6 of the eight codewords on
    the line form).  Every byte against the CPU reference; the one-orientation strategies against the oracle port."""
    h, w = 1024, 2048
    rng = np.random.default_rng(41)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    level = 8 + (yy // 16) * (240.0 / (h // 16 - 1))                         # band mean 8 .. 248
    wave = 40 * np.sin(xx / 37.0 + yy / 91.0)[..., None] * np.array([1.0, 0.6, -0.8], np.float32)
    amp = (1 + (xx // 256) * 6)[..., None]                                   # noise +-1 .. +-43 by column band
    img = level[..., None] + wave + rng.integers(-1, 2, (h, w, 3)) * amp
    spikes = rng.random((h // 4, w // 4)) < 0.01                             # 1 % of the blocks: one saturated pixel
    sy, sx = np.nonzero(spikes)
    img[sy * 4 + 1, sx * 4 + 2] = rng.choice(np.array([0.0, 255.0]), (sy.size, 1))
    host = np.ascontiguousarray(np.clip(img, 0, 255).astype(np.uint8))
    want, kind = ck.cpu_encode_full("etc1_rgb8", host, h, w)
    src = torch.from_numpy(host.ravel()).cuda()
    got = icb.encode_device(2, ck.RGB, src, h, w).cpu().numpy()
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, "%d bytes differ from the CPU %s, first in block %d" % (bad.size, kind, bad[0] // 8)
    for strategy in (0, 1):  # one orientation only: the same search, against the oracle port on a slice
        rows = 64
        sl = np.ascontiguousarray(host[:rows])
        part = icb.encode_device(2, ck.RGB, torch.from_numpy(sl.ravel()).cuda(), rows, w, strategy=strategy).cpu().numpy()
        assert np.array_equal(part, ck.oracle_etc1(strategy, sl.ravel(), rows, w)), strategy


def test_full_size_structured_content_dxt(icb):
    """8192^2 structured content (what real textures look like: flat regions, gradients, dark areas, 2-colour cells),
    DXT1 and DXT5, every byte against the CPU reference -- the general (non-fast-path) index search, the constant-
    colour table and the 6-alpha mode at scale, which uniform random bytes almost never reach."""
    n = 8192
    yy, xx = torch.meshgrid(torch.arange(n, device="cuda", dtype=torch.int32), torch.arange(n, device="cuda", dtype=torch.int32), indexing="ij")
    noise = torch.empty(n * n * 4, dtype=torch.uint8, device="cuda")
    icb.fill_synthetic(noise, 77)
    noise = noise.view(n, n, 4).int()
    r = torch.where((xx // 512 + yy // 512) % 4 == 0, (xx // 32) % 256, torch.where((xx // 512 + yy // 512) % 4 == 1, noise[..., 0] // 16, (xx // 8 ^ yy // 8) % 2 * 255))
    g = torch.where((xx // 512 + yy // 512) % 4 == 2, torch.full_like(xx, 150), (yy // 16) % 256)
    b = torch.where((xx // 1024) % 2 == 0, ((xx + yy) // 64) % 256, noise[..., 2] % 3)
    a = torch.where((yy // 256) % 3 == 0, torch.full_like(xx, 255), torch.where((yy // 256) % 3 == 1, noise[..., 3] // 128 * 255, (xx // 4) % 256))
    img = torch.stack([r, g, b, a], -1).to(torch.uint8).contiguous()
    del noise, r, g, b, a, xx, yy
    host = img.cpu().numpy()
    for workload, codec in (("dxt1_rgba8", 0), ("dxt5_rgba8", 1)):
        want, kind = ck.cpu_encode_full(workload, host, n, n)
        got = icb.encode_device(codec, ck.RGBA, img.view(-1), n, n).cpu().numpy()
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, "%s structured: %d bytes differ from the CPU %s, first at byte %d" % (workload, bad.size, kind, bad[0])


def test_full_size_properties(icb):
    """BASELINE.json full sizes, where the scalar oracle would take minutes: size-independent properties.
    (a) TMA path == generic path byte for byte; (b) block-row translation invariance: encoding rows [k, k+256)
    as its own image gives the same bytes as that slice of the full output; (c) a random sample of blocks agrees
    with the oracle applied to just those 4x4 windows."""
    n = 8192
    rgba = torch.empty(n * n * 4, dtype=torch.uint8, device="cuda")
    icb.fill_synthetic(rgba, 2)
    for codec, bb in ((0, 8), (1, 16)):
        icb.set_tma_mode(1)
        fast = icb.encode_device(codec, ck.RGBA, rgba, n, n)
        icb.set_tma_mode(0)
        slow = icb.encode_device(codec, ck.RGBA, rgba, n, n)
        icb.set_tma_mode(-1)
        assert torch.equal(fast, slow)
        # (b) rows 4096..5119 as a standalone 1024 x 8192 image
        sub = rgba[4096 * n * 4:(4096 + 1024) * n * 4]
        sub_out = icb.encode_device(codec, ck.RGBA, sub, 1024, n)
        assert torch.equal(sub_out, fast[1024 * (n // 4) * bb:(1024 + 256) * (n // 4) * bb])
        # (c) 4096 sampled blocks against the oracle
        rng = np.random.default_rng(3)
        brs, bcs = rng.integers(0, n // 4, 4096), rng.integers(0, n // 4, 4096)
        img = rgba.view(n, n, 4)
        idx_r = torch.from_numpy(brs).cuda()[:, None] * 4 + torch.arange(4, device="cuda")[None, :]
        idx_c = torch.from_numpy(bcs).cuda()[:, None] * 4 + torch.arange(4, device="cuda")[None, :]
        windows = img[idx_r[:, :, None], idx_c[:, None, :]].cpu().numpy()  # [4096, 4, 4, 4]
        strip = np.ascontiguousarray(windows.transpose(1, 0, 2, 3).reshape(4, 4096 * 4, 4))  # one row of blocks
        want = ck.oracle_dxt1_rgba(strip.ravel(), 4, 4096 * 4) if codec == 0 else ck.oracle_dxt(ck.RGBA, strip.ravel(), 4, 4096 * 4)
        got = fast.view(n // 4, n // 4, bb)[torch.from_numpy(brs).cuda(), torch.from_numpy(bcs).cuda()].cpu().numpy()
        assert np.array_equal(got.ravel(), want)


def test_decoders_vs_oracle(icb):
    """Block decoders on arbitrary bit patterns (3-colour DXT1 blocks, equal endpoints, ETC1 differential overflow),
    ragged sizes, both channel orders; device and host entry points."""
    rng = np.random.default_rng(41)
    for codec, bb in ((0, 8), (1, 16), (2, 8)):
        for (h, w) in ((4, 4), (16, 16), (5, 7), (33, 18), (64, 256), (1, 1)):
            nb = ck.nblocks(h) * ck.nblocks(w)
            for kind in range(3):
                blocks = rng.integers(0, 256, nb * bb, dtype=np.uint8)
                if kind == 1 and codec != 2:  # c0 == c1 in every block
                    v = blocks.reshape(-1, bb)
                    v[:, bb - 6:bb - 4] = v[:, bb - 8:bb - 6]
                if kind == 2:  # blocks an encoder really produces
                    fmt = ck.RGBA if codec == 1 else ck.RGB
                    img = imagegen.make("smooth_noise", h, w, ck.ncomp(fmt), seed=42).ravel()
                    blocks = ck.oracle_etc1(2, img, h, w) if codec == 2 else ck.oracle_dxt(fmt, img, h, w)
                for swap in ((0, 1) if codec != 2 else (0,)):
                    want = ck.oracle_decode(codec, blocks, h, w, swap_rb=swap)
                    got = icb.decode_device(codec, dev(blocks), h, w, swap_rb=swap).cpu().numpy()
                    assert np.array_equal(got, want), (codec, h, w, kind, swap)
                    fmt = (ck.BGRA if swap else ck.RGBA) if codec == 1 else (ck.BGR if swap else ck.RGB)
                    assert np.array_equal(icb.decompress_host(codec, fmt, blocks, h, w), want)


def test_encode_decode_round_trip_full_size(icb):
    """Full-size property: decode(encode(x)) stays close to x (DXT5 on smooth 4096^2 content), all on the device."""
    n = 4096
    yy, xx = torch.meshgrid(torch.arange(n, device="cuda"), torch.arange(n, device="cuda"), indexing="ij")
    img = torch.stack([(xx // 16) % 256, (yy // 16) % 256, ((xx + yy) // 32) % 256, 255 - (xx // 16) % 256], -1).to(torch.uint8).contiguous()
    blocks = icb.encode_device(1, ck.RGBA, img.view(-1), n, n)
    back = icb.decode_device(1, blocks, n, n).view(n, n, 4)
    err = (back.int() - img.int()).abs()
    assert int(err.max()) <= 24 and float(err.float().mean()) < 3.0


def test_tma_path_small_wide_and_odd_stripes(icb):
    """Forced TMA path on shapes smaller than a tile, much wider than a tile row, and stripes that start in the
    middle of a tile row."""
    for (h, w) in ((4, 4), (4, 16), (16, 16), (8, 4096), (12, 16388), (100, 260)):
        img = imagegen.make("random", h, w, 4, seed=51)
        for codec, want in ((0, ck.oracle_dxt1_rgba(img.ravel(), h, w)), (1, ck.oracle_dxt(ck.RGBA, img.ravel(), h, w))):
            got = gpu_encode(icb, codec, ck.RGBA, img.ravel(), h, w, tma=1)
            assert np.array_equal(got, want), (codec, h, w)
    h, w = 160, 512
    img = dev(imagegen.make("smooth_noise", h, w, 4, seed=52).ravel())
    whole = icb.encode_device(1, ck.RGBA, img, h, w).cpu().numpy()
    prev = icb.set_tma_mode(1)
    try:
        for (r0, r1) in ((0, 1), (1, 7), (7, 40), (3, 39), (39, 40)):
            out = torch.empty((r1 - r0) * (w // 4) * 16, dtype=torch.uint8, device="cuda")
            icb.encode_stripe_device(1, ck.RGBA, img.data_ptr(), h, w, w * 4, h, w, r0, r1, out)
            assert np.array_equal(out.cpu().numpy(), whole[r0 * (w // 4) * 16:r1 * (w // 4) * 16]), (r0, r1)
    finally:
        icb.set_tma_mode(prev)


def test_large_row_padding(icb):
    """Row pitch far larger than the row (aligned and unaligned), device and host paths."""
    for padding in (4096, 4099):
        for fmt, codec in ((ck.RGBA, 1), (ck.RGB, 0), (ck.RGB, 2)):
            nc = ck.ncomp(fmt)
            img = imagegen.make("gradient", 24, 40, nc, seed=53)
            buf, pitch = imagegen.with_row_padding(img, padding)
            want = ck.oracle_etc1(2, buf, 24, 40, None, None, padding) if codec == 2 else ck.oracle_dxt(fmt, buf, 24, 40, None, None, padding)
            assert np.array_equal(gpu_encode(icb, codec, fmt, buf, 24, 40, pitch=pitch), want), (padding, fmt, codec)
            assert np.array_equal(icb.compress_host(codec, fmt, buf, 24, 40, padding=padding), want), (padding, fmt, codec)


def test_host_path_is_thread_safe(icb):
    """Concurrent Compress calls from several host threads (the reference is re-entrant; so is this build)."""
    import threading
    jobs = []
    for i, (codec, fmt) in enumerate(((0, ck.RGB), (1, ck.RGBA), (2, ck.RGB), (3, ck.RGBA), (0, ck.BGR), (1, ck.BGRA))):
        n = 64 if codec == 3 else 72 + 4 * i
        img = imagegen.make("smooth_noise", n, n, ck.ncomp(fmt), seed=60 + i)
        if codec == 3:
            want = ck.oracle_pvrtc(img.ravel(), n, n)
        elif codec == 2:
            want = ck.oracle_etc1(2, img.ravel(), n, n)
        else:
            want = ck.oracle_dxt(fmt, img.ravel(), n, n)
        jobs.append((codec, fmt, img, n, want))
    failures = []

    def worker(job):
        codec, fmt, img, n, want = job
        for _ in range(20):
            got = icb.compress_host(codec, fmt, img.ravel(), n, n)
            if not np.array_equal(got, want):
                failures.append((codec, fmt))
                return

    threads = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not failures


_DRIVER_SNIPPET = r"""
import sys
import numpy as np, torch
sys.path.insert(0, "tests")
import checkers as ck, imagegen
import image_compression_b200 as icb
for kind in ("random", "smooth_noise", "two_colour"):
    for (h, w) in ((64, 256), (40, 516), (260, 1024), (72, 1280)):
        rgba = imagegen.make(kind, h, w, 4, seed=7)
        rgb = np.ascontiguousarray(rgba[..., :3])
        cases = [(0, ck.RGBA, rgba, ck.oracle_dxt1_rgba(rgba.ravel(), h, w)), (1, ck.RGBA, rgba, ck.oracle_dxt(ck.RGBA, rgba.ravel(), h, w)),
                 (0, ck.RGB, rgb, ck.oracle_dxt(ck.RGB, rgb.ravel(), h, w)), (2, ck.RGB, rgb, ck.oracle_etc1(2, rgb.ravel(), h, w))]
        for codec, fmt, img, want in cases:
            got = icb.encode_device(codec, fmt, torch.from_numpy(img.ravel().copy()).cuda(), h, w).cpu().numpy()
            assert np.array_equal(got, want), (kind, h, w, codec, fmt)
print("driver ok")
"""


@pytest.mark.parametrize("driver,stages", [("ring", "2"), ("ring", "3"), ("producer", "3"), ("producer", "4")])
def test_both_ring_drivers_every_codec(icb, driver, stages):
    """The launcher picks one ring-refill scheme per codec (producer warp for DXT1/ETC1, counted release for DXT5);
    ICB_DRIVER / ICB_TMA_STAGES force the other combinations, which must stay bit-exact too.  The choice is cached
    per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ICB_DRIVER=driver, ICB_TMA_STAGES=stages)
    res = subprocess.run([sys.executable, "-c", _DRIVER_SNIPPET], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "driver ok" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


def _pvrtc_stripe_rows(img, h, w, r0, r1):
    """Image rows 4*(r0-1) .. 4*(r1+1)-1, wrapped: what icb_pvrtc2_encode_stripe wants resident."""
    from image_compression_b200 import sharding
    return np.ascontiguousarray(img.reshape(h, w * 4)[sharding.pvrtc_stripe_row_indices(h, r0, r1)]).ravel()


@pytest.mark.parametrize("n,parts", [(64, 2), (64, 4), (256, 2), (256, 8), (512, 3), (32, 2)])
def test_pvrtc_stripes_assemble_whole_image(icb, n, parts):
    """Every rank of a sharded PVRTC encode holds only its block rows plus one halo block row each side (wrapped) and
    pixel (0,0); each stripe's blocks land at their Z-order slots of the whole-image buffer.  The union must equal the
    reference (oracle) encoding of the whole image -- including quirk P1 blocks, which read pixel (0,0)."""
    from image_compression_b200 import sharding
    for kind in ("random", "zero_channel", "alpha_extremes", "smooth_noise"):
        img = imagegen.make(kind, n, n, 4, seed=61).ravel()
        want = ck.oracle_pvrtc(img, n, n)
        out = torch.zeros(n * n // 4, dtype=torch.uint8, device="cuda")
        first = dev(img[:4])
        lh = n // 4
        for r in range(parts):
            r0, r1 = sharding.stripe_rows(lh, r, parts)
            if r1 - r0 + 2 > lh:
                pytest.skip("stripe + halo larger than the image")
            icb.pvrtc_encode_stripe_device(dev(_pvrtc_stripe_rows(img, n, n, r0, r1)), first, n, n, r0, r1, out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want), (kind, n, parts)


def test_pvrtc_stripe_rejections(icb):
    rows = torch.zeros(64 * 4 * 4 * 6, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64 * 64 // 4, dtype=torch.uint8, device="cuda")
    first = torch.zeros(4, dtype=torch.uint8, device="cuda")
    for (h, w, r0, r1) in ((64, 64, 0, 16), (64, 64, 3, 3), (64, 64, 5, 17), (48, 48, 0, 2), (64, 32, 0, 2)):
        with pytest.raises(icb.IcbError):
            icb.pvrtc_encode_stripe_device(rows, first, h, w, r0, r1, out)


def test_host_pipes_are_pooled_not_leaked_per_thread(icb):
    """the host path used to keep its streams, events, pinned rings and image-sized device buffers in
    thread_local storage for ever.  Now they are leased from a pool: 24 short-lived threads (at most 3 at a time)
    leave at most 3 pipes behind, free device memory stops shrinking after the first wave, and icb_trim() gives the
    memory back."""
    import threading
    L = icb.lib()
    L.icb_trim()
    h = w = 1024
    img = ck.synthetic(h * w * 4, 15)
    want = ck.oracle_dxt(ck.RGBA, img, h, w)
    failures = []

    def request():
        got = icb.compress_host(icb.CODEC_DXT5, icb.RGBA, img, h, w)
        if not np.array_equal(got, want):
            failures.append(1)

    free_after_wave = []
    for wave in range(8):
        ts = [threading.Thread(target=request) for _ in range(3)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        torch.cuda.synchronize()
        free_after_wave.append(torch.cuda.mem_get_info()[0])
    assert not failures
    assert 1 <= L.icb_idle_pipes() <= 3
    assert free_after_wave[-1] >= free_after_wave[1] - (8 << 20), free_after_wave   # no growth after warm-up
    released = L.icb_trim()
    assert released >= h * w * 4 and L.icb_idle_pipes() == 0
    assert np.array_equal(icb.compress_host(icb.CODEC_DXT5, icb.RGBA, img, h, w), want)   # and it comes back


def test_failed_host_call_leaves_the_pipe_usable(icb):
    """An error after work has been queued (a destination-size check passes, then a refused block operation) must not
    poison the pooled pipe for the next caller."""
    img = ck.synthetic(64 * 64 * 3, 16)
    blocks = icb.compress_host(icb.CODEC_DXT1, icb.RGB, img, 64, 64)
    with pytest.raises(icb.IcbError):
        icb.blockop_host(icb.OP_PAD, icb.CODEC_DXT1, [64, 64, 32, 128], blocks, 8 * 32 * 8)   # shrinks one way, grows the other
    assert np.array_equal(icb.compress_host(icb.CODEC_DXT1, icb.RGB, img, 64, 64), ck.oracle_dxt(ck.RGB, img, 64, 64))


@pytest.mark.parametrize("codec,fmt", [(0, ck.RGBA), (0, ck.RGB), (0, ck.BGR), (1, ck.BGRA), (2, ck.RGB)])
def test_unaligned_device_sources(icb, codec, fmt):
    """Device-resident images TMA cannot describe -- base pointer off by 1..15 bytes, row pitch not a multiple of 16 (or
    of 4), RGB888 widths that are not multiples of 16 pixels -- are encoded by the generic kernel, whatever the
    alignment: same bytes as the oracle."""
    nc = ck.ncomp(fmt)
    launches = icb.launch_count
    for (h, w, padding, offset) in ((64, 512, 0, 1), (67, 300, 0, 3), (40, 1028, 5, 2), (130, 260, 7, 13), (16, 256, 1, 4), (96, 2052, 0, 8)):
        img = imagegen.make("smooth_noise" if (h + w) & 1 else "random", h, w, nc, seed=h + w)
        buf, pitch = imagegen.with_row_padding(img, padding)
        if codec == 2:
            want = ck.oracle_etc1(2, buf, h, w, padding=padding)
        elif codec == 0 and nc == 4:
            want = ck.oracle_dxt1_rgba(buf, h, w, swap_rb=1 if fmt == ck.BGRA else 0, padding=padding)
        else:
            want = ck.oracle_dxt(fmt, buf, h, w, padding=padding)
        big = torch.zeros(buf.size + 64, dtype=torch.uint8, device="cuda")
        big[offset:offset + buf.size] = torch.from_numpy(np.ascontiguousarray(buf).ravel()).cuda()
        src = big[offset:offset + buf.size]
        assert src.data_ptr() % 16 != 0 or pitch % 16 != 0
        for mode in (-1, 0):
            prev = icb.set_tma_mode(mode)
            try:
                before = launches()
                got = icb.encode_device(codec, fmt, src, h, w, pitch=pitch)
                torch.cuda.synchronize()
            finally:
                icb.set_tma_mode(prev)
            assert np.array_equal(got.cpu().numpy(), want), (codec, fmt, h, w, padding, offset, mode)
            if mode != 0 and h >= 16 and w >= 256:
                assert launches() - before >= 1


def test_unaligned_full_size_rgb888_width_not_multiple_of_16(icb):
    """A large RGB888 image whose pitch is not 16-byte aligned (8184 px wide: 24552 bytes per row), whole image through
    the generic kernel, every byte against the CPU reference run on row stripes."""
    h, w = 4096, 8184
    src = torch.empty(h * w * 3, dtype=torch.uint8, device="cuda")
    icb.fill_synthetic(src, 21)
    host = src.cpu().numpy()
    want, kind = ck.cpu_encode_full("dxt1_rgb8", host, h, w)
    got = icb.encode_device(0, ck.RGB, src, h, w).cpu().numpy()
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, "%d bytes differ from the CPU %s, first at %d" % (bad.size, kind, bad[0])


def test_registered_caller_buffers_take_the_pinned_path(icb):
    """icb_host_register: a buffer the caller already owns (numpy memory here) becomes page-locked; the host path then
    DMAs it directly instead of staging it, with the same bytes out; unregistering puts it back on the staged path."""
    L = icb.lib()
    h = w = 2048
    img = ck.synthetic(h * w * 4, 31)
    out = np.zeros(icb.compressed_size(icb.CODEC_DXT5, h, w), np.uint8)
    want = ck.oracle_dxt(ck.RGBA, img, h, w)
    assert L.icb_host_register(img.ctypes.data, img.size) == 0 and L.icb_host_register(out.ctypes.data, out.size) == 0
    try:
        icb.compress_host(icb.CODEC_DXT5, icb.RGBA, img, h, w, out=out)
        assert np.array_equal(out, want)
    finally:
        assert L.icb_host_unregister(img.ctypes.data) == 0 and L.icb_host_unregister(out.ctypes.data) == 0
    out[:] = 0
    icb.compress_host(icb.CODEC_DXT5, icb.RGBA, img, h, w, out=out)
    assert np.array_equal(out, want)
    assert L.icb_host_register(None, 16) == -1


def test_pvrtc_fused_kernel_equals_three_kernel_pipeline_and_oracle(icb):
    """ICB_PVRTC_FUSED=1 runs Morph + the fused Modulate/Pack kernel on whole images >= 256 x 256 (opt-in: it measured
    slower than the three-kernel pipeline).  Both must give the oracle's bytes (256 .. 1024, every content kind) and
    each other's at 2048."""
    import os

    def fused(fn):
        os.environ["ICB_PVRTC_FUSED"] = "1"
        try:
            return fn()
        finally:
            del os.environ["ICB_PVRTC_FUSED"]

    for n in (256, 512, 1024):
        for kind in imagegen.KINDS if n < 1024 else ("random", "alpha_extremes"):
            img = imagegen.make(kind, n, n, 4, seed=n + 3)
            want = ck.oracle_pvrtc(img.ravel(), n, n)
            d = dev(img.ravel())
            before = icb.launch_count()
            got3 = icb.pvrtc_encode_device(d, n, n).cpu().numpy()
            assert icb.launch_count() - before == 3          # Morph, Modulate, Pack
            assert np.array_equal(got3, want), (n, kind, "three kernels")
            before = icb.launch_count()
            got2 = fused(lambda: icb.pvrtc_encode_device(d, n, n).cpu().numpy())
            assert icb.launch_count() - before == 2          # Morph, Modulate+Pack
            assert np.array_equal(got2, want), (n, kind, "fused")
    n = 2048
    src = torch.empty(n * n * 4, dtype=torch.uint8, device="cuda")
    icb.fill_synthetic(src, 12)
    a = icb.pvrtc_encode_device(src, n, n)
    b = fused(lambda: icb.pvrtc_encode_device(src, n, n))
    assert torch.equal(a, b)
