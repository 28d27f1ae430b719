"""GPU parity tests of the compressed-domain operations: Downsample, Pad,
CopySubimage, CreateSolidImage and the DXT1 -> ETC1 transcoder, called through the C ABI (device-resident and
host-buffer forms) and through the C++ classes, against the CPU oracle and the reference-generated fixtures.
Bit-exact: every comparison is array_equal."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import checkers as ck
import imagegen

pytestmark = pytest.mark.gpu

CODEC_FORMATS = ((0, ck.RGB), (0, ck.BGR), (1, ck.RGBA), (1, ck.BGRA), (2, ck.RGB))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def compress(codec, fmt, img, h, w, strategy=2):
    return ck.oracle_etc1(strategy, img.ravel(), h, w) if codec == 2 else ck.oracle_dxt(fmt, img.ravel(), h, w)


def test_golden_ops_device_and_host(icb, golden_ops):
    for meta, src, want in golden_ops:
        op, codec = meta["op"], meta["codec"]
        src = np.ascontiguousarray(src)
        st = meta.get("strategy", 2)
        if op == "downsample":
            h, w = meta["h"], meta["w"]
            if meta["refused"]:
                with pytest.raises(icb.IcbError):
                    icb.downsample_device(codec, dev(src), h, w, strategy=st)
                continue
            assert np.array_equal(host(icb.downsample_device(codec, dev(src), h, w, strategy=st)), want), meta
            assert np.array_equal(icb.blockop_host(icb.OP_DOWNSAMPLE, codec, [h, w], src, want.size, strategy=st), want), meta
        elif op == "pad":
            ch, cw = 4 * ck.nblocks(meta["h"]), 4 * ck.nblocks(meta["w"])
            assert np.array_equal(host(icb.pad_device(codec, dev(src), ch, cw, meta["ph"], meta["pw"], strategy=st)), want), meta
            assert np.array_equal(icb.blockop_host(icb.OP_PAD, codec, [ch, cw, meta["ph"], meta["pw"]], src, want.size, strategy=st), want), meta
        elif op == "copy_subimage":
            args = (4 * ck.nblocks(meta["h"]), 4 * ck.nblocks(meta["w"]), meta["row"], meta["col"], meta["sh"], meta["sw"])
            if meta["refused"]:
                with pytest.raises(icb.IcbError):
                    icb.copy_subimage_device(codec, dev(src), *args)
                continue
            assert np.array_equal(host(icb.copy_subimage_device(codec, dev(src), *args)), want), meta
            assert np.array_equal(icb.blockop_host(icb.OP_COPY_SUBIMAGE, codec, list(args), src, want.size), want), meta
        elif op == "solid":
            assert np.array_equal(host(icb.fill_solid_device(codec, src.tolist(), meta["h"], meta["w"])), want), meta
            packed = int(src[0]) | int(src[1]) << 8 | int(src[2]) << 16 | int(src[3]) << 24
            assert np.array_equal(icb.blockop_host(icb.OP_SOLID, codec, [meta["h"], meta["w"], packed], None, want.size), want), meta
        else:
            assert np.array_equal(host(icb.transcode_dxt1_to_etc1_device(dev(src))), want), meta
            assert np.array_equal(icb.blockop_host(icb.OP_TRANSCODE, 0, [], src, want.size), want), meta


@pytest.mark.parametrize("codec,fmt", CODEC_FORMATS)
def test_downsample_vs_oracle(icb, codec, fmt):
    sizes = ((8, 8), (16, 24), (32, 8), (13, 29), (5, 7), (4, 16), (24, 4), (4, 4), (2, 2), (1, 1), (1, 4), (4, 2), (2, 1),
             (64, 64), (256, 512), (120, 1000))
    for kind in ("random", "smooth_noise", "constant", "two_colour", "alpha_extremes", "dark", "zero_channel"):
        for (h, w) in sizes:
            if h * w > 4096 and kind not in ("random", "smooth_noise"):
                continue
            for st in ((0, 1, 2, 3) if codec == 2 and h * w <= 4096 else (2,)):
                blocks = compress(codec, fmt, imagegen.make(kind, h, w, ck.ncomp(fmt), seed=31), h, w, st)
                got = host(icb.downsample_device(codec, dev(blocks), h, w, strategy=st))
                assert np.array_equal(got, ck.oracle_downsample(codec, blocks, h, w, strategy=st)), (kind, h, w, st)


@pytest.mark.parametrize("codec,fmt", CODEC_FORMATS)
def test_downsample_arbitrary_blocks_and_refusals(icb, codec, fmt):
    rng = np.random.default_rng(41 + codec)
    for (h, w) in ((16, 16), (8, 32), (4, 8), (4, 4), (64, 128)):
        blocks = rng.integers(0, 256, ck.nblocks(h) * ck.nblocks(w) * ck.block_bytes(codec), dtype=np.uint8)
        got = host(icb.downsample_device(codec, dev(blocks), h, w))
        assert np.array_equal(got, ck.oracle_downsample(codec, blocks, h, w)), (h, w)
    blocks = dev(rng.integers(0, 256, 3 * 3 * 16, dtype=np.uint8))
    for (h, w) in ((12, 8), (8, 12), (3, 4), (4, 3), (9, 9)):  # odd block counts; 3-pixel single block
        assert ck.oracle_downsample(codec, host(blocks), h, w) is None
        with pytest.raises(icb.IcbError) as e:
            icb.downsample_device(codec, blocks, h, w)
        assert e.value.status == -4


def test_mip_chain_stays_on_device(icb):
    """Compress once, then halve down to one block without leaving HBM; every level equals the oracle's."""
    n = 256
    for codec, fmt in CODEC_FORMATS:
        img = imagegen.make("smooth_noise", n, n, ck.ncomp(fmt), seed=5)
        cid = {0: icb.CODEC_DXT1, 1: icb.CODEC_DXT5, 2: icb.CODEC_ETC1}[codec]
        level = icb.encode_device(cid, fmt, dev(img.ravel()), n, n)
        want = compress(codec, fmt, img, n, n)
        assert np.array_equal(host(level), want)
        size = n
        while size > 1:
            level = icb.downsample_device(codec, level, size, size)
            want = ck.oracle_downsample(codec, want, size, size)
            size = (size + 1) // 2
            assert np.array_equal(host(level), want), (codec, fmt, size)
            if size == 3:
                break


@pytest.mark.parametrize("codec,fmt", CODEC_FORMATS)
def test_pad_vs_oracle(icb, codec, fmt):
    rng = np.random.default_rng(51 + codec)
    for (h, w) in ((8, 8), (5, 7), (16, 4), (4, 4), (12, 20), (64, 96)):
        ch, cw = 4 * ck.nblocks(h), 4 * ck.nblocks(w)
        for (ph, pw) in ((h + 9, w + 6), (h, w + 8), (h + 4, w), (h, w), (4, 4), (h + 1, w + 1), (128, 128), (ch, cw + 1)):
            if ck.nblocks(ph) < ck.nblocks(ch) and pw > cw or ck.nblocks(pw) < ck.nblocks(cw) and ph > ch:
                continue
            for content in ("image", "random"):
                if content == "image":
                    blocks = compress(codec, fmt, imagegen.make("smooth_noise", h, w, ck.ncomp(fmt), 1), h, w)
                else:
                    blocks = rng.integers(0, 256, ck.nblocks(h) * ck.nblocks(w) * ck.block_bytes(codec), dtype=np.uint8)
                for st in ((2, 3) if codec == 2 else (2,)):
                    got = host(icb.pad_device(codec, dev(blocks), ch, cw, ph, pw, strategy=st))
                    assert np.array_equal(got, ck.oracle_pad(codec, blocks, ch, cw, ph, pw, strategy=st)), (h, w, ph, pw, content, st)
    with pytest.raises(icb.IcbError) as e:  # the reference overruns its buffer here; this build refuses
        icb.pad_device(codec, dev(np.zeros(4 * 4 * 16, np.uint8)), 16, 16, 8, 32)
    assert e.value.status == -4


def test_solid_and_transcode_vs_oracle(icb):
    rng = np.random.default_rng(61)
    for codec, _ in CODEC_FORMATS:
        for (h, w) in ((4, 4), (9, 6), (1, 1), (256, 260)):
            colour = rng.integers(0, 256, 4, dtype=np.uint8)
            want = np.tile(ck.oracle_solid_block(codec, colour), ck.nblocks(h) * ck.nblocks(w))
            assert np.array_equal(host(icb.fill_solid_device(codec, colour.tolist(), h, w)), want), (codec, h, w)
    for kind in imagegen.KINDS:
        img = imagegen.make(kind, 64, 96, 3, seed=2)
        blocks = ck.oracle_dxt(ck.RGB, img.ravel(), 64, 96)
        assert np.array_equal(host(icb.transcode_dxt1_to_etc1_device(dev(blocks))), ck.oracle_transcode(blocks)), kind
    blocks = rng.integers(0, 256, 8 * 4099, dtype=np.uint8)
    assert np.array_equal(host(icb.transcode_dxt1_to_etc1_device(dev(blocks))), ck.oracle_transcode(blocks))


def test_large_downsample_and_transcode(icb):
    """2048 x 2048 (262,144 input blocks): whole-image comparison with the oracle, DXT1 and DXT5; transcode of the
    DXT1 stream; and a size-independent property of solid images."""
    n = 2048
    rgba = torch.empty(n * n * 4, dtype=torch.uint8, device="cuda")
    icb.fill_synthetic(rgba, 2)
    h_rgba = host(rgba)
    d5 = icb.encode_device(icb.CODEC_DXT5, icb.RGBA, rgba, n, n)
    want5 = ck.oracle_dxt(ck.RGBA, h_rgba, n, n)
    assert np.array_equal(host(d5), want5)
    assert np.array_equal(host(icb.downsample_device(1, d5, n, n)), ck.oracle_downsample(1, want5, n, n))
    d1 = icb.encode_device(icb.CODEC_DXT1, icb.RGBA, rgba, n, n)
    want1 = ck.oracle_dxt1_rgba(h_rgba, n, n)
    assert np.array_equal(host(icb.downsample_device(0, d1, n, n)), ck.oracle_downsample(0, want1, n, n))
    assert np.array_equal(host(icb.transcode_dxt1_to_etc1_device(d1.clone())), ck.oracle_transcode(want1))
    for codec in (0, 1, 2):
        # size-independent property: a solid image downsamples to a tiling of ONE block, the one the oracle gets
        # from an 8 x 8 solid image
        solid = icb.fill_solid_device(codec, [40, 90, 200, 77], n, n)
        half = host(icb.downsample_device(codec, solid, n, n))
        bb = ck.block_bytes(codec)
        small = np.tile(ck.oracle_solid_block(codec, (40, 90, 200, 77)), 4)
        one = ck.oracle_downsample(codec, small, 8, 8)
        assert one.size == bb and np.array_equal(half.reshape(-1, bb), np.tile(one, (half.size // bb, 1)))


# ---- through the C++ classes ----------------------------------------------------------------------------------

_u8p = C.POINTER(C.c_uint8)


@pytest.fixture(scope="module", params=["libicb_api_test.so", "libicb_refabi_test.so"])
def api(request):
    # second build: the same doorway compiled against the reference's own headers (see tests/test_cpp_api.py)
    path = os.path.join(ROOT, "image_compression_b200", "lib", request.param)
    if not os.path.exists(path) and "refabi" in request.param and not os.path.isdir("/root/reference"):
        pytest.skip("%s was not built (no /root/reference where build() ran)" % request.param)
    lib = C.CDLL(path)
    lib.icapi_block_op.restype = C.c_long
    lib.icapi_block_op.argtypes = [C.c_int] * 4 + [C.c_uint] * 6 + [_u8p, C.c_size_t, _u8p, C.c_size_t, C.POINTER(C.c_uint32)]
    return lib


def class_op(api, op, codec, fmt, blocks, h, w, a=0, b=0, c=0, d=0, strategy=2, cap=1 << 20):
    blocks = np.ascontiguousarray(blocks)
    out = np.zeros(cap, np.uint8)
    meta = (C.c_uint32 * 7)()
    n = api.icapi_block_op(op, codec, strategy, fmt, h, w, a, b, c, d, blocks.ctypes.data_as(_u8p), blocks.size,
                           out.ctypes.data_as(_u8p), out.size, meta)
    return (out[:n].copy(), [int(x) for x in meta]) if n > 0 else (None, None)


def test_classes_block_ops(api, golden_ops):
    """Compressor::Downsample / Pad / CopySubimage / CreateSolidImage and TranscodeDxt1ToEtc1 of this build reproduce
    the reference's bytes, refusals and metadata (fixtures generated from the reference's own classes)."""
    for meta, src, want in golden_ops:
        op, codec, fmt = meta["op"], meta["codec"], meta["format"]
        st = meta.get("strategy", 2)
        if op == "downsample":
            got, m = class_op(api, 0, codec, fmt, src, meta["h"], meta["w"], strategy=st)
            if not meta["refused"]:
                assert m[1:5] == [(meta["h"] + 1) // 2, (meta["w"] + 1) // 2, 4 * ck.nblocks((meta["h"] + 1) // 2), 4 * ck.nblocks((meta["w"] + 1) // 2)]
        elif op == "pad":
            got, m = class_op(api, 1, codec, fmt, src, meta["h"], meta["w"], meta["ph"], meta["pw"], strategy=st)
        elif op == "copy_subimage":
            got, m = class_op(api, 2, codec, fmt, src, meta["h"], meta["w"], meta["row"], meta["col"], meta["sh"], meta["sw"])
        elif op == "solid":
            got, m = class_op(api, 3, codec, fmt, src, meta["h"], meta["w"])
        else:
            got, m = class_op(api, 4, codec, fmt, src, 32, 24 if src.size == 8 * 48 else src.size // 8 // 8 * 4)
        if meta["refused"]:
            assert got is None, meta
        else:
            assert got is not None and np.array_equal(got, want), meta
