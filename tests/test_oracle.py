"""CPU tests of the checker itself: the plain-C oracle against (a) the reference-generated golden fixtures,
(b) the known-answer vectors SURVEY.md section 8c captured from the compiled reference, and (c) the compiled
reference directly when oracle/_ref exists (build container only)."""
import numpy as np
import pytest

import checkers as ck
import imagegen


def run_oracle(meta, buf):
    h, w, padded, padding = meta["h"], meta["w"], meta["padded"], meta["padding"]
    ch, cw = (padded if padded else (None, None))
    if meta["codec"] == "dxt":
        return ck.oracle_dxt(meta["format"], buf, h, w, ch, cw, padding)
    if meta["codec"] == "etc":
        return ck.oracle_etc1(meta["strategy"], buf, h, w, ch, cw, padding)
    return ck.oracle_pvrtc(buf, h, w)


def test_oracle_matches_golden(golden):
    assert len(golden) > 400
    for meta, buf, want in golden:
        got = run_oracle(meta, np.ascontiguousarray(buf))
        assert np.array_equal(got, want), meta


def _hex(a):
    return a.tobytes().hex()


def test_survey_known_answers():
    i = np.arange(16)
    A = np.stack([16 * i, 255 - 16 * i, 8 * i], -1).astype(np.uint8)
    assert _hex(ck.oracle_dxt(ck.RGB, A.ravel(), 4, 4)) == "8fe8e007d5ffaa02"
    assert _hex(ck.oracle_dxt(ck.BGR, A.ravel(), 4, 4)) == "9d78e007d5ffaa02"

    def rgba(alpha):
        return np.ascontiguousarray(np.concatenate([A, np.asarray(alpha, np.uint8)[:, None]], -1)).ravel()
    assert _hex(ck.oracle_dxt(ck.RGBA, rgba(17 * i), 4, 4)) == "ff00c96fb7e42601" "8fe8e007d5ffaa02"
    assert _hex(ck.oracle_dxt(ck.RGBA, rgba(64 + 8 * i), 4, 4)) == "b840c96fb7e42601" "8fe8e007d5ffaa02"
    a = 64 + 8 * i
    a[i < 3] = 0
    a[i > 12] = 255
    assert _hex(ck.oracle_dxt(ck.RGBA, rgba(a), 4, 4)) == "58a0b6216d649bff" "8fe8e007d5ffaa02"
    assert _hex(ck.oracle_etc1(ck.ETC_SMALLER_ERROR, A.ravel(), 4, 4)) == "3bc415010555aaa5"
    assert _hex(ck.oracle_etc1(ck.ETC_HEURISTIC, A.ravel(), 4, 4)) == "3bc4154905550000"
    assert _hex(ck.oracle_etc1(ck.ETC_SPLIT_H, A.ravel(), 4, 4)) == "3bc415010555aaa5"
    assert _hex(ck.oracle_etc1(ck.ETC_SPLIT_V, A.ravel(), 4, 4)) == "6897342433339999"
    B = np.tile(np.array([100, 150, 200], np.uint8), (16, 1))
    # constant block: identical bytes for kRGB and kBGR (the reference swaps twice, dxtc_compressor.cc:360)
    assert _hex(ck.oracle_dxt(ck.RGB, B.ravel(), 4, 4)) == "b864b86400000000"
    assert _hex(ck.oracle_dxt(ck.BGR, B.ravel(), 4, 4)) == "b864b86400000000"
    B4 = np.tile(np.array([100, 150, 200, 128], np.uint8), (16, 1))
    assert _hex(ck.oracle_dxt(ck.RGBA, B4.ravel(), 4, 4)) == "8080000000000000" "b864b86400000000"
    assert _hex(ck.oracle_etc1(ck.ETC_SMALLER_ERROR, B.ravel(), 4, 4)) == "6090c802ffff0000"
    yy, xx = np.mgrid[0:5, 0:5]
    Cc = np.ascontiguousarray(np.stack([40 * xx, 50 * yy, 25 * (xx + yy)], -1).astype(np.uint8))
    assert _hex(ck.oracle_dxt(ck.RGB, Cc.ravel(), 5, 5)) == ("b27c0000f5bfab0a" "b59c0c9855ffaa00" "357e2c062d2d2d2d" "9244ffffaaaaaaaa")
    assert _hex(ck.oracle_etc1(ck.ETC_SMALLER_ERROR, Cc.ravel(), 5, 5)) == ("3317366d005faa07" "aa177a6555554444" "16cc7a440f0f0f00" "a0c8c826ffff0000")
    yy, xx = np.mgrid[0:8, 0:8]
    D = np.ascontiguousarray(np.stack([32 * xx, 32 * yy, 16 * (xx + yy), np.full_like(xx, 255)], -1).astype(np.uint8))
    assert _hex(ck.oracle_pvrtc(D.ravel(), 8, 8)) == "51a4b4e5018094f1" "a4e5a4f909829cf3"
    D = np.ascontiguousarray(np.stack([32 * xx, 32 * yy, 16 * (xx + yy), 36 * xx], -1).astype(np.uint8))
    assert _hex(ck.oracle_pvrtc(D.ravel(), 8, 8)) == "91e4b4e5010095f5" "a5e5b4e58500bdf7"


def test_rgba_extension_equals_alpha_strip():
    """The DXT1-from-RGBA8 extension is defined as: strip alpha, run the 3-component reference path."""
    for kind in imagegen.KINDS:
        for swap, fmt in ((0, ck.RGB), (1, ck.BGR)):
            img = imagegen.make(kind, 24, 20, 4, seed=5)
            rgb = np.ascontiguousarray(img[..., :3])
            want = ck.oracle_dxt(fmt, rgb.ravel(), 24, 20)
            got = ck.oracle_dxt1_rgba(img.ravel(), 24, 20, swap_rb=swap)
            assert np.array_equal(got, want), (kind, swap)


def test_synthetic_stream_is_offset_consistent():
    whole = ck.synthetic(1000, 7)
    for off, n in ((0, 10), (3, 50), (8, 64), (13, 987)):
        assert np.array_equal(ck.synthetic(n, 7, off), whole[off:off + n])
    assert ck.fnv1a64(np.frombuffer(b"hello", np.uint8)) == 0xa430d84680aabd0b


@pytest.mark.skipif(not ck.have_ref(), reason="compiled reference (oracle/_ref) only exists in the build container")
class TestAgainstCompiledReference:
    def test_dxt_all_formats(self):
        for fmt in (ck.RGB, ck.BGR, ck.RGBA, ck.BGRA):
            nc = ck.ncomp(fmt)
            for kind in imagegen.KINDS:
                for (h, w) in ((4, 4), (7, 9), (20, 16), (1, 5), (33, 47)):
                    img = imagegen.make(kind, h, w, nc, seed=3)
                    for padded in (None, (h + 5, w + 9), (h, w + 4), (h + 4, w)):
                        r = ck.ref_dxt(fmt, img.ravel(), h, w, padded=padded)
                        o = ck.oracle_dxt(fmt, img.ravel(), h, w, *(padded or (None, None)))
                        assert np.array_equal(r, o), (fmt, kind, h, w, padded)

    def test_dxt_row_padding(self):
        for fmt in (ck.RGB, ck.RGBA):
            img = imagegen.make("random", 19, 11, ck.ncomp(fmt), 1)
            buf, _ = imagegen.with_row_padding(img, 13)
            assert np.array_equal(ck.ref_dxt(fmt, buf, 19, 11, padding=13), ck.oracle_dxt(fmt, buf, 19, 11, padding=13))

    def test_etc_all_strategies(self):
        for st in range(4):
            for kind in imagegen.KINDS:
                for (h, w) in ((4, 4), (7, 9), (20, 16), (33, 21)):
                    img = imagegen.make(kind, h, w, 3, seed=st)
                    for padded in (None, (h + 5, w + 9)):
                        r = ck.ref_etc(st, img.ravel(), h, w, padded=padded)
                        o = ck.oracle_etc1(st, img.ravel(), h, w, *(padded or (None, None)))
                        assert np.array_equal(r, o), (st, kind, h, w, padded)

    def test_pvrtc(self):
        for s in (8, 16, 32, 64, 128):
            for kind in imagegen.KINDS:
                img = imagegen.make(kind, s, s, 4, seed=s)
                assert np.array_equal(ck.ref_pvrtc(img.ravel(), s, s), ck.oracle_pvrtc(img.ravel(), s, s)), (s, kind)

    def test_reference_rejections(self):
        img = imagegen.make("random", 16, 16, 4, 0).ravel()
        assert ck.ref_pvrtc(img, 8, 16) is None        # not square
        assert ck.ref_pvrtc(img, 4, 4) is None         # smaller than a block
        assert ck.ref_pvrtc(img, 8, 8, padding=4) is None
        assert ck.ref_etc(2, img, 8, 8, fmt=ck.RGBA) is None
        assert ck.ref_etc(2, img, 8, 8, fmt=ck.BGR) is None

    def test_synthetic_1024_medium(self):
        n = 256
        rgb, rgba = ck.synthetic(n * n * 3, 1), ck.synthetic(n * n * 4, 2)
        assert np.array_equal(ck.ref_dxt(ck.RGB, rgb, n, n), ck.oracle_dxt(ck.RGB, rgb, n, n))
        assert np.array_equal(ck.ref_dxt(ck.RGBA, rgba, n, n), ck.oracle_dxt(ck.RGBA, rgba, n, n))
        assert np.array_equal(ck.ref_etc(2, rgb, n, n), ck.oracle_etc1(2, rgb, n, n))
        assert np.array_equal(ck.ref_pvrtc(rgba, n, n), ck.oracle_pvrtc(rgba, n, n))


@pytest.mark.skipif(not ck.have_ref(), reason="compiled reference (oracle/_ref) only exists in the build container")
def test_oracle_decoders_match_reference_decompress():
    rng = np.random.default_rng(5)
    for fmt, codec in ((ck.RGB, 0), (ck.BGR, 0), (ck.RGBA, 1), (ck.BGRA, 1), (ck.RGB, 2)):
        bb = 16 if codec == 1 else 8
        for (h, w) in ((4, 4), (8, 12), (5, 7), (16, 16), (1, 1), (13, 3)):
            nb = ck.nblocks(h) * ck.nblocks(w)
            for kind in range(3):
                blocks = rng.integers(0, 256, nb * bb, dtype=np.uint8)
                if kind == 1:
                    img = imagegen.make("smooth_noise", h, w, ck.ncomp(fmt), 3).ravel()
                    blocks = ck.oracle_etc1(2, img, h, w) if codec == 2 else ck.oracle_dxt(fmt, img, h, w)
                elif kind == 2 and codec != 2:
                    v = blocks.reshape(-1, bb)
                    v[:, bb - 6:bb - 4] = v[:, bb - 8:bb - 6]  # c0 == c1
                r = ck.ref_decompress(codec, fmt, blocks, h, w)
                o = ck.oracle_decode(codec, blocks, h, w, swap_rb=1 if fmt in (ck.BGR, ck.BGRA) else 0)
                assert r is not None and np.array_equal(r, o), (fmt, codec, h, w, kind)


def test_decode_golden(golden):
    """Decoder fixture: decode of every golden DXT/ETC output must reproduce a frozen checksum table."""
    import json
    import os
    import zlib
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode_crc_v1.json")
    want = json.load(open(path))
    got = {}
    for i, (meta, _, blocks) in enumerate(golden):
        if meta["codec"] == "pvrtc" or meta["padded"]:
            continue
        codec = 2 if meta["codec"] == "etc" else (0 if meta["ncomp"] == 3 else 1)
        swap = 1 if meta["format"] in (ck.BGR, ck.BGRA) else 0
        px = ck.oracle_decode(codec, np.ascontiguousarray(blocks), meta["h"], meta["w"], swap_rb=swap)
        got[str(i)] = zlib.crc32(px.tobytes())
    assert got == want


def _compress_any(codec, fmt, img, h, w, strategy=ck.ETC_SMALLER_ERROR):
    return ck.oracle_etc1(strategy, img.ravel(), h, w) if codec == 2 else ck.oracle_dxt(fmt, img.ravel(), h, w)


CODEC_FORMATS = ((0, ck.RGB), (0, ck.BGR), (1, ck.RGBA), (1, ck.BGRA), (2, ck.RGB))


@pytest.mark.skipif(not ck.have_ref(), reason="compiled reference (oracle/_ref) only exists in the build container")
class TestBlockOpsAgainstCompiledReference:
    """Compressed-domain operations: oracle restatement vs the reference."""

    def test_downsample(self):
        sizes = ((8, 8), (16, 24), (32, 8), (8, 40), (13, 29), (5, 7), (4, 16), (24, 4), (3, 16), (16, 2), (4, 4),
                 (2, 2), (1, 1), (1, 4), (4, 2), (2, 1), (3, 4), (4, 3), (12, 8), (8, 20), (64, 64))
        for codec, fmt in CODEC_FORMATS:
            for kind in ("random", "smooth_noise", "constant", "two_colour", "alpha_extremes", "dark"):
                for (h, w) in sizes:
                    img = imagegen.make(kind, h, w, ck.ncomp(fmt), seed=7)
                    for st in ((2, 3) if codec == 2 else (2,)):
                        blocks = _compress_any(codec, fmt, img, h, w, st)
                        r, meta = ck.ref_downsample(codec, fmt, blocks, h, w, strategy=st)
                        o = ck.oracle_downsample(codec, blocks, h, w, strategy=st)
                        if r is None:
                            assert o is None, (codec, fmt, kind, h, w)
                            continue
                        assert o is not None and np.array_equal(r, o), (codec, fmt, kind, h, w, st)
                        assert (meta["uncompressed_height"], meta["uncompressed_width"]) == ((h + 1) // 2, (w + 1) // 2)

    def test_downsample_random_blocks(self):
        """Arbitrary bit patterns (3-colour DXT1 blocks, ETC differential overflow) go through decode -> encode."""
        rng = np.random.default_rng(11)
        for codec, fmt in CODEC_FORMATS:
            for (h, w) in ((16, 16), (8, 32), (4, 4), (4, 8)):
                blocks = rng.integers(0, 256, ck.nblocks(h) * ck.nblocks(w) * ck.block_bytes(codec), dtype=np.uint8)
                r, _ = ck.ref_downsample(codec, fmt, blocks, h, w)
                assert np.array_equal(r, ck.oracle_downsample(codec, blocks, h, w)), (codec, fmt, h, w)

    def test_pad(self):
        rng = np.random.default_rng(12)
        for codec, fmt in CODEC_FORMATS:
            for (h, w) in ((8, 8), (5, 7), (16, 4), (4, 4), (12, 20)):
                for (ph, pw) in ((h + 9, w + 6), (h, w + 8), (h + 4, w), (h, w), (4, 4), (h + 1, w + 1), (32, 32)):
                    if ph < 4 * ck.nblocks(h) and pw > 4 * ck.nblocks(w) or pw < 4 * ck.nblocks(w) and ph > 4 * ck.nblocks(h):
                        continue  # the reference overruns its output buffer here (helper.h:419-440)
                    for content in ("image", "random"):
                        if content == "image":
                            blocks = _compress_any(codec, fmt, imagegen.make("smooth_noise", h, w, ck.ncomp(fmt), 1), h, w)
                        else:
                            blocks = rng.integers(0, 256, ck.nblocks(h) * ck.nblocks(w) * ck.block_bytes(codec), dtype=np.uint8)
                        for st in ((2, 3) if codec == 2 else (2,)):
                            r, meta = ck.ref_pad(codec, fmt, blocks, h, w, ph, pw, strategy=st)
                            o = ck.oracle_pad(codec, blocks, 4 * ck.nblocks(h), 4 * ck.nblocks(w), ph, pw, strategy=st)
                            assert r is not None and np.array_equal(r, o), (codec, fmt, h, w, ph, pw, content, st)

    def test_solid_and_transcode(self):
        rng = np.random.default_rng(13)
        for codec, fmt in CODEC_FORMATS:
            for _ in range(40):
                color = rng.integers(0, 256, 4, dtype=np.uint8)
                r, meta = ck.ref_solid(codec, fmt, 9, 6, color)
                blk = ck.oracle_solid_block(codec, color)
                assert np.array_equal(r, np.tile(blk, 3 * 2)), (codec, fmt, color)
        assert ck.oracle_solid_block(2, (100, 150, 200)).tobytes().hex() == "6090c80200000000"
        for kind in ("random", "smooth_noise", "constant", "two_colour", "dark"):
            img = imagegen.make(kind, 32, 24, 3, seed=2)
            blocks = ck.oracle_dxt(ck.RGB, img.ravel(), 32, 24)
            assert np.array_equal(ck.ref_transcode(blocks), ck.oracle_transcode(blocks)), kind
        blocks = rng.integers(0, 256, 8 * 500, dtype=np.uint8)
        assert np.array_equal(ck.ref_transcode(blocks), ck.oracle_transcode(blocks))

    def test_copy_subimage_reference_rules(self):
        img = imagegen.make("random", 16, 24, 3, 0)
        blocks = ck.oracle_dxt(ck.RGB, img.ravel(), 16, 24)
        r, meta = ck.ref_copy_subimage(0, ck.RGB, blocks, 16, 24, 4, 8, 8, 12)
        want = blocks.reshape(4, 6, 8)[1:3, 2:5].ravel()
        assert np.array_equal(r, want) and meta["uncompressed_height"] == 8 and meta["uncompressed_width"] == 12
        assert ck.ref_copy_subimage(0, ck.RGB, blocks, 16, 24, 2, 8, 8, 12)[0] is None   # not a multiple of 4
        assert ck.ref_copy_subimage(0, ck.RGB, blocks, 16, 24, 12, 8, 8, 12)[0] is None  # leaves the image


def test_oracle_block_ops_match_golden(golden_ops):
    """The restated compressed-domain operations against reference outputs frozen by tools/gen_golden_ops.py."""
    assert len(golden_ops) > 150
    seen = set()
    for meta, src, want in golden_ops:
        op, codec = meta["op"], meta["codec"]
        seen.add(op)
        src = np.ascontiguousarray(src)
        if op == "downsample":
            got = ck.oracle_downsample(codec, src, meta["h"], meta["w"], strategy=meta["strategy"])
            assert (got is None) == meta["refused"], meta
        elif op == "pad":
            got = ck.oracle_pad(codec, src, 4 * ck.nblocks(meta["h"]), 4 * ck.nblocks(meta["w"]), meta["ph"], meta["pw"],
                                strategy=meta["strategy"])
        elif op == "copy_subimage":
            if meta["refused"]:
                continue  # argument rules live in the host classes (tests/test_cpp_api.py)
            bb, cols = ck.block_bytes(codec), ck.nblocks(meta["w"])
            grid = src.reshape(ck.nblocks(meta["h"]), cols, bb)
            got = grid[meta["row"] // 4:(meta["row"] + meta["sh"]) // 4, meta["col"] // 4:(meta["col"] + meta["sw"]) // 4].ravel()
        elif op == "solid":
            got = np.tile(ck.oracle_solid_block(codec, src), ck.nblocks(meta["h"]) * ck.nblocks(meta["w"]))
        else:
            got = ck.oracle_transcode(src)
        if not meta["refused"]:
            assert np.array_equal(got, want), meta
    assert seen == {"downsample", "pad", "copy_subimage", "solid", "transcode"}


def test_threaded_stripe_reference_equals_one_thread():
    """cpu_encode_full (the full-size checker of tests/test_parity_gpu.py and bench.py's in-run parity flag): T row stripes
    from T threads == the one-thread answer, for every 4x4 workload, uneven stripe counts included."""
    for wl, nc in (("dxt1_rgba8", 4), ("dxt1_rgb8", 3), ("dxt5_rgba8", 4), ("etc1_rgb8", 3)):
        h, w = 52, 64
        img = ck.synthetic(h * w * nc, 3)
        if wl == "dxt1_rgba8":
            want = ck.oracle_dxt1_rgba(img, h, w)
        elif wl == "etc1_rgb8":
            want = ck.oracle_etc1(ck.ETC_SMALLER_ERROR, img, h, w)
        else:
            want = ck.oracle_dxt(ck.RGB if nc == 3 else ck.RGBA, img, h, w)
        for threads in (1, 3, 5, 64):
            got, kind = ck.cpu_encode_full(wl, img, h, w, threads=threads)
            assert kind == ("reference" if ck.have_ref() else "port")
            assert np.array_equal(got, want), (wl, threads)
    img = ck.synthetic(32 * 32 * 4, 2)
    got, _ = ck.cpu_encode_full("pvrtc2_rgba8", img, 32, 32)
    assert np.array_equal(got, ck.oracle_pvrtc(img, 32, 32))
