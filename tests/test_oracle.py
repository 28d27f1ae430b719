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
