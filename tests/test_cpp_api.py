"""The drop-in C++ classes (image_codec_compression::{Dxtc,Etc,Pvrtc}Compressor of this build) driven through
lib/libicb_api_test.so exactly the way oracle/ref_shim.cc drives the reference's classes."""
import ctypes as C
import os

import numpy as np
import pytest

import checkers as ck
import imagegen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


@pytest.fixture(scope="module")
def api():
    path = os.path.join(ROOT, "image_compression_b200", "lib", "libicb_api_test.so")
    if not os.path.exists(path):
        pytest.fail("%s not built (run __graft_entry__.build())" % path)
    lib = C.CDLL(path)
    lib.icapi_dxt.restype = C.c_long
    lib.icapi_dxt.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t, _u32p, C.c_int]
    lib.icapi_etc.restype = C.c_long
    lib.icapi_etc.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t, _u32p, C.c_int]
    lib.icapi_pvrtc.restype = C.c_long
    lib.icapi_pvrtc.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t, _u32p, C.c_int]
    lib.icapi_size.restype = C.c_size_t
    lib.icapi_size.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint]
    return lib


def call(fn, args, cap, external=0):
    out = np.zeros(max(cap, 16), np.uint8)
    meta = (C.c_uint32 * 7)()
    n = fn(*args, out.ctypes.data_as(_u8p), cap if external else out.size, meta, external)
    if n <= 0:
        return None, None
    return out[:n].copy(), [int(x) for x in meta]


def test_sizes_match_reference_rules(api):
    cases = [(0, ck.RGB, 8192, 8192, 33554432), (0, ck.RGBA, 8192, 8192, 67108864), (1, ck.RGB, 4096, 4096, 8388608),
             (2, ck.RGBA, 4096, 4096, 4194304), (0, ck.RGB, 5, 5, 32), (0, ck.BGRA, 1, 1, 16), (1, ck.RGBA, 8, 8, 0),
             (0, ck.RGB, 0, 7, 0), (1, ck.RGB, 7, 0, 0)]
    for codec, fmt, h, w, want in cases:
        assert api.icapi_size(codec, fmt, h, w) == want
        if ck.have_ref():
            assert ck.ref().icref_size(codec, fmt, h, w) == want


def test_rejections_before_any_cuda_call(api):
    img = np.zeros(16 * 16 * 4, np.uint8)
    p = img.ctypes.data_as(_u8p)
    assert call(api.icapi_dxt, (ck.RGB, 0, 4, 0, 0, 0, 0, p), 64)[0] is None            # zero height
    assert call(api.icapi_etc, (2, ck.RGBA, 8, 8, 0, 0, 0, 0, p), 64)[0] is None         # ETC is kRGB only
    assert call(api.icapi_etc, (2, ck.BGR, 8, 8, 0, 0, 0, 0, p), 64)[0] is None
    assert call(api.icapi_pvrtc, (ck.RGBA, 8, 16, 0, p), 64)[0] is None                  # not square
    assert call(api.icapi_pvrtc, (ck.RGBA, 4, 4, 0, p), 64)[0] is None                   # smaller than a block
    assert call(api.icapi_pvrtc, (ck.RGBA, 8, 8, 4, p), 64)[0] is None                   # padded rows
    assert call(api.icapi_dxt, (ck.RGBA, 8, 8, 0, 0, 0, 0, p), 7, external=1)[0] is None  # external storage, wrong size


@pytest.mark.gpu
def test_compress_matches_oracle_and_metadata(api):
    for fmt in (ck.RGB, ck.BGR, ck.RGBA, ck.BGRA):
        nc = ck.ncomp(fmt)
        for (h, w, padded, padding) in ((16, 16, None, 0), (9, 13, None, 3), (21, 34, (40, 64), 0), (64, 256, None, 0)):
            img = imagegen.make("smooth_noise", h, w, nc, seed=21)
            buf, _ = imagegen.with_row_padding(img, padding)
            ph, pw = padded if padded else (0, 0)
            ch, cw = max(h, ph), max(w, pw)
            cap = ck.nblocks(ch) * ck.nblocks(cw) * 16
            for external in (0, 1):
                exact = ck.nblocks(ch) * ck.nblocks(cw) * (8 if nc == 3 else 16)
                got, meta = call(api.icapi_dxt, (fmt, h, w, 1 if padded else 0, ph, pw, padding, buf.ctypes.data_as(_u8p)),
                                 exact if external else cap, external)
                want = ck.oracle_dxt(fmt, buf, h, w, *(padded or (None, None)), padding)
                assert got is not None and np.array_equal(got, want), (fmt, h, w, padded, padding, external)
                # CompressAndPad reports the padded size as the uncompressed size (compressor4x4_helper.h:486-492)
                assert meta == [fmt, ch, cw, 4 * ck.nblocks(ch), 4 * ck.nblocks(cw), padding, 4]
                if ck.have_ref():
                    r, rmeta = ck.ref_dxt(fmt, buf, h, w, padded=padded, padding=padding, want_meta=True)
                    assert np.array_equal(r, got) and list(rmeta.values()) == meta
    for strategy in range(4):
        img = imagegen.make("random", 20, 24, 3, seed=22)
        got, meta = call(api.icapi_etc, (strategy, ck.RGB, 20, 24, 0, 0, 0, 0, img.ctypes.data_as(_u8p)), 4096)
        assert np.array_equal(got, ck.oracle_etc1(strategy, img.ravel(), 20, 24))
        assert meta == [ck.RGB, 20, 24, 20, 24, 0, 3]
    img = imagegen.make("alpha_extremes", 32, 32, 4, seed=23)
    for fmt in (ck.RGBA, ck.RGB):  # PvrtcCompressor::Compress ignores the format value (pvrtc_compressor.cc:664)
        got, meta = call(api.icapi_pvrtc, (fmt, 32, 32, 0, img.ctypes.data_as(_u8p)), 4096)
        assert np.array_equal(got, ck.oracle_pvrtc(img.ravel(), 32, 32))
        assert meta == [fmt, 32, 32, 32, 32, 0, 5]


@pytest.mark.gpu
def test_compress_then_decompress_matches_reference_round_trip(api):
    """Decompress() of the drop-in classes == oracle decode of oracle encode (and == the reference when present)."""
    api.icapi_roundtrip.restype = C.c_long
    api.icapi_roundtrip.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t]
    for codec, fmt in ((0, ck.RGB), (0, ck.BGR), (1, ck.RGBA), (1, ck.BGRA), (2, ck.RGB)):
        nc = ck.ncomp(fmt)
        for (h, w) in ((16, 16), (9, 13), (40, 28)):
            img = imagegen.make("smooth_noise", h, w, nc, seed=24)
            out = np.zeros(h * w * 4, np.uint8)
            n = api.icapi_roundtrip(codec, 2, fmt, h, w, img.ctypes.data_as(_u8p), out.ctypes.data_as(_u8p), out.size)
            assert n == h * w * nc
            blocks = ck.oracle_etc1(2, img.ravel(), h, w) if codec == 2 else ck.oracle_dxt(fmt, img.ravel(), h, w)
            want = ck.oracle_decode(codec, blocks, h, w, swap_rb=1 if fmt in (ck.BGR, ck.BGRA) else 0)
            assert np.array_equal(out[:n], want), (codec, fmt, h, w)
            if ck.have_ref():
                assert np.array_equal(ck.ref_decompress(codec, fmt, blocks, h, w), want)
