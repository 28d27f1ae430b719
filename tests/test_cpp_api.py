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


# Every test below runs twice: through the doorway compiled against THIS build's headers (libicb_api_test.so) and
# through the same source compiled against the REFERENCE's own headers (libicb_refabi_test.so, built by
# image_compression_b200/cpp/Makefile where /root/reference is mounted; the .so travels to the GPU box).  The second
# is the link-compatibility check: an object that only ever saw upstream's compressor.h dispatching into this library.
API_BUILDS = {"repo_headers": "libicb_api_test.so", "reference_headers": "libicb_refabi_test.so"}


def load_api(build):
    path = os.path.join(ROOT, "image_compression_b200", "lib", API_BUILDS[build])
    if not os.path.exists(path):
        if build == "reference_headers" and not os.path.isdir("/root/reference"):
            pytest.skip("%s was not built (no /root/reference where build() ran)" % API_BUILDS[build])
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
    lib.icapi_all_virtuals.restype = C.c_long
    lib.icapi_all_virtuals.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t, _u32p]
    return lib


@pytest.fixture(scope="module", params=sorted(API_BUILDS))
def api(request):
    return load_api(request.param)


def test_vtable_layout_matches_reference():
    """Member offsets, class sizes and the ORDER OF THE VIRTUALS of this build's public headers equal the reference's
    (public/compressor.h:52-137), as g++ lays them out: against the frozen fixture everywhere, and against the
    reference's headers themselves where they are mounted."""
    import subprocess
    import sys
    tool = os.path.join(ROOT, "tools", "gen_class_layout.py")
    mine = subprocess.run([sys.executable, tool, "--print", os.path.join(ROOT, "image_compression_b200", "cpp")],
                          check=True, capture_output=True, text=True).stdout
    frozen = open(os.path.join(ROOT, "tests", "golden", "ref_class_layout_v1.txt")).read()
    assert "DxtcCompressor::CompressAndPad" in frozen and frozen.count("Vtable for") == 4
    assert mine == frozen
    if os.path.isdir("/root/reference"):
        live = subprocess.run([sys.executable, tool, "--print", "/root/reference"], check=True, capture_output=True, text=True).stdout
        assert live == frozen


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


def all_virtuals(api, codec, fmt, img, h, w):
    out = np.zeros(1 << 16, np.uint8)
    sizes = (C.c_uint32 * 10)()
    mask = api.icapi_all_virtuals(codec, fmt, h, w, img.ctypes.data_as(_u8p), out.ctypes.data_as(_u8p), out.size, sizes)
    segs, off = [], 0
    for n in sizes:
        segs.append(out[off:off + n].copy())
        off += n
    return mask, segs


def test_non_compute_virtuals_dispatch_without_gpu(api):
    """SupportsFormat / ComputeCompressedDataSize land in the right vtable slot whichever headers the caller was
    compiled against (no device needed: slots 0 and 2; slot 1 needs a compressed image)."""
    img = np.zeros(16 * 16 * 4, np.uint8)
    for codec, fmt, size in ((0, ck.RGB, 128), (0, ck.RGBA, 256), (2, ck.RGB, 128)):
        mask, segs = all_virtuals(api, codec, fmt, img, 16, 16)
        assert mask & 0b101 == 0b101
        assert segs[0][0] == 1 and int.from_bytes(segs[2].tobytes(), "little") == size


@pytest.mark.gpu
def test_all_ten_virtuals_through_base_pointer(api):
    """Every virtual of Compressor, called through a Compressor* in vtable order, returns what the oracle says that
    METHOD returns -- a caller whose vtable slots were shifted (CompressAndPad declared out of order)
    would get Decompress where it asked for Downsample and fail here."""
    for codec, fmt in ((0, ck.RGB), (0, ck.RGBA), (0, ck.BGRA), (2, ck.RGB)):
        nc = ck.ncomp(fmt)
        h, w = 24, 32
        img = imagegen.make("smooth_noise", h, w, nc, seed=31)
        mask, segs = all_virtuals(api, codec, fmt, img, h, w)
        assert mask == 0x3ff, (codec, fmt, bin(mask))
        icodec = 2 if codec == 2 else (0 if nc == 3 else 1)   # checkers' numbering: DXT1, DXT5, ETC1
        bb = ck.block_bytes(icodec)
        enc = (lambda buf, hh, ww, ph=None, pw=None: ck.oracle_etc1(2, buf, hh, ww, ph, pw)) if codec == 2 else \
              (lambda buf, hh, ww, ph=None, pw=None: ck.oracle_dxt(fmt, buf, hh, ww, ph, pw))
        blocks = enc(img.ravel(), h, w)
        assert segs[0][0] == 1 and segs[1][0] == 1
        assert int.from_bytes(segs[2].tobytes(), "little") == blocks.size
        assert np.array_equal(segs[3], blocks)                                                    # Compress
        assert np.array_equal(segs[4], ck.oracle_decode(icodec, blocks, h, w, swap_rb=1 if fmt in (ck.BGR, ck.BGRA) else 0))  # Decompress
        assert np.array_equal(segs[5], ck.oracle_downsample(icodec, blocks, h, w))                # Downsample
        assert np.array_equal(segs[6], ck.oracle_pad(icodec, blocks, h, w, h + 8, w + 12))        # Pad
        assert np.array_equal(segs[7], enc(img.ravel(), h, w, h + 8, w + 12))                     # CompressAndPad
        solid = ck.oracle_solid_block(icodec, img.ravel()[:nc])
        assert np.array_equal(segs[8], np.tile(solid, ck.nblocks(h) * ck.nblocks(w)))             # CreateSolidImage
        grid = blocks.reshape(ck.nblocks(h), ck.nblocks(w), bb)
        assert np.array_equal(segs[9], grid[1:3, 1:3].reshape(-1))                                # CopySubimage
