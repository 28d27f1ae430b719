"""CPU-only parity tests of the DEVICE arithmetic: the CUDA headers of image_compression_b200/csrc compiled for the
host with the intrinsics emulated (tests/hostemu/ -- test infrastructure, never part of the product) and stepped
through one emulated thread at a time, against the CPU oracle and the reference-generated golden fixtures.
Bit-exact: every comparison is array_equal.

What this covers that the GPU suite cannot cover here (the container has no GPU): the block encoders, decoders,
PVRTC kernels and compressed-domain operations themselves.  What it does not cover: the TMA ring drivers, launch
configuration and the memory system (tests/*_gpu.py, run on the B200 box)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import checkers as ck
import imagegen

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu")
CSRC = os.path.join(ck.ROOT, "image_compression_b200", "csrc")
_u8p = C.POINTER(C.c_uint8)
VOTES = (0, 1, 2)  # what the other lanes of the warp answer to a warp vote: like this lane / "no" / "no" to the first vote only (cuda_emulation.h)


def _ptr(a):
    return a.ctypes.data_as(_u8p)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "libhostemu.so")
    sources = [os.path.join(HERE, f) for f in ("hostemu.cc", "cuda_emulation.h")]
    sources += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc"))]
    # ICB_HOSTEMU_DEFS="-DICB_DXT_INT_PAIRS=0x0f ...": emulate an A/B variant of the device code (tools/build_variants.sh
    # builds the same flags for the GPU); such a build goes to its own file and never replaces the default one.
    defs = os.environ.get("ICB_HOSTEMU_DEFS", "").split()
    if defs:
        so = os.path.join(HERE, "libhostemu_variant.so")
    if defs or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in sources):
        # -ffp-contract=off: every float operation rounds on its own, as the SASS does (FADD / FFMA as written)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                        "-Wno-unused-function"] + defs + ["-o", so, os.path.join(HERE, "hostemu.cc")], check=True)
    lib = C.CDLL(so)
    lib.emu_encode4x4.argtypes = [C.c_int, C.c_int, _u8p] + [C.c_uint32] * 5 + [C.c_int, C.c_int, _u8p]
    lib.emu_dxt1_rgb888_rows.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _u8p]
    lib.emu_pvrtc2.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p]
    lib.emu_decode4x4.argtypes = [C.c_int, _u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _u8p, C.c_uint32]
    lib.emu_downsample4x4.argtypes = [C.c_int, C.c_int, _u8p, C.c_uint32, C.c_uint32, _u8p]
    lib.emu_pad4x4.argtypes = [C.c_int, C.c_int, _u8p] + [C.c_uint32] * 4 + [_u8p]
    lib.emu_fill_solid4x4.argtypes = [C.c_int, C.c_uint32, C.c_uint64, _u8p]
    lib.emu_transcode_dxt1_to_etc1.argtypes = [_u8p, C.c_uint64]
    yield lib
    lib.emu_set_vote(0)


def encode(emu, codec, nc, buf, h, w, pitch=None, coded=None, swap=0, strategy=2, vote=0):
    ch, cw = coded if coded else (h, w)
    ch, cw = max(h, ch), max(w, cw)
    out = np.zeros(ck.nblocks(ch) * ck.nblocks(cw) * ck.block_bytes(codec), np.uint8)
    buf = np.ascontiguousarray(buf)
    emu.emu_set_vote(vote)
    assert emu.emu_encode4x4(codec, nc, _ptr(buf), h, w, pitch or w * nc, ch, cw, swap, strategy, _ptr(out)) == 0
    return out


def pvrtc(emu, img, n, stripes=1):
    out = np.zeros(n * n // 4, np.uint8)
    img = np.ascontiguousarray(img)
    assert emu.emu_pvrtc2(_ptr(img), n, n, stripes, _ptr(out)) == 0
    return out


def test_emulated_instructions(emu):
    """The emulation layer itself: byte permutes, dp4a, SIMD min/max/add-clamp, packed-half rounding and saturation."""
    assert emu.emu_self_check() == 0


def test_golden_fixtures(emu, golden):
    """Every reference-generated fixture (tests/golden/golden_v1.npz) through the device code."""
    for meta, buf, want in golden:
        h, w, nc = meta["h"], meta["w"], meta["ncomp"]
        if meta["codec"] == "pvrtc":
            got = pvrtc(emu, buf, h)
        else:
            codec = 2 if meta["codec"] == "etc" else (0 if nc == 3 else 1)
            swap = 1 if meta["format"] in (ck.BGR, ck.BGRA) else 0
            for vote in VOTES:
                got = encode(emu, codec, nc, buf, h, w, pitch=w * nc + meta["padding"], coded=meta["padded"], swap=swap,
                             strategy=meta["strategy"], vote=vote)
                assert np.array_equal(got, want), (meta, vote)
        assert np.array_equal(got, want), meta


@pytest.mark.parametrize("fmt", [ck.RGB, ck.BGR, ck.RGBA, ck.BGRA])
def test_dxt_vs_oracle(emu, fmt):
    """DXT1 (3 components) / DXT5 (4), every image kind, ragged sizes; the warp-uniform fast path of the colour index
    search and the general path must both give the oracle's bytes (the vote only picks which one runs)."""
    nc = ck.ncomp(fmt)
    codec, swap = (0 if nc == 3 else 1), (1 if fmt in (ck.BGR, ck.BGRA) else 0)
    for kind in imagegen.KINDS:
        for (h, w) in ((32, 64), (13, 30), (1, 1), (5, 3)):
            img = imagegen.make(kind, h, w, nc, seed=21)
            want = ck.oracle_dxt(fmt, img.ravel(), h, w)
            for vote in VOTES:
                assert np.array_equal(encode(emu, codec, nc, img.ravel(), h, w, swap=swap, vote=vote), want), (kind, h, w, vote)


def test_dxt_padding_and_compress_and_pad(emu):
    for fmt in (ck.RGB, ck.BGRA):
        nc = ck.ncomp(fmt)
        codec, swap = (0 if nc == 3 else 1), (1 if fmt == ck.BGRA else 0)
        for (h, w, padding, coded) in ((3, 3, 1, None), (9, 13, 7, None), (8, 8, 0, (8, 40)), (8, 8, 0, (40, 8)), (33, 67, 5, (64, 128))):
            img = imagegen.make("smooth_noise", h, w, nc, seed=14)
            buf, pitch = imagegen.with_row_padding(img, padding)
            ch, cw = coded if coded else (None, None)
            got = encode(emu, codec, nc, buf, h, w, pitch=pitch, coded=coded, swap=swap)
            assert np.array_equal(got, ck.oracle_dxt(fmt, buf, h, w, ch, cw, padding)), (fmt, h, w, padding, coded)


@pytest.mark.parametrize("swap", [0, 1])
def test_dxt1_from_rgba_extension(emu, swap):
    for kind in imagegen.KINDS:
        img = imagegen.make(kind, 24, 36, 4, seed=22)
        want = ck.oracle_dxt1_rgba(img.ravel(), 24, 36, swap_rb=swap)
        for vote in VOTES:
            assert np.array_equal(encode(emu, 0, 4, img.ravel(), 24, 36, swap=swap, vote=vote), want), (kind, vote)


@pytest.mark.parametrize("swap", [0, 1])
def test_dxt1_rgb888_row_word_form(emu, swap):
    """The form the TMA consumers use for RGB888: luminance keys straight from the three 32-bit words of a block row."""
    for kind in imagegen.KINDS:
        for padding in (0, 4):
            img = imagegen.make(kind, 16, 48, 3, seed=23)
            buf, pitch = imagegen.with_row_padding(img, padding)
            want = ck.oracle_dxt(ck.BGR if swap else ck.RGB, buf, 16, 48, None, None, padding)
            for vote in VOTES:
                emu.emu_set_vote(vote)
                got = np.zeros(want.size, np.uint8)
                assert emu.emu_dxt1_rgb888_rows(_ptr(buf), 16, 48, pitch, swap, _ptr(got)) == 0
                assert np.array_equal(got, want), (kind, padding, vote)


def test_dxt5_alpha_statistics_corner_cases(emu):
    """ComputeBaseAlphas' rules (dxtc_compressor.cc:374-424): counts of 0 / 255 around the '> 1' threshold, all-extreme
    blocks, equal endpoints -- one block per case, colour held constant."""
    rng = np.random.default_rng(5)
    cases = []
    for n0 in (0, 1, 2, 15, 16):
        for n255 in (0, 1, 2, 16 - n0):
            if n0 + n255 > 16:
                continue
            for _ in range(6):
                a = rng.integers(1, 255, 16)
                a[:n0] = 0
                a[n0:n0 + n255] = 255
                cases.append(rng.permutation(a))
    for v in (0, 1, 127, 254, 255):
        cases.append(np.full(16, v))
    for lo, hi in ((1, 2), (100, 101), (253, 254), (1, 254), (7, 9)):
        cases.append(rng.choice([lo, hi], 16))
        cases.append(np.concatenate([[0], rng.choice([lo, hi], 15)]))
        cases.append(np.concatenate([[255], rng.choice([lo, hi], 15)]))
    img = np.zeros((4, 4 * len(cases), 4), np.uint8)
    img[..., :3] = (90, 160, 30)
    for k, a in enumerate(cases):
        img[:, 4 * k:4 * k + 4, 3] = np.asarray(a, np.uint8).reshape(4, 4)
    h, w = img.shape[:2]
    assert np.array_equal(encode(emu, 1, 4, img.ravel(), h, w), ck.oracle_dxt(ck.RGBA, img.ravel(), h, w))


@pytest.mark.parametrize("strategy", [0, 1, 2, 3])
def test_etc1_vs_oracle(emu, strategy):
    for kind in imagegen.KINDS:
        for (h, w) in ((16, 32), (9, 14)):
            img = imagegen.make(kind, h, w, 3, seed=24)
            want = ck.oracle_etc1(strategy, img.ravel(), h, w)
            # vote 0: this lane's no-clamp codewords take the line form; vote 1: the general form for every codeword
            for vote in (0, 1):
                assert np.array_equal(encode(emu, 2, 3, img.ravel(), h, w, strategy=strategy, vote=vote), want), (kind, h, w, vote)


def test_pvrtc_vs_oracle_whole_and_striped(emu):
    """Morph / Modulate / Pack over whole images, and the halo-stripe form (each stripe from a private copy of its rows
    plus wrapped halo rows, scratch poisoned in between): any number of stripes assembles the whole image."""
    for n in (8, 16, 32, 64):
        for kind in imagegen.KINDS:
            img = imagegen.make(kind, n, n, 4, seed=n + 1)
            want = ck.oracle_pvrtc(img.ravel(), n, n)
            assert np.array_equal(pvrtc(emu, img.ravel(), n), want), (n, kind)
            for parts in (2, 3, 4):
                if n // 4 // parts + 2 <= n // 4 and parts <= n // 4:
                    assert np.array_equal(pvrtc(emu, img.ravel(), n, parts), want), (n, kind, parts)


def test_pvrtc_fused_modulate_pack_kernel(emu):
    """Whole images of 256 x 256 and more run Morph + the fused Modulate/Pack kernel (one CTA per 32 x 8-block tile,
    modulation values through shared memory, the tile's right-hand pixel column and the row below included): its two
    phases, emulated CTA by CTA, give the oracle's bytes -- and the same bytes as the three-kernel pipeline, which
    stays in use for small images and stripes -- on every kind of content, tile seams and the toroidal wrap included."""
    emu.emu_set_pvrtc_unfused.argtypes = [C.c_int]
    try:
        for n in (256, 512):
            for kind in imagegen.KINDS:
                img = imagegen.make(kind, n, n, 4, seed=n + 7)
                want = ck.oracle_pvrtc(img.ravel(), n, n)
                emu.emu_set_pvrtc_unfused(0)
                assert np.array_equal(pvrtc(emu, img.ravel(), n), want), (n, kind, "fused")
                emu.emu_set_pvrtc_unfused(1)
                assert np.array_equal(pvrtc(emu, img.ravel(), n), want), (n, kind, "three kernels")
    finally:
        emu.emu_set_pvrtc_unfused(0)


def test_decoders_every_kind_of_block(emu):
    """Random bit patterns (blocks no encoder produces included) and encoder output, whole and cropped images."""
    rng = np.random.default_rng(9)
    for codec in (0, 1, 2):
        nc = 4 if codec == 1 else 3
        for (h, w) in ((16, 64), (7, 10), (4, 4), (1, 1), (33, 130)):
            blocks = rng.integers(0, 256, ck.nblocks(h) * ck.nblocks(w) * ck.block_bytes(codec), dtype=np.uint8)
            for swap in ((0, 1) if codec != 2 else (0,)):
                got = np.zeros(h * w * nc, np.uint8)
                assert emu.emu_decode4x4(codec, _ptr(blocks), h, w, ck.nblocks(w), swap, _ptr(got), w * nc) == 0
                assert np.array_equal(got, ck.oracle_decode(codec, blocks, h, w, swap_rb=swap)), (codec, h, w, swap)


def test_block_operations_golden_and_oracle(emu, golden_ops):
    """Downsample / Pad / CreateSolidImage / TranscodeDxt1ToEtc1 kernels against the reference-generated fixtures."""
    for meta, src, want in golden_ops:
        op, codec = meta["op"], meta["codec"]
        src = np.ascontiguousarray(src)
        st = meta.get("strategy", 2)
        got = np.zeros(want.size, np.uint8)
        if op == "downsample":
            status = emu.emu_downsample4x4(codec, st, _ptr(src), meta["h"], meta["w"], _ptr(got))
            if meta["refused"]:
                assert status == -4, meta
                continue
            assert status == 0
        elif op == "pad":
            ch, cw = 4 * ck.nblocks(meta["h"]), 4 * ck.nblocks(meta["w"])
            assert emu.emu_pad4x4(codec, st, _ptr(src), ch, cw, meta["ph"], meta["pw"], _ptr(got)) == 0
        elif op == "solid":
            packed = int(src[0]) | int(src[1]) << 8 | int(src[2]) << 16 | int(src[3]) << 24
            assert emu.emu_fill_solid4x4(codec, packed, want.size // ck.block_bytes(codec), _ptr(got)) == 0
        elif op == "transcode":
            got = src.copy()
            assert emu.emu_transcode_dxt1_to_etc1(_ptr(got), got.size // 8) == 0
        else:
            continue  # copy_subimage is a strided memcpy on the device, no kernel to emulate
        assert np.array_equal(got, want), meta


@pytest.mark.parametrize("codec,fmt", [(0, ck.RGB), (1, ck.RGBA), (2, ck.RGB)])
def test_mip_chain(emu, codec, fmt):
    """Compress, then halve down to one block: every level equals the oracle's (decode + average + re-encode)."""
    n, nc = 64, ck.ncomp(fmt)
    img = imagegen.make("smooth_noise", n, n, nc, seed=5)
    level = encode(emu, codec, nc, img.ravel(), n, n)
    want = ck.oracle_etc1(2, img.ravel(), n, n) if codec == 2 else ck.oracle_dxt(fmt, img.ravel(), n, n)
    size = n
    while size > 4:
        assert np.array_equal(level, want), size
        out = np.zeros(want.size // 4 if size > 4 else want.size, np.uint8)
        assert emu.emu_downsample4x4(codec, 2, _ptr(level), size, size, _ptr(out)) == 0
        level, want = out, ck.oracle_downsample(codec, want, size, size)
        size //= 2
    assert np.array_equal(level, want)


# ---- exhaustive and large sweeps: the emulated device code is fast enough to leave random sampling behind ----------

def _blocks_to_plane(blocks16, cols=1024):
    """(n, 16) block rows -> a (4*rows, 4*cols) plane whose 4x4 block (br, bc) is blocks16[br * cols + bc]."""
    n = len(blocks16)
    rows = (n + cols - 1) // cols
    padded = np.concatenate([blocks16, np.tile(blocks16[:1], (rows * cols - n, 1))])
    return np.ascontiguousarray(padded.reshape(rows, cols, 4, 4).transpose(0, 2, 1, 3).reshape(rows * 4, cols * 4))


def _families(make):
    out = []

    def family(prefix, vals, prefix_last=False):
        k = 16 - len(prefix)
        vals = np.asarray(vals, np.uint8)
        vals = np.concatenate([vals, np.full((-len(vals)) % k, vals[0], np.uint8)]).reshape(-1, k)
        pre = np.tile(np.asarray(prefix, np.uint8), (len(vals), 1))
        out.append(np.concatenate([vals, pre] if prefix_last else [pre, vals], axis=1))

    make(family)
    return np.concatenate(out)


def test_dxt5_alpha_search_every_mode_endpoint_pair_and_alpha(emu):
    """EXHAUSTIVE for the alpha half: every (6-/8-alpha mode, a0, a1) the statistics can produce, with every alpha value
    between the endpoints (and 0 / 255 where the mode has them as explicit candidates) -- 536,064 blocks.  The crossing
    table was verified as a Python model when it was generated (tools/gen_dxt5_alpha_table.py); this runs the device
    function itself (packed-half statistics, table look-up, seven-crossing search, bit packing)."""
    def make(family):
        for lo in range(1, 255):
            for hi in range(lo, 255):
                inner = np.arange(lo, hi + 1)
                family([lo, hi], inner)                    # 8-alpha mode: a0 = hi, a1 = lo (lo == hi: the '<=' branch)
                family([0, 0, lo, hi], inner)              # 6-alpha mode through two zeros: a0 = lo, a1 = hi
                if (lo + hi) % 7 == 0:                     # the other ways into 6-alpha mode, thinned
                    family([255, 255, lo, hi], inner)
                    family([0, 255, 255, lo, hi], inner)
        for hi in range(1, 255):
            family([0, hi], np.arange(1, hi + 1))          # exactly one zero: a1 = 0
            family([hi, 255], np.arange(hi, 255))          # exactly one 255: a0 = 255
        family([0, 255], np.arange(1, 255))
        family([0, 0, 255, 255], [0, 255])                 # nothing but extremes: 0 / 255 endpoints
    plane = _blocks_to_plane(_families(make))
    img = np.empty(plane.shape + (4,), np.uint8)
    img[..., :3] = (90, 160, 30)
    img[..., 3] = plane
    h, w = plane.shape
    want = ck.oracle_dxt(ck.RGBA, img.ravel(), h, w)
    # the walk has a warp-uniform form for 8-alpha blocks with distinct candidates and a table-driven one: every answer
    # of the vote that picks between them must give the same bytes
    for vote in VOTES:
        assert np.array_equal(encode(emu, 1, 4, img.ravel(), h, w, vote=vote), want), vote


@pytest.mark.parametrize("fmt", [ck.RGB, ck.BGR])
def test_dxt1_colour_search_every_endpoint_pair_on_an_axis(emu, fmt):
    """EXHAUSTIVE along single colour axes: every pair of extreme values lo <= hi with every value in between, the
    extremes placed first or last in raster order (first-minimum / first-maximum rules), on the gray axis, on each
    single channel and on a skewed axis -- 217,645 blocks per axis, under both answers of the warp vote."""
    def make(family):
        for lo in range(256):
            for hi in range(lo, 256):
                family([lo, hi] if (lo + hi) & 1 else [hi, lo], np.arange(lo, hi + 1), prefix_last=bool(lo & 1))
    plane = _blocks_to_plane(_families(make)).astype(np.uint16)
    h, w = plane.shape
    swap = 1 if fmt == ck.BGR else 0
    for weights in ((256, 256, 256), (256, 0, 0), (0, 256, 0), (0, 0, 256), (0, 128, 256), (256, 85, 170)):
        img = np.ascontiguousarray(np.stack([(plane * k) >> 8 for k in weights], -1).astype(np.uint8))
        want = ck.oracle_dxt(fmt, img.ravel(), h, w)
        for vote in VOTES:
            assert np.array_equal(encode(emu, 0, 3, img.ravel(), h, w, swap=swap, vote=vote), want), (weights, vote)


def test_dxt_soak_small_palettes(emu):
    """Two million blocks whose 16 pixels come from palettes of 1-4 nearby colours: equal luminances with different
    colours, constant blocks (the endpoint table path), crossed and coinciding interpolants -- the general path of the
    colour search and its tie rules, which uniform random pixels almost never reach."""
    rng = np.random.default_rng(77)
    n = 2_000_000
    base = rng.integers(0, 256, (n, 1, 4), dtype=np.int16)
    spread = rng.choice(np.array([0, 1, 2, 3, 8, 40], np.int16), (n, 1, 1))
    palette = np.clip(base + rng.integers(-1, 2, (n, 4, 4), dtype=np.int16) * spread, 0, 255).astype(np.uint8)
    pick = rng.integers(0, 4, (n, 16)) % rng.integers(1, 5, (n, 1))
    blocks = np.take_along_axis(palette, pick[..., None].astype(np.intp), axis=1)          # (n, 16, 4)
    cols = 2000
    img = np.ascontiguousarray(blocks.reshape(n // cols, cols, 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(n // cols * 4, cols * 4, 4))
    h, w = img.shape[:2]
    want1 = ck.oracle_dxt1_rgba(img.ravel(), h, w)
    for vote in VOTES:
        assert np.array_equal(encode(emu, 0, 4, img.ravel(), h, w, vote=vote), want1), vote
    quarter = np.ascontiguousarray(img[: h // 4])
    assert np.array_equal(encode(emu, 1, 4, quarter.ravel(), h // 4, w), ck.oracle_dxt(ck.RGBA, quarter.ravel(), h // 4, w))


@pytest.mark.parametrize("strategy", [0, 1, 2, 3])
def test_etc1_soak_clamping_and_ties(emu, strategy):
    """Sixty thousand blocks per strategy near black and near white (every codeword clamps), on flat and two-level
    content (ties between codewords and between the two orientations: '<=' keeps the unflipped block)."""
    rng = np.random.default_rng(78 + strategy)
    n = 60_000
    centre = rng.choice(np.array([0, 3, 16, 60, 128, 200, 240, 252, 255], np.int16), (n, 1, 1))
    spread = rng.choice(np.array([0, 1, 2, 6, 20, 90], np.int16), (n, 1, 1))
    blocks = np.clip(centre + rng.integers(-1, 2, (n, 16, 3), dtype=np.int16) * spread + rng.integers(-2, 3, (n, 1, 3), dtype=np.int16), 0, 255)
    cols = 500
    img = np.ascontiguousarray(blocks.astype(np.uint8).reshape(n // cols, cols, 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(n // cols * 4, cols * 4, 3))
    h, w = img.shape[:2]
    want = ck.oracle_etc1(strategy, img.ravel(), h, w)
    for vote in (0, 1):  # line form where this block's codewords do not clamp / general form everywhere
        assert np.array_equal(encode(emu, 2, 3, img.ravel(), h, w, strategy=strategy, vote=vote), want), vote


def test_etc1_line_form_prefix_lengths(emu):
    """The codewords that take the line form are a prefix 0 .. n-1 decided by the WARP's smallest margin: whatever the
    other lanes contribute (every margin at which n changes, and the values next to it) the bytes are the same."""
    rng = np.random.default_rng(5)
    n = 6000
    centre = rng.integers(20, 236, (n, 1, 3), dtype=np.int16)
    spread = rng.choice(np.array([1, 4, 12, 40, 120], np.int16), (n, 1, 1))
    blocks = np.clip(centre + rng.integers(-1, 2, (n, 16, 3), dtype=np.int16) * spread + rng.integers(-3, 4, (n, 16, 3), dtype=np.int16), 0, 255)
    cols = 500
    img = np.ascontiguousarray(blocks.astype(np.uint8).reshape(n // cols, cols, 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(n // cols * 4, cols * 4, 3))
    h, w = img.shape[:2]
    want = ck.oracle_etc1(2, img.ravel(), h, w)
    for large in (8, 17, 29, 42, 60, 80, 106):
        for margin in (large - 1, large, large + 1):
            assert np.array_equal(encode(emu, 2, 3, img.ravel(), h, w, strategy=2, vote=1000 + margin + 128), want), margin


def test_device_code_addressing_under_sanitizers():
    """The same device code built with AddressSanitizer + UBSan and driven through the encoders that index memory in
    interesting ways: clamp-to-edge windows on ragged images with row padding, CompressAndPad grids, and PVRTC halo
    stripes that see nothing but a private copy of their own rows (an out-of-stripe read would leave that buffer)."""
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    ubsan = subprocess.run(["gcc", "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    if not (os.path.isabs(asan) and os.path.isabs(ubsan)):
        pytest.skip("no sanitizer runtimes in this toolchain")
    so = os.path.join(HERE, "libhostemu_asan.so")
    sources = [os.path.join(HERE, f) for f in ("hostemu.cc", "cuda_emulation.h")]
    sources += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in sources):
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize=alignment",
                        "-fno-sanitize-recover=undefined", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-o", so, os.path.join(HERE, "hostemu.cc")], check=True)
    script = r"""
import ctypes as C, sys
import numpy as np
lib = C.CDLL(sys.argv[1])
u8p = C.POINTER(C.c_uint8)
ptr = lambda a: a.ctypes.data_as(u8p)
lib.emu_encode4x4.argtypes = [C.c_int, C.c_int, u8p] + [C.c_uint32] * 5 + [C.c_int, C.c_int, u8p]
lib.emu_pvrtc2.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, u8p]
lib.emu_decode4x4.argtypes = [C.c_int, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u8p, C.c_uint32]
rng = np.random.default_rng(3)
for codec, nc in ((0, 3), (0, 4), (1, 4), (2, 3)):
    for h, w, padding, ch, cw in ((1, 1, 0, 1, 1), (5, 7, 3, 5, 7), (9, 13, 0, 16, 20), (33, 67, 5, 64, 128)):
        pitch = w * nc + padding
        src = rng.integers(0, 256, h * pitch - padding, dtype=np.uint8)   # the last row carries no padding
        out = np.zeros(((ch + 3) // 4) * ((cw + 3) // 4) * (16 if codec == 1 else 8), np.uint8)
        assert lib.emu_encode4x4(codec, nc, ptr(src), h, w, pitch, ch, cw, 0, 2, ptr(out)) == 0
        if (ch, cw) == (h, w):
            dec = np.zeros(h * w * (4 if codec == 1 else 3), np.uint8)
            assert lib.emu_decode4x4(codec, ptr(out), h, w, (w + 3) // 4, 0, ptr(dec), w * (4 if codec == 1 else 3)) == 0
for n in (8, 16, 64):
    img = rng.integers(0, 256, n * n * 4, dtype=np.uint8)
    for stripes in (1, 2, 3, 4, 5):
        out = np.zeros(n * n // 4, np.uint8)
        status = lib.emu_pvrtc2(ptr(img), n, n, stripes, ptr(out))
        assert status in (0, -1), status
print("clean")
"""
    env = dict(os.environ, LD_PRELOAD=asan + ":" + ubsan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0")
    run = subprocess.run([sys.executable, "-c", script, so], capture_output=True, text=True, env=env, timeout=600)
    assert run.returncode == 0 and run.stdout.strip().endswith("clean"), run.stderr[-3000:]
