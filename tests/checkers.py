"""ctypes doors to the two CPU checkers (TEST INFRASTRUCTURE): oracle/liboracle.so (plain-C restatement) and,
when it was built in the container that has /root/reference, oracle/_ref/libicref.so (the unmodified reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
RGB, BGR, RGBA, BGRA = 0, 1, 2, 3
ETC_SPLIT_H, ETC_SPLIT_V, ETC_SMALLER_ERROR, ETC_HEURISTIC = 0, 1, 2, 3

_u8p = C.POINTER(C.c_uint8)


def _ptr(a):
    return a.ctypes.data_as(_u8p)


def build_oracle():
    """(Re)builds liboracle.so and, if the reference is mounted, _ref/libicref.so."""
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True, stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        for name in ("orc_dxt_compress", "orc_dxt1_compress_rgba", "orc_etc1_compress"):
            getattr(lib, name).restype = C.c_size_t
            getattr(lib, name).argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u8p]
        lib.orc_pvrtc2_compress.restype = C.c_size_t
        lib.orc_pvrtc2_compress.argtypes = [C.c_uint32, C.c_uint32, _u8p, _u8p]
        lib.orc_decode4x4.restype = None
        lib.orc_decode4x4.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u8p]
        lib.orc_downsample.restype = C.c_size_t
        lib.orc_downsample.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, _u8p, _u8p]
        lib.orc_pad.restype = C.c_size_t
        lib.orc_pad.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u8p]
        lib.orc_solid_block.restype = None
        lib.orc_solid_block.argtypes = [C.c_int, _u8p, _u8p]
        lib.orc_transcode_dxt1_to_etc1.restype = None
        lib.orc_transcode_dxt1_to_etc1.argtypes = [_u8p, C.c_size_t]
        lib.orc_fill_synthetic.restype = None
        lib.orc_fill_synthetic.argtypes = [_u8p, C.c_size_t, C.c_uint64, C.c_uint64]
        lib.orc_fnv1a64.restype = C.c_uint64
        lib.orc_fnv1a64.argtypes = [_u8p, C.c_size_t]
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libicref.so"))


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libicref.so"))
        u32p = C.POINTER(C.c_uint32)
        lib.icref_dxt.restype = C.c_long
        lib.icref_dxt.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t, u32p]
        lib.icref_etc.restype = C.c_long
        lib.icref_etc.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t, u32p]
        lib.icref_pvrtc.restype = C.c_long
        lib.icref_pvrtc.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t, u32p]
        lib.icref_dxt_external.restype = C.c_int
        lib.icref_dxt_external.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t]
        lib.icref_etc_external.restype = C.c_int
        lib.icref_etc_external.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t]
        lib.icref_decompress.restype = C.c_long
        lib.icref_decompress.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, _u8p, C.c_size_t, _u8p, C.c_size_t]
        lib.icref_block_op.restype = C.c_long
        lib.icref_block_op.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_uint] * 6 + [_u8p, C.c_size_t, _u8p, C.c_size_t, u32p]
        lib.icref_solid.restype = C.c_long
        lib.icref_solid.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint, _u8p, _u8p, C.c_size_t, u32p]
        lib.icref_transcode.restype = None
        lib.icref_transcode.argtypes = [_u8p, C.c_size_t]
        lib.icref_size.restype = C.c_size_t
        lib.icref_size.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint]
        _ref = lib
    return _ref


def ncomp(fmt):
    return 3 if fmt in (RGB, BGR) else 4


def nblocks(n):
    return (n + 3) // 4


# ---- plain-C oracle -------------------------------------------------------------------------------------------

def oracle_dxt(fmt, img, h, w, coded_h=None, coded_w=None, padding=0):
    coded_h = h if coded_h is None else max(h, coded_h)
    coded_w = w if coded_w is None else max(w, coded_w)
    out = np.zeros(nblocks(coded_h) * nblocks(coded_w) * (8 if ncomp(fmt) == 3 else 16), np.uint8)
    n = oracle().orc_dxt_compress(fmt, h, w, coded_h, coded_w, padding, _ptr(img), _ptr(out))
    assert n == out.size
    return out


def oracle_dxt1_rgba(img, h, w, swap_rb=0, coded_h=None, coded_w=None, padding=0):
    coded_h = h if coded_h is None else max(h, coded_h)
    coded_w = w if coded_w is None else max(w, coded_w)
    out = np.zeros(nblocks(coded_h) * nblocks(coded_w) * 8, np.uint8)
    n = oracle().orc_dxt1_compress_rgba(swap_rb, h, w, coded_h, coded_w, padding, _ptr(img), _ptr(out))
    assert n == out.size
    return out


def oracle_etc1(strategy, img, h, w, coded_h=None, coded_w=None, padding=0):
    coded_h = h if coded_h is None else max(h, coded_h)
    coded_w = w if coded_w is None else max(w, coded_w)
    out = np.zeros(nblocks(coded_h) * nblocks(coded_w) * 8, np.uint8)
    n = oracle().orc_etc1_compress(strategy, h, w, coded_h, coded_w, padding, _ptr(img), _ptr(out))
    assert n == out.size
    return out


def oracle_pvrtc(img, h, w):
    out = np.zeros(h * w // 4, np.uint8)
    n = oracle().orc_pvrtc2_compress(h, w, _ptr(img), _ptr(out))
    assert n == out.size
    return out


def oracle_decode(codec, blocks, h, w, swap_rb=0, block_cols=None):
    """codec: 0 DXT1 -> RGB888, 1 DXT5 -> RGBA8888, 2 ETC1 -> RGB888."""
    block_cols = nblocks(w) if block_cols is None else block_cols
    out = np.zeros(h * w * (4 if codec == 1 else 3), np.uint8)
    oracle().orc_decode4x4(codec, swap_rb, h, w, block_cols, _ptr(blocks), _ptr(out))
    return out


def ref_decompress(codec, fmt, blocks, h, w):
    """The reference's Decompress() on a block stream of an h x w image (codec 0/1 = DxtcCompressor, 2 = ETC)."""
    out = np.zeros(h * w * 4 + 16, np.uint8)
    n = ref().icref_decompress(codec, 2, fmt, h, w, _ptr(blocks), blocks.size, _ptr(out), out.size)
    return out[:n].copy() if n > 0 else None


def block_bytes(codec):
    return 16 if codec == 1 else 8


def oracle_downsample(codec, blocks, uh, uw, strategy=ETC_SMALLER_ERROR):
    """Compressed-domain 2:1 downsample of the blocks of a uh x uw image; None where the reference refuses."""
    out = np.zeros(max(1, nblocks((uh + 1) // 2)) * max(1, nblocks((uw + 1) // 2)) * block_bytes(codec), np.uint8)
    blocks = np.ascontiguousarray(blocks)
    n = oracle().orc_downsample(codec, strategy, uh, uw, _ptr(blocks), _ptr(out))
    return out[:n].copy() if n else None


def oracle_pad(codec, blocks, ch, cw, ph, pw, strategy=ETC_SMALLER_ERROR):
    rows, cols = max(nblocks(ch), nblocks(ph)), max(nblocks(cw), nblocks(pw))
    out = np.zeros(rows * cols * block_bytes(codec), np.uint8)
    blocks = np.ascontiguousarray(blocks)
    n = oracle().orc_pad(codec, strategy, ch, cw, ph, pw, _ptr(blocks), _ptr(out))
    return out[:n].copy()


def oracle_solid_block(codec, color):
    out = np.zeros(block_bytes(codec), np.uint8)
    color = np.ascontiguousarray(np.asarray(list(color) + [0] * (4 - len(color)), np.uint8))
    oracle().orc_solid_block(codec, _ptr(color), _ptr(out))
    return out


def oracle_transcode(blocks):
    out = np.ascontiguousarray(blocks).copy()
    oracle().orc_transcode_dxt1_to_etc1(_ptr(out), out.size // 8)
    return out


def synthetic(nbytes, seed, offset=0):
    out = np.zeros(nbytes, np.uint8)
    oracle().orc_fill_synthetic(_ptr(out), nbytes, seed, offset)
    return out


def fnv1a64(a):
    a = np.ascontiguousarray(a, np.uint8)
    return int(oracle().orc_fnv1a64(_ptr(a), a.size))


# ---- compiled reference ---------------------------------------------------------------------------------------

def _meta_dict(meta):
    return dict(format=int(meta[0]), uncompressed_height=int(meta[1]), uncompressed_width=int(meta[2]),
                compressed_height=int(meta[3]), compressed_width=int(meta[4]), padding_bytes_per_row=int(meta[5]),
                name_len=int(meta[6]))


def _run_ref(fn, args, cap):
    out = np.zeros(max(cap, 16), np.uint8)
    meta = (C.c_uint32 * 7)()
    n = fn(*args, _ptr(out), out.size, meta)
    if n <= 0:
        return None, None
    return out[:n].copy(), _meta_dict(meta)


def ref_dxt(fmt, img, h, w, padded=None, padding=0, want_meta=False):
    ph, pw = padded if padded else (0, 0)
    ch, cw = (max(h, ph), max(w, pw)) if padded else (h, w)
    cap = nblocks(ch) * nblocks(cw) * 16
    out, meta = _run_ref(ref().icref_dxt, (fmt, h, w, 1 if padded else 0, ph, pw, padding, _ptr(img)), cap)
    return (out, meta) if want_meta else out


def ref_etc(strategy, img, h, w, padded=None, padding=0, fmt=RGB, want_meta=False):
    ph, pw = padded if padded else (0, 0)
    ch, cw = (max(h, ph), max(w, pw)) if padded else (h, w)
    cap = nblocks(ch) * nblocks(cw) * 8
    out, meta = _run_ref(ref().icref_etc, (strategy, fmt, h, w, 1 if padded else 0, ph, pw, padding, _ptr(img)), cap)
    return (out, meta) if want_meta else out


def ref_pvrtc(img, h, w, padding=0, fmt=RGBA, want_meta=False):
    out, meta = _run_ref(ref().icref_pvrtc, (fmt, h, w, padding, _ptr(img)), h * w // 4 + 16)
    return (out, meta) if want_meta else out


def _ref_block_op(op, codec, fmt, blocks, h, w, a=0, b=0, c=0, d=0, strategy=ETC_SMALLER_ERROR, cap=None):
    blocks = np.ascontiguousarray(blocks)
    cap = cap if cap is not None else blocks.size * 4 + 64
    return _run_ref(ref().icref_block_op, (op, codec, strategy, fmt, h, w, a, b, c, d, _ptr(blocks), blocks.size), cap)


def ref_downsample(codec, fmt, blocks, h, w, strategy=ETC_SMALLER_ERROR):
    """The reference's Downsample() on the block stream Compress() made for an h x w image -> (bytes, meta)."""
    return _ref_block_op(0, codec, fmt, blocks, h, w, strategy=strategy)


def ref_pad(codec, fmt, blocks, h, w, ph, pw, strategy=ETC_SMALLER_ERROR):
    cap = max(nblocks(h), nblocks(ph)) * max(nblocks(w), nblocks(pw)) * 16 + 64
    return _ref_block_op(1, codec, fmt, blocks, h, w, ph, pw, strategy=strategy, cap=cap)


def ref_copy_subimage(codec, fmt, blocks, h, w, row, col, sh, sw):
    return _ref_block_op(2, codec, fmt, blocks, h, w, row, col, sh, sw)


def ref_solid(codec, fmt, h, w, color):
    color = np.ascontiguousarray(np.asarray(list(color) + [0] * (4 - len(color)), np.uint8))
    return _run_ref(ref().icref_solid, (codec, fmt, h, w, _ptr(color)), nblocks(h) * nblocks(w) * 16 + 64)


def ref_transcode(blocks):
    out = np.ascontiguousarray(blocks).copy()
    ref().icref_transcode(_ptr(out), out.size)
    return out


# ---- whole-image CPU answers at benchmark sizes ------------------------------------------------------------------

def cpu_encode_full(workload, img, h, w, threads=None):
    """The CPU answer for a whole BASELINE-size image in about a second: the unmodified reference (oracle/_ref) when it
    was built, else the oracle port, run on T row stripes from T threads, each writing its slice of the output through
    an external-storage CompressedImage -- byte-identical to the one-thread result because
    4x4 blocks are independent (tests/test_oracle.py checks that claim on small images).  PVRTC cannot be split
    (toroidal wrap, Z-order): one thread.  workload: dxt1_rgb8 | dxt1_rgba8 | dxt5_rgba8 | etc1_rgb8 | pvrtc2_rgba8;
    `img` is the image in the workload's own pixel format (dxt1_rgba8: RGBA, alpha stripped here for the reference,
    which has no such entry).  Returns (blocks, kind) with kind "reference" or "port"."""
    import threading
    use_ref = have_ref()
    kind = "reference" if use_ref else "port"
    img = np.ascontiguousarray(img).reshape(-1)
    if workload == "pvrtc2_rgba8":
        return (ref_pvrtc(img, h, w) if use_ref else oracle_pvrtc(img, h, w)), kind
    assert h % 4 == 0 and w % 4 == 0, "stripe split is for block-aligned sizes"
    nc_in = 3 if workload in ("dxt1_rgb8", "etc1_rgb8") else 4
    nc = nc_in
    if workload == "dxt1_rgba8":
        img = np.ascontiguousarray(img.reshape(-1, 4)[:, :3]).reshape(-1)
        nc = 3
    bb = 16 if workload == "dxt5_rgba8" else 8
    threads = threads or min(os.cpu_count() or 1, 64)
    grid_rows = h // 4
    threads = max(1, min(threads, grid_rows))
    out = np.zeros(grid_rows * (w // 4) * bb, np.uint8)
    bounds = [grid_rows * t // threads for t in range(threads + 1)]
    errors = []

    def job(t):
        r0, r1 = bounds[t], bounds[t + 1]
        if r1 == r0:
            return
        rows = (r1 - r0) * 4
        s = img[r0 * 4 * w * nc:r1 * 4 * w * nc]
        o = out[r0 * (w // 4) * bb:r1 * (w // 4) * bb]
        if use_ref:
            if workload == "etc1_rgb8":
                ok = ref().icref_etc_external(ETC_SMALLER_ERROR, rows, w, 0, _ptr(s), _ptr(o), o.size)
            else:
                ok = ref().icref_dxt_external(RGB if nc == 3 else RGBA, rows, w, 0, _ptr(s), _ptr(o), o.size)
            if ok != 1:
                errors.append(t)
        elif workload == "etc1_rgb8":
            oracle().orc_etc1_compress(ETC_SMALLER_ERROR, rows, w, rows, w, 0, _ptr(s), _ptr(o))
        else:
            oracle().orc_dxt_compress(RGB if nc == 3 else RGBA, rows, w, rows, w, 0, _ptr(s), _ptr(o))

    ts = [threading.Thread(target=job, args=(t,)) for t in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, "reference refused stripes %s" % errors
    return out, kind
