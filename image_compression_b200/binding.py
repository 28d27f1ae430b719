"""ctypes binding of include/icb200.h plus thin helpers over torch tensors (device memory and streams only)."""
import ctypes as C
import os

import numpy as np

RGB, BGR, RGBA, BGRA = 0, 1, 2, 3
CODEC_DXT1, CODEC_DXT5, CODEC_ETC1, CODEC_PVRTC2 = 0, 1, 2, 3
ETC_SPLIT_HORIZONTALLY, ETC_SPLIT_VERTICALLY, ETC_SMALLER_ERROR, ETC_HEURISTIC = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    # ICB200_LIB: another build of the same library (kernel A/B experiments, tools/build_variants.sh)
    return os.environ.get("ICB200_LIB") or os.path.join(_HERE, "lib", "libicb200.so")


class IcbError(RuntimeError):
    def __init__(self, status, text):
        super().__init__("icb status %d: %s" % (status, text))
        self.status = status


_lib = None

# name -> (restype, argtypes); mirrors include/icb200.h one to one (tests check the header against this table)
PROTOTYPES = {
    "icb_abi_version": (C.c_int, []),
    "icb_last_error": (C.c_char_p, []),
    "icb_device_count": (C.c_int, []),
    "icb_compressed_size": (C.c_size_t, [C.c_int, C.c_uint32, C.c_uint32]),
    "icb_dxt1_encode_rgb8": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]),
    "icb_dxt1_encode_rgba8": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]),
    "icb_dxt5_encode_rgba8": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]),
    "icb_etc1_encode_rgb8": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]),
    "icb_encode4x4_stripe": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "icb_pvrtc2_scratch_size": (C.c_size_t, [C.c_uint32, C.c_uint32]),
    "icb_pvrtc2_encode_rgba8": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "icb_pvrtc2_encode_stripe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "icb_compress_host": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "icb_decode4x4": (C.c_int, [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "icb_decompress_host": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "icb_downsample4x4": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "icb_pad4x4": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "icb_copy_subimage4x4": (C.c_int, [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "icb_fill_solid4x4": (C.c_int, [C.c_int, C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "icb_transcode_dxt1_to_etc1": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "icb_blockop_host": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "icb_device_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "icb_device_free": (C.c_int, [C.c_void_p]),
    "icb_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "icb_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "icb_ipc_close": (C.c_int, [C.c_void_p]),
    "icb_host_alloc": (C.c_void_p, [C.c_size_t]),
    "icb_host_free": (None, [C.c_void_p]),
    "icb_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "icb_host_unregister": (C.c_int, [C.c_void_p]),
    "icb_fill_synthetic": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_void_p]),
    "icb_launch_count": (C.c_uint64, []),
    "icb_set_tma_mode": (C.c_int, [C.c_int]),
    "icb_set_host_devices": (C.c_int, [C.c_int]),
    "icb_trim": (C.c_size_t, []),
    "icb_idle_pipes": (C.c_size_t, []),
    "icb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "icb_ctx_destroy": (C.c_int, [C.c_void_p]),
    "icb_ctx_device_count": (C.c_int, [C.c_void_p]),
    "icb_ctx_device": (C.c_int, [C.c_void_p, C.c_int]),
    "icb_ctx_peer_stores": (C.c_int, [C.c_void_p]),
    "icb_stripe_partition": (C.c_int, [C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]),
    "icb_root_share_permille": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "icb_encode_sharded": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_size_t, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "icb_ctx_compress_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
}


def lib():
    """Loads lib/libicb200.so.  Raises if it is missing -- there is nothing to fall back to."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a); there is no CPU fallback" % path)
        handle = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _check(status):
    if status != 0:
        raise IcbError(status, lib().icb_last_error().decode())


def compressed_size(codec, coded_h, coded_w):
    return lib().icb_compressed_size(codec, coded_h, coded_w)


def launch_count():
    return int(lib().icb_launch_count())


def set_tma_mode(mode):
    return lib().icb_set_tma_mode(mode)


def _stream_ptr(stream):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def ncomp_of(fmt):
    return 3 if fmt in (RGB, BGR) else 4


def encode_device(codec, fmt, src, h, w, pitch=None, coded_h=None, coded_w=None, strategy=ETC_SMALLER_ERROR,
                  out=None, stream=None):
    """Device-resident encode of a uint8 CUDA tensor `src` (flat or [h, pitch]) -> uint8 CUDA tensor of blocks.

    codec DXT1 with a 4-component format uses the RGBA8 extension entry point (alpha ignored)."""
    import torch
    nc = ncomp_of(fmt)
    pitch = w * nc if pitch is None else pitch
    coded_h = h if coded_h is None else coded_h
    coded_w = w if coded_w is None else coded_w
    size = compressed_size(codec, coded_h, coded_w)
    if out is None:
        out = torch.empty(size, dtype=torch.uint8, device=src.device)
    assert out.numel() == size and src.is_cuda and src.dtype == torch.uint8
    swap = 1 if fmt in (BGR, BGRA) else 0
    L = lib()
    sp = _stream_ptr(stream)
    with torch.cuda.device(src.device):
        if codec == CODEC_DXT1 and nc == 3:
            _check(L.icb_dxt1_encode_rgb8(src.data_ptr(), h, w, pitch, coded_h, coded_w, swap, out.data_ptr(), sp))
        elif codec == CODEC_DXT1:
            _check(L.icb_dxt1_encode_rgba8(src.data_ptr(), h, w, pitch, coded_h, coded_w, swap, out.data_ptr(), sp))
        elif codec == CODEC_DXT5:
            if nc != 4:
                raise IcbError(-1, "DXT5 needs a 4-component format")
            _check(L.icb_dxt5_encode_rgba8(src.data_ptr(), h, w, pitch, coded_h, coded_w, swap, out.data_ptr(), sp))
        elif codec == CODEC_ETC1:
            if fmt != RGB:
                raise IcbError(-1, "ETC1 supports kRGB only")
            _check(L.icb_etc1_encode_rgb8(src.data_ptr(), h, w, pitch, coded_h, coded_w, strategy, out.data_ptr(), sp))
        elif codec == CODEC_PVRTC2:
            return pvrtc_encode_device(src, h, w, out=out, stream=stream)
        else:
            raise IcbError(-1, "unknown codec")
    return out


from .sharding import stripe_rows  # noqa: E402,F401  (re-exported)


def encode_stripe_device(codec, fmt, src_base_ptr, h, w, pitch, coded_h, coded_w, r0, r1, out, strategy=ETC_SMALLER_ERROR,
                         stream=None):
    """Encodes block rows [r0, r1).  src_base_ptr is the (possibly virtual) address of pixel (0,0) of the whole
    image; only the rows the stripe reads must be resident.  `out` receives the stripe's blocks: a uint8 tensor, or
    a raw device address (int) -- e.g. a peer-mapped one from sharding.PeerStream."""
    swap = 1 if fmt in (BGR, BGRA) else 0
    out_ptr = out if isinstance(out, int) else out.data_ptr()
    _check(lib().icb_encode4x4_stripe(codec, ncomp_of(fmt), C.c_void_p(src_base_ptr), h, w, pitch, coded_h, coded_w, swap,
                                      strategy, r0, r1, C.c_void_p(out_ptr), _stream_ptr(stream)))
    return out


def pvrtc_encode_device(src, h, w, out=None, scratch=None, stream=None):
    import torch
    if out is None:
        out = torch.empty(w * h // 4, dtype=torch.uint8, device=src.device)
    with torch.cuda.device(src.device):
        _check(lib().icb_pvrtc2_encode_rgba8(src.data_ptr(), h, w, out.data_ptr(),
                                             scratch.data_ptr() if scratch is not None else None, _stream_ptr(stream)))
    return out


def pvrtc_encode_stripe_device(rows, first_pixel, h, w, r0, r1, out, scratch=None, stream=None):
    """Block rows [r0, r1) of an h x w image.  rows: uint8 tensor with image rows 4*(r0-1) .. 4*(r1+1)-1 (wrapped),
    first_pixel: 4-byte uint8 tensor holding pixel (0,0), out: the WHOLE image's block buffer (tensor or raw address)."""
    import torch
    with torch.cuda.device(rows.device):
        out_ptr = out if isinstance(out, int) else out.data_ptr()
        _check(lib().icb_pvrtc2_encode_stripe(rows.data_ptr(), first_pixel.data_ptr(), h, w, r0, r1, C.c_void_p(out_ptr),
                                              scratch.data_ptr() if scratch is not None else None, _stream_ptr(stream)))
    return out


def decode_device(codec, blocks, h, w, swap_rb=0, block_cols=None, out=None, stream=None):
    """Device-resident decode of a uint8 CUDA tensor of blocks -> uint8 CUDA tensor of h*w*(3|4) pixel bytes."""
    import torch
    nc = 4 if codec == CODEC_DXT5 else 3
    block_cols = (w + 3) // 4 if block_cols is None else block_cols
    if out is None:
        out = torch.empty(h * w * nc, dtype=torch.uint8, device=blocks.device)
    with torch.cuda.device(blocks.device):
        _check(lib().icb_decode4x4(codec, blocks.data_ptr(), h, w, block_cols, swap_rb, out.data_ptr(), w * nc, _stream_ptr(stream)))
    return out


def decompress_host(codec, fmt, blocks, h, w, block_cols=None):
    nc = 4 if codec == CODEC_DXT5 else 3
    block_cols = (w + 3) // 4 if block_cols is None else block_cols
    out = np.empty(h * w * nc, np.uint8)
    _check(lib().icb_decompress_host(codec, fmt, h, w, block_cols, blocks.ctypes.data, blocks.size, out.ctypes.data, out.size))
    return out


def fill_synthetic(dst, seed, byte_offset=0, stream=None):
    import torch
    with torch.cuda.device(dst.device):
        _check(lib().icb_fill_synthetic(dst.data_ptr(), dst.numel(), seed, byte_offset, _stream_ptr(stream)))
    return dst


def compress_host(codec, fmt, src, h, w, padded=None, padding=0, strategy=ETC_SMALLER_ERROR, out=None):
    """Host path: numpy uint8 in, numpy uint8 out, through icb_compress_host (H2D + kernels + D2H inside)."""
    ph, pw = padded if padded else (0, 0)
    coded_h, coded_w = max(h, ph), max(w, pw)
    size = compressed_size(codec, coded_h, coded_w) if codec != CODEC_PVRTC2 else w * h // 4
    if out is None:
        out = np.empty(size, np.uint8)
    _check(lib().icb_compress_host(codec, fmt, h, w, ph, pw, padding, strategy, src.ctypes.data, out.ctypes.data, out.size))
    return out


# ---- compressed-domain operations ---------------------------------------------

OP_DOWNSAMPLE, OP_PAD, OP_COPY_SUBIMAGE, OP_SOLID, OP_TRANSCODE = 0, 1, 2, 3, 4


def _nb(n):
    return (n + 3) // 4


def block_bytes(codec):
    return 16 if codec == CODEC_DXT5 else 8


def downsample_device(codec, blocks, h, w, strategy=ETC_SMALLER_ERROR, out=None, stream=None):
    """Blocks of an h x w image -> blocks of the ceil(h/2) x ceil(w/2) image (decode, 2x2 average, re-encode)."""
    import torch
    if out is None:
        out = torch.empty(_nb((h + 1) // 2) * _nb((w + 1) // 2) * block_bytes(codec), dtype=torch.uint8, device=blocks.device)
    with torch.cuda.device(blocks.device):
        _check(lib().icb_downsample4x4(codec, strategy, blocks.data_ptr(), h, w, out.data_ptr(), _stream_ptr(stream)))
    return out


def pad_device(codec, blocks, ch, cw, ph, pw, strategy=ETC_SMALLER_ERROR, out=None, stream=None):
    import torch
    if out is None:
        rows, cols = (_nb(ch), _nb(cw)) if (ch >= ph and cw >= pw) else (_nb(ph), _nb(pw))
        out = torch.empty(rows * cols * block_bytes(codec), dtype=torch.uint8, device=blocks.device)
    with torch.cuda.device(blocks.device):
        _check(lib().icb_pad4x4(codec, strategy, blocks.data_ptr(), ch, cw, ph, pw, out.data_ptr(), _stream_ptr(stream)))
    return out


def copy_subimage_device(codec, blocks, ch, cw, row, col, h, w, out=None, stream=None):
    import torch
    if out is None:
        out = torch.empty(_nb(h) * _nb(w) * block_bytes(codec), dtype=torch.uint8, device=blocks.device)
    with torch.cuda.device(blocks.device):
        _check(lib().icb_copy_subimage4x4(codec, blocks.data_ptr(), ch, cw, row, col, h, w, out.data_ptr(), _stream_ptr(stream)))
    return out


def fill_solid_device(codec, colour, h, w, device="cuda", out=None, stream=None):
    import torch
    if out is None:
        out = torch.empty(_nb(h) * _nb(w) * block_bytes(codec), dtype=torch.uint8, device=device)
    colour = bytes(list(colour) + [0] * (4 - len(colour)))
    with torch.cuda.device(out.device):
        _check(lib().icb_fill_solid4x4(codec, colour, h, w, out.data_ptr(), _stream_ptr(stream)))
    return out


def transcode_dxt1_to_etc1_device(blocks, stream=None):
    """In place."""
    import torch
    with torch.cuda.device(blocks.device):
        _check(lib().icb_transcode_dxt1_to_etc1(blocks.data_ptr(), blocks.numel() // 8, _stream_ptr(stream)))
    return blocks


def blockop_host(op, codec, args, src, out_size, strategy=ETC_SMALLER_ERROR):
    """Host-buffer form: numpy in, numpy out, through icb_blockop_host."""
    out = np.empty(out_size, np.uint8)
    arr = (C.c_uint32 * max(1, len(args)))(*args)
    src_ptr, src_size = (src.ctypes.data, src.size) if src is not None else (None, 0)
    _check(lib().icb_blockop_host(op, codec, strategy, arr, src_ptr, src_size, out.ctypes.data, out.size))
    return out


# ---- one image over several GPUs from one process (icb_ctx_*) -------------------------------

def stripe_partition(n, grid_rows, root_share_permille=-1):
    """Block-row splits [s0=0, s1, ..., sn=grid_rows] of icb_stripe_partition (no device needed)."""
    splits = (C.c_uint32 * (n + 1))()
    _check(lib().icb_stripe_partition(n, grid_rows, root_share_permille, splits))
    return [int(x) for x in splits]


def root_share_permille(codec, fmt, n):
    return int(lib().icb_root_share_permille(codec, ncomp_of(fmt), n))


class ShardContext:
    """icb_ctx: N GPUs of this process, device ids[0] the root that receives the packed stream by peer stores."""

    def __init__(self, device_ids=None, n=None):
        self._h = C.c_void_p()
        ids = list(device_ids) if device_ids is not None else None
        count = len(ids) if ids is not None else (0 if n is None else n)
        arr = (C.c_int * count)(*ids) if ids is not None else None
        _check(lib().icb_ctx_create(count, arr, C.byref(self._h)))
        self.devices = [lib().icb_ctx_device(self._h, k) for k in range(lib().icb_ctx_device_count(self._h))]

    @property
    def peer_stores(self):
        return bool(lib().icb_ctx_peer_stores(self._h))

    def encode(self, codec, fmt, stripes, h, w, splits, out, pitch=None, strategy=ETC_SMALLER_ERROR, stream=None):
        """stripes[r]: uint8 CUDA tensor on device r of the context holding pixel rows 4*splits[r] .. of the image;
        out: uint8 CUDA tensor on the root for the whole stream.  Asynchronous on `stream` (a stream of the root)."""
        import torch
        pitch = w * ncomp_of(fmt) if pitch is None else pitch
        n = len(self.devices)
        ptrs = (C.c_void_p * n)(*[(t.data_ptr() if t is not None and t.numel() else None) for t in stripes])
        sp = (C.c_uint32 * (n + 1))(*splits)
        with torch.cuda.device(self.devices[0]):
            _check(lib().icb_encode_sharded(self._h, codec, fmt, h, w, pitch, strategy, sp, ptrs, out.data_ptr(), _stream_ptr(stream)))
        return out

    def compress_host(self, codec, fmt, src, h, w, padded=None, padding=0, strategy=ETC_SMALLER_ERROR, out=None):
        ph, pw = padded if padded else (0, 0)
        size = compressed_size(codec, max(h, ph), max(w, pw)) if codec != CODEC_PVRTC2 else w * h // 4
        if out is None:
            out = np.empty(size, np.uint8)
        _check(lib().icb_ctx_compress_host(self._h, codec, fmt, h, w, ph, pw, padding, strategy, src.ctypes.data, out.ctypes.data, out.size))
        return out

    def close(self):
        if self._h:
            lib().icb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
