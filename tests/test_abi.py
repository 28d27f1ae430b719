"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/icb200.h declares,
the Python prototype table matches the header, size arithmetic agrees with the reference rules, and compute entry
points fail loudly (never fall back) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import checkers as ck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "icb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"ICB_API[^;(]*?\b(icb_\w+)\s*\(", text)))


def test_header_declares_what_binding_uses(icb):
    from image_compression_b200 import binding
    assert header_functions() == sorted(binding.PROTOTYPES)


def test_library_exports_every_declared_symbol(icb):
    handle = C.CDLL(icb.lib_path())
    for name in header_functions():
        assert hasattr(handle, name), name
    assert icb.lib().icb_abi_version() == 1


def test_compressed_size_rules(icb):
    assert icb.compressed_size(icb.CODEC_DXT1, 8192, 8192) == 33554432
    assert icb.compressed_size(icb.CODEC_DXT5, 8192, 8192) == 67108864
    assert icb.compressed_size(icb.CODEC_ETC1, 4096, 4096) == 8388608
    assert icb.compressed_size(icb.CODEC_PVRTC2, 4096, 4096) == 4194304
    assert icb.compressed_size(icb.CODEC_DXT1, 5, 5) == 4 * 8
    assert icb.compressed_size(icb.CODEC_DXT1, 0, 5) == 0
    if ck.have_ref():
        for (h, w) in ((1, 1), (4, 4), (5, 9), (0, 3), (17, 33)):
            assert icb.compressed_size(icb.CODEC_DXT1, h, w) == ck.ref().icref_size(0, ck.RGB, h, w)
            assert icb.compressed_size(icb.CODEC_DXT5, h, w) == ck.ref().icref_size(0, ck.RGBA, h, w)
            assert icb.compressed_size(icb.CODEC_ETC1, h, w) == ck.ref().icref_size(1, ck.RGB, h, w)


def test_stripe_partition_covers_grid(icb):
    for rows in (1, 7, 8, 2048, 2049, 513):
        for world in (1, 2, 3, 4, 8):
            spans = [icb.stripe_rows(rows, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device failure mode")
def test_no_device_fails_loudly(icb):
    src = np.zeros(16 * 16 * 4, np.uint8)
    with pytest.raises(icb.IcbError) as e:
        icb.compress_host(icb.CODEC_DXT5, icb.RGBA, src, 16, 16)
    assert e.value.status == -2  # ICB_ERR_CUDA: no fallback path exists
    assert icb.lib().icb_device_count() < 0


def test_argument_validation_needs_no_device(icb):
    src = np.zeros(64, np.uint8)
    out = np.zeros(8, np.uint8)
    L = icb.lib()
    # null / zero dimension / format-codec mismatches are rejected before any CUDA call
    assert L.icb_compress_host(icb.CODEC_DXT1, icb.RGB, 0, 4, 0, 0, 0, 2, src.ctypes.data, out.ctypes.data, 8) == -1
    assert L.icb_compress_host(icb.CODEC_DXT1, icb.RGB, 4, 4, 0, 0, 0, 2, None, out.ctypes.data, 8) == -1
    assert L.icb_compress_host(icb.CODEC_ETC1, icb.RGBA, 4, 4, 0, 0, 0, 2, src.ctypes.data, out.ctypes.data, 8) == -1
    assert L.icb_compress_host(icb.CODEC_ETC1, icb.BGR, 4, 4, 0, 0, 0, 2, src.ctypes.data, out.ctypes.data, 8) == -1
    assert L.icb_compress_host(icb.CODEC_DXT5, icb.RGB, 4, 4, 0, 0, 0, 2, src.ctypes.data, out.ctypes.data, 16) == -1
    assert b"" != L.icb_last_error()


def test_host_emulation_macro_is_test_only():
    """ICB_HOST_EMULATION (device headers compiled for the CPU) is defined by tests/hostemu alone: no build recipe of
    the product sets it, and nothing in the package outside the guarded #ifdef branches mentions it."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    recipes = ["image_compression_b200/csrc/Makefile", "image_compression_b200/cpp/Makefile", "oracle/Makefile",
               "tools/build_variants.sh", "__graft_entry__.py", "bench.py", "image_compression_b200/binding.py"]
    for rel in recipes:
        assert "ICB_HOST_EMULATION" not in open(os.path.join(root, rel)).read(), rel
    definers = []
    for base, _, files in os.walk(root):
        if any(part in base for part in (".git", "gpurun_out", "__pycache__")):
            continue
        for f in files:
            if f.endswith((".h", ".cuh", ".cu", ".cc", ".c", ".cpp")):
                text = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"^\s*#\s*define\s+ICB_HOST_EMULATION", text, re.M):
                    definers.append(os.path.relpath(os.path.join(base, f), root))
    assert definers == ["tests/hostemu/cuda_emulation.h"], definers
