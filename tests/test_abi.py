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
    assert icb.lib().icb_abi_version() == 2


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


def test_c_abi_stripe_partition(icb):
    """icb_stripe_partition: contiguous cover of the grid, whole tile rows (4 block rows) per stripe where the grid allows,
    even split == sizes within one unit, root-heavy split gives the root at least its share and splits the rest evenly."""
    for rows in (1, 7, 8, 31, 2048, 2049, 513, 1024):
        for n in (1, 2, 3, 4, 8):
            s = icb.stripe_partition(n, rows)
            assert s[0] == 0 and s[-1] == rows and all(a <= b for a, b in zip(s, s[1:])) and len(s) == n + 1
            unit = 4 if rows >= 4 * n else 1
            assert all(x % unit == 0 for x in s[:-1])
            sizes = [b - a for a, b in zip(s, s[1:])]
            assert max(sizes) - min(sizes) <= unit + rows % unit
            for share in (0, 300, 475, 900, 1000):
                t = icb.stripe_partition(n, rows, share)
                assert t[0] == 0 and t[-1] == rows and all(a <= b for a, b in zip(t, t[1:]))
                if n > 1 and rows >= 4 * n:
                    assert t[1] >= max(s[1], (rows // 4 * share + 500) // 1000 * 4) - 4
                    rest = [b - a for a, b in zip(t[1:], t[2:])]
                    assert max(rest[:-1] or [0]) - min(rest[:-1] or [0]) <= 4
    assert icb.stripe_partition(8, 2048, 475)[1] == 972
    # B200 balance points: DXT1/DXT5 are NVLink-ingress bound from N = 4 up (root-heavy), ETC1 never is (even split)
    assert icb.root_share_permille(icb.CODEC_DXT1, icb.RGBA, 1) == 1000
    assert icb.root_share_permille(icb.CODEC_DXT1, icb.RGBA, 2) == -1 or icb.root_share_permille(icb.CODEC_DXT1, icb.RGBA, 2) >= 500
    assert 400 <= icb.root_share_permille(icb.CODEC_DXT1, icb.RGBA, 8) <= 600
    assert 400 <= icb.root_share_permille(icb.CODEC_DXT5, icb.RGBA, 8) <= 650
    assert icb.root_share_permille(icb.CODEC_ETC1, icb.RGB, 8) == -1


def test_misaligned_block_pointers_are_rejected_before_any_launch(icb):
    """Blocks move as 8/16-byte vectors: a misaligned stream pointer is ICB_ERR_INVALID, not a sticky device fault."""
    L = icb.lib()
    base = 0x7f0000000000
    assert L.icb_dxt1_encode_rgba8(base, 16, 16, 64, 16, 16, 0, base + 4, None) == -1
    assert L.icb_dxt5_encode_rgba8(base, 16, 16, 64, 16, 16, 0, base + 8, None) == -1
    assert L.icb_encode4x4_stripe(icb.CODEC_ETC1, 3, base, 16, 16, 48, 16, 16, 0, 2, 0, 4, base + 2, None) == -1
    assert L.icb_decode4x4(icb.CODEC_DXT1, base + 4, 16, 16, 4, 0, base, 48, None) == -1
    assert L.icb_downsample4x4(icb.CODEC_DXT5, 2, base, 16, 16, base + 8, None) == -1
    assert L.icb_pvrtc2_encode_rgba8(base, 8, 8, base + 4, None, None) == -1
    assert b"align" in L.icb_last_error()


def test_pipe_pool_entry_points_need_no_device(icb):
    L = icb.lib()
    assert L.icb_idle_pipes() == 0
    assert L.icb_trim() == 0
    assert L.icb_set_host_devices(-1) == 0x7fffffff  # nothing was set before: "no previous setting"


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


def test_library_sass_is_free_of_a_known_ptxas_miscompile():
    """CUDA 12.9's ptxas has been seen to split a lane-wise add with an immediate addend (VIADD.16x2 R, R, imm) into
    "VIADD.16x2 R17,R4,0x0 ; VIADD.16x2 R15,R4,imm ; PRMT R19,R17,0x7610,R15" -- one lane loses its addend -- in some
    instantiations of the DXT5 alpha statistics (found on the GPU by the parity tests; profiles/r02b_driver_ab.txt).
    The addends now live in constant memory; this scans every kernel of the shipped library for the signature, so that
    the build container already refuses a library whose GPU results would be wrong."""
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from image_compression_b200 import binding
    sass = subprocess.run([cuobjdump, "-sass", binding.lib_path()], capture_output=True, text=True, check=True).stdout
    assert sass.count("Function :") > 20
    bad, fn = [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
        elif re.search(r"VIADD\.16x2 R\d+, R\d+(\.reuse)?, 0x0 ", line):
            bad.append(fn)
    assert not bad, bad
