#!/usr/bin/env python3
"""End-to-end icb_compress_host from PINNED vs ordinary PAGEABLE host memory (what a C++ caller that malloc'ed its image
hands to Compress()).  DXT1 8192x8192 RGBA8.  Prints one JSON line per case.  ICB_HOST_DEVICES=N spreads the call
over N GPUs of the box (each chunk over its own PCIe link)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import image_compression_b200 as icb  # noqa: E402

n = 8192
in_bytes, out_bytes = n * n * 4, n * n // 2
L = icb.lib()
rng = np.random.default_rng(1)
pageable_in = rng.integers(0, 256, in_bytes, dtype=np.uint8)
pageable_out = np.zeros(out_bytes, np.uint8)
pin_in_p, pin_out_p = L.icb_host_alloc(in_bytes), L.icb_host_alloc(out_bytes)
pin_in = np.ctypeslib.as_array(C.cast(pin_in_p, C.POINTER(C.c_uint8)), shape=(in_bytes,))
pin_out = np.ctypeslib.as_array(C.cast(pin_out_p, C.POINTER(C.c_uint8)), shape=(out_bytes,))
pin_in[:] = pageable_in
for name, src, dst in (("pinned", pin_in, pin_out), ("pageable", pageable_in, pageable_out), ("pageable_in_pinned_out", pageable_in, pin_out)):
    for _ in range(2):
        icb.compress_host(icb.CODEC_DXT1, icb.RGBA, src, n, n, out=dst)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        icb.compress_host(icb.CODEC_DXT1, icb.RGBA, src, n, n, out=dst)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    print(json.dumps({"host_memory": name, "host_devices": os.environ.get("ICB_HOST_DEVICES", "1"), "ms": ms, "mpix_s": n * n / ms / 1e3, "h2d_gb_s": in_bytes / ms / 1e6,
                      "same_output": bool(np.array_equal(dst, pin_out))}))
