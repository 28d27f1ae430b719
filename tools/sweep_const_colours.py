#!/usr/bin/env python3
"""EXHAUSTIVE check of the constant-colour path of the DXT encoders (csrc/dxt_encode.cuh: dxt_const_colour and its
compile-time error table) on the CPU: every one of the 2^24 colours as a constant 4x4 block, DXT1 (three-colour halves
allowed) and DXT5 (always four colours), RGB and BGR order, device code under host emulation (tests/hostemu) against
the plain-C oracle.  ~2 minutes; run after touching the table or the search:  python tools/sweep_const_colours.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck  # noqa: E402
import test_hostemu as th  # noqa: E402

emu = next(th.emu.__wrapped__())
chunk = 1 << 19
bad = 0
for start in range(0, 1 << 24, chunk):
    colours = np.arange(start, start + chunk, dtype=np.uint32)
    rgb = np.stack([colours & 255, (colours >> 8) & 255, colours >> 16], -1).astype(np.uint8)      # (chunk, 3)
    row = np.repeat(rgb, 4, axis=0)                                                                # 4 pixels per block, x direction
    img3 = np.ascontiguousarray(np.broadcast_to(row[None], (4, chunk * 4, 3)))
    img4 = np.ascontiguousarray(np.concatenate([img3, np.full((4, chunk * 4, 1), 200, np.uint8)], -1))
    for swap, fmt3, fmt4 in ((0, ck.RGB, ck.RGBA), (1, ck.BGR, ck.BGRA)):
        got = th.encode(emu, 0, 3, img3.ravel(), 4, chunk * 4, swap=swap)
        bad += int(np.count_nonzero(got != ck.oracle_dxt(fmt3, img3.ravel(), 4, chunk * 4)))
        got = th.encode(emu, 1, 4, img4.ravel(), 4, chunk * 4, swap=swap)
        bad += int(np.count_nonzero(got != ck.oracle_dxt(fmt4, img4.ravel(), 4, chunk * 4)))
    print("colours %08x..%08x: %d differing bytes so far" % (start, start + chunk - 1, bad), flush=True)
print("PASS" if bad == 0 else "FAIL", bad)
sys.exit(1 if bad else 0)
