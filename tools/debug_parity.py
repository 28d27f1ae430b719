"""Diagnoses a parity failure of a 4x4 encoder build on the GPU box: encodes the bench image several times with the
library named by ICB200_LIB, byte-compares every run with the CPU reference and prints where the differences are
(block index, tile, which of the block's bytes).  python tools/debug_parity.py <workload> [runs]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck  # noqa: E402
import image_compression_b200 as icb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "dxt5_rgba8"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
codec, fmt, nc, n = {"dxt5_rgba8": (icb.CODEC_DXT5, icb.RGBA, 4, 8192), "dxt1_rgba8": (icb.CODEC_DXT1, icb.RGBA, 4, 8192),
                     "dxt1_rgb8": (icb.CODEC_DXT1, icb.RGB, 3, 8192)}[wl]
src = ck.synthetic(n * n * nc, 12345)
want, kind = ck.cpu_encode_full(wl, src, n, n)
d = torch.from_numpy(src).cuda()
bb = icb.block_bytes(codec)
out = torch.empty(want.size, dtype=torch.uint8, device="cuda")
prev = None
for r in range(runs):
    out.zero_()
    for _ in range(1 + r % 3):  # back-to-back launches (programmatic dependent launch) on some runs
        icb.encode_device(codec, fmt, d, n, n, out=out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    diff = np.flatnonzero(got != want)
    blocks = np.unique(diff // bb)
    print("run %d: %d differing bytes in %d blocks (reference: %s)" % (r, diff.size, blocks.size, kind))
    if blocks.size:
        cols = n // 4
        by, bx = blocks // cols, blocks % cols
        print("   block rows %d..%d cols %d..%d; tiles (x/64, y/4): %s" % (by.min(), by.max(), bx.min(), bx.max(),
              sorted(set(zip((bx // 64).tolist(), (by // 4).tolist())))[:12]))
        print("   byte-in-block histogram:", np.bincount(diff % bb, minlength=bb).tolist())
        b0 = int(blocks[0])
        print("   first block %d: got %s want %s" % (b0, got[b0 * bb:(b0 + 1) * bb].tobytes().hex(), want[b0 * bb:(b0 + 1) * bb].tobytes().hex()))
        if prev is not None:
            print("   same set of bytes as the previous run:", np.array_equal(prev, diff))
    prev = diff
