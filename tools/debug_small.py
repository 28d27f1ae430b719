"""GPU-box diagnosis: small images of every content kind through the device path, compared with the oracle block by block.
python tools/debug_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck  # noqa: E402
import imagegen  # noqa: E402
import image_compression_b200 as icb  # noqa: E402

for codec, fmt in ((0, ck.RGB), (0, ck.RGBA), (0, ck.BGRA), (1, ck.RGBA)):
    nc = ck.ncomp(fmt)
    bb = 16 if codec == 1 else 8
    for mode in (0, 2):  # generic only / automatic (TMA where possible)
        icb.set_tma_mode(mode)
        for kind in imagegen.KINDS:
            h, w = 64, 512
            img = imagegen.make(kind, h, w, nc, seed=5)
            want = ck.oracle_dxt(fmt, img.ravel(), h, w) if codec == 1 or nc == 3 else ck.oracle_dxt1_rgba(img.ravel(), h, w, swap_rb=1 if fmt == ck.BGRA else 0)
            got = icb.encode_device(codec, fmt, torch.from_numpy(img.ravel()).cuda(), h, w).cpu().numpy()
            if want.size != got.size:
                print("size mismatch", codec, fmt, kind, want.size, got.size)
                continue
            diff = np.flatnonzero(got != want)
            blocks = np.unique(diff // bb)
            line = "codec %d fmt %d mode %d %-14s: %5d of %d blocks differ" % (codec, fmt, mode, kind, blocks.size, want.size // bb)
            if blocks.size:
                b = int(blocks[0])
                by, bx = divmod(b, w // 4)
                line += "  bytes %s | first block %d got %s want %s px %s" % (
                    np.bincount(diff % bb, minlength=bb).tolist(), b, got[b * bb:(b + 1) * bb].tobytes().hex(),
                    want[b * bb:(b + 1) * bb].tobytes().hex(), img[4 * by:4 * by + 4, 4 * bx:4 * bx + 4].reshape(16, nc).tolist())
            print(line)
icb.set_tma_mode(2)
