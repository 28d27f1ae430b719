"""Kernel time of every encoder on different image content (random, smooth, gradients, dark, checkerboard, constant,
two-colour): the DXT index search has a warp-uniform fast path and a general path, constant blocks take a table path,
so speed may depend on the data; output is bit-exact either way.  Run on the GPU box:
python tools/bench_content.py [workload ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import imagegen  # noqa: E402
import image_compression_b200 as icb  # noqa: E402

n = 4096
cases = (("dxt1_rgba8", icb.CODEC_DXT1, icb.RGBA, 4), ("dxt1_rgb8", icb.CODEC_DXT1, icb.RGB, 3), ("dxt5_rgba8", icb.CODEC_DXT5, icb.RGBA, 4),
         ("etc1_rgb8", icb.CODEC_ETC1, icb.RGB, 3), ("pvrtc2_rgba8", icb.CODEC_PVRTC2, icb.RGBA, 4))
kinds = ("random", "smooth_noise", "gradient", "dark", "checker", "constant", "two_colour", "alpha_extremes", "photo_like", "photo_dark")


def photo_like(h, w, nc, lo, hi, seed):
    """Mid-frequency structure the way photographs and painted textures have it: a few sinusoids with periods of 40 to
    600 pixels per channel, mapped to [lo, hi], plus +-8 of noise -- not one of the parity kinds (tests/imagegen.py),
    only a timing input: ETC1's line form depends on how far a warp's 128 x 4 pixels stay from black and white."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    chans = []
    for k in range(nc):
        v = np.zeros((h, w), np.float32)
        for _ in range(5):
            period, angle, phase = rng.uniform(40, 600), rng.uniform(0, np.pi), rng.uniform(0, 2 * np.pi)
            v += np.sin((xx * np.cos(angle) + yy * np.sin(angle)) * (2 * np.pi / period) + phase)
        v = (v - v.min()) / (v.max() - v.min())
        chans.append(lo + v * (hi - lo) + rng.integers(-8, 9, (h, w)))
    return np.ascontiguousarray(np.clip(np.stack(chans, -1), 0, 255).astype(np.uint8))


def content(kind, nc):
    if kind == "photo_like":
        return photo_like(1024, 1024, nc, 30, 225, 11)
    if kind == "photo_dark":
        return photo_like(1024, 1024, nc, 0, 120, 12)
    return imagegen.make(kind, 1024, 1024, nc, seed=3)


if len(sys.argv) > 1:  # optional: only the workloads (and contents, "kind=dark") named on the command line
    names = [a for a in sys.argv[1:] if "=" not in a]
    cases = tuple(c for c in cases if not names or c[0] in names)
    kinds = tuple(a.split("=")[1] for a in sys.argv[1:] if a.startswith("kind=")) or kinds
for kind in kinds:
    row = {"content": kind}
    for name, codec, fmt, nc in cases:
        img = np.tile(content(kind, nc), (4, 4, 1))
        d = torch.from_numpy(np.ascontiguousarray(img).ravel()).cuda()
        out = torch.empty(icb.compressed_size(codec, n, n), dtype=torch.uint8, device="cuda")
        run = (lambda: icb.pvrtc_encode_device(d, n, n, out=out)) if codec == icb.CODEC_PVRTC2 else (lambda: icb.encode_device(codec, fmt, d, n, n, out=out))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        row[name + "_us"] = round(e0.elapsed_time(e1) * 100, 1)
    print(json.dumps(row))
