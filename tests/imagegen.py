"""Deterministic test images (numpy only).  Random bytes exercise few encoder branches, so
the structured kinds matter: constant, two-colour, gradients, narrow ranges, alpha extremes, all-zero channels."""
import zlib

import numpy as np

KINDS = ("random", "constant", "two_colour", "gradient", "narrow", "alpha_extremes", "zero_channel", "checker",
         "smooth_noise", "dark")


def make(kind, h, w, nc, seed=0):
    rng = np.random.default_rng(zlib.crc32(repr((kind, h, w, nc, seed)).encode()))
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "random":
        a = rng.integers(0, 256, (h, w, nc), dtype=np.uint8)
    elif kind == "constant":
        a = np.empty((h, w, nc), np.uint8)
        a[...] = rng.integers(0, 256, (nc,), dtype=np.uint8)
    elif kind == "two_colour":
        c = rng.integers(0, 256, (2, nc), dtype=np.uint8)
        a = c[rng.integers(0, 2, (h, w))]
    elif kind == "gradient":
        a = np.stack([(xx * (3 + k) + yy * (5 - k) + 37 * k + seed) % 256 for k in range(nc)], -1).astype(np.uint8)
    elif kind == "narrow":
        a = (rng.integers(0, 5, (h, w, nc)) + rng.integers(0, 250)).astype(np.uint8)
    elif kind == "alpha_extremes":
        a = rng.integers(0, 256, (h, w, nc), dtype=np.uint8)
        if nc == 4:
            a[..., 3] = rng.choice(np.array([0, 0, 255, 255, 1, 254, 128, 224, 223, 37], np.uint8), (h, w))
            a[: h // 2, : w // 2, 3] = rng.choice(np.array([0, 255], np.uint8), (h // 2, w // 2))
        else:
            a[..., 1] = rng.choice(np.array([0, 255], np.uint8), (h, w))
    elif kind == "zero_channel":
        a = rng.integers(0, 256, (h, w, nc), dtype=np.uint8)
        a[..., seed % nc] = 0
        a[: max(1, h // 2)] //= 32
    elif kind == "checker":
        cell = 1 << (seed % 4)
        v = ((((xx // cell) + (yy // cell)) & 1) * 255).astype(np.uint8)
        a = np.stack([v] * nc, -1)
    elif kind == "smooth_noise":
        base = np.stack([(np.sin(xx / 9.0 + k) * 60 + np.cos(yy / 7.0 - k) * 60 + 128) for k in range(nc)], -1)
        a = np.clip(base + rng.integers(-6, 7, (h, w, nc)), 0, 255).astype(np.uint8)
    elif kind == "dark":
        a = rng.integers(0, 3, (h, w, nc), dtype=np.uint8)
        a[rng.integers(0, h), rng.integers(0, w)] = 255
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(a)


def with_row_padding(img, padding, fill=0xA5):
    """Returns (flat buffer, pitch) with `padding` junk bytes after every row."""
    h, w, nc = img.shape
    pitch = w * nc + padding
    buf = np.full(h * pitch, fill, np.uint8)
    buf.reshape(h, pitch)[:, : w * nc] = img.reshape(h, w * nc)
    return buf, pitch
