#!/usr/bin/env python3
"""Generates tests/golden/golden_ops_v1.npz: outputs of the UNMODIFIED reference (oracle/_ref/libicref.so) for the
compressed-domain operations -- Downsample, Pad, CopySubimage, CreateSolidImage, TranscodeDxt1ToEtc1.  Build
container only; the fixtures travel to the GPU box, the reference does not.  Each case stores the input block
stream, the call parameters and the reference's output bytes (empty when the reference returned false).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck  # noqa: E402
import imagegen  # noqa: E402

CODEC_FORMATS = ((0, ck.RGB), (0, ck.BGR), (1, ck.RGBA), (1, ck.BGRA), (2, ck.RGB))


def compress(codec, fmt, img, h, w, strategy=2):
    if codec == 2:
        return ck.ref_etc(strategy, img.ravel(), h, w)
    return ck.ref_dxt(fmt, img.ravel(), h, w)


def main():
    ck.build_oracle()
    assert ck.have_ref(), "reference not built: /root/reference must be mounted"
    rng = np.random.default_rng(2024)
    arrays, index = {}, []

    def add(op, codec, fmt, blocks, out, **kw):
        i = len(index)
        arrays["in_%d" % i] = np.ascontiguousarray(blocks)
        arrays["out_%d" % i] = out if out is not None else np.zeros(0, np.uint8)
        index.append(dict(op=op, codec=codec, format=fmt, refused=out is None, **kw))

    for codec, fmt in CODEC_FORMATS:
        nc = ck.ncomp(fmt)
        strategies = (2, 3) if codec == 2 else (2,)
        for st in strategies:
            for kind, (h, w) in (("smooth_noise", (32, 32)), ("random", (16, 24)), ("alpha_extremes", (8, 8)),
                                 ("two_colour", (13, 29)), ("gradient", (4, 16)), ("dark", (24, 4)), ("random", (4, 4)),
                                 ("smooth_noise", (2, 2)), ("random", (1, 4)), ("random", (4, 1)), ("random", (3, 4)),
                                 ("random", (12, 8)), ("constant", (8, 8))):
                blocks = compress(codec, fmt, imagegen.make(kind, h, w, nc, seed=5), h, w, st)
                out, _ = ck.ref_downsample(codec, fmt, blocks, h, w, strategy=st)
                add("downsample", codec, fmt, blocks, out, h=h, w=w, strategy=st, note=kind)
            blocks = rng.integers(0, 256, 4 * 4 * ck.block_bytes(codec), dtype=np.uint8)  # arbitrary bit patterns
            out, _ = ck.ref_downsample(codec, fmt, blocks, 16, 16, strategy=st)
            add("downsample", codec, fmt, blocks, out, h=16, w=16, strategy=st, note="random blocks")
            for (h, w, ph, pw) in ((8, 8, 17, 14), (5, 7, 8, 16), (16, 4, 20, 4), (12, 20, 12, 20), (4, 4, 16, 16)):
                blocks = compress(codec, fmt, imagegen.make("smooth_noise", h, w, nc, seed=6), h, w, st)
                out, _ = ck.ref_pad(codec, fmt, blocks, h, w, ph, pw, strategy=st)
                add("pad", codec, fmt, blocks, out, h=h, w=w, ph=ph, pw=pw, strategy=st)
            blocks = rng.integers(0, 256, 2 * 3 * ck.block_bytes(codec), dtype=np.uint8)
            out, _ = ck.ref_pad(codec, fmt, blocks, 8, 12, 20, 24, strategy=st)
            add("pad", codec, fmt, blocks, out, h=8, w=12, ph=20, pw=24, strategy=st, note="random blocks")
        blocks = compress(codec, fmt, imagegen.make("random", 16, 24, nc, seed=7), 16, 24)
        for (row, col, sh, sw) in ((4, 8, 8, 12), (0, 0, 16, 24), (12, 20, 4, 4), (2, 8, 8, 12), (12, 8, 8, 12)):
            out, _ = ck.ref_copy_subimage(codec, fmt, blocks, 16, 24, row, col, sh, sw)
            add("copy_subimage", codec, fmt, blocks, out, h=16, w=24, row=row, col=col, sh=sh, sw=sw)
        for _ in range(6):
            colour = rng.integers(0, 256, 4, dtype=np.uint8)
            out, _ = ck.ref_solid(codec, fmt, 9, 6, colour)
            add("solid", codec, fmt, colour, out, h=9, w=6)
    for kind in ("random", "smooth_noise", "constant", "two_colour", "dark", "gradient"):
        blocks = ck.ref_dxt(ck.RGB, imagegen.make(kind, 32, 24, 3, seed=8).ravel(), 32, 24)
        add("transcode", 0, ck.RGB, blocks, ck.ref_transcode(blocks), note=kind)
    blocks = rng.integers(0, 256, 8 * 256, dtype=np.uint8)
    add("transcode", 0, ck.RGB, blocks, ck.ref_transcode(blocks), note="random blocks")

    arrays["index_json"] = np.frombuffer(json.dumps(index).encode(), np.uint8)
    out_path = os.path.join(ROOT, "tests", "golden", "golden_ops_v1.npz")
    np.savez_compressed(out_path, **arrays)
    print("wrote %s: %d cases, %d bytes" % (out_path, len(index), os.path.getsize(out_path)))


if __name__ == "__main__":
    main()
