#!/usr/bin/env python3
"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference (oracle/_ref/libicref.so, built by
oracle/Makefile from /root/reference).  Runs only in the build container; the fixtures travel, the reference does
not.  Each case stores the exact input bytes, the call parameters and the reference's output bytes.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck  # noqa: E402
import imagegen  # noqa: E402


def main():
    ck.build_oracle()
    assert ck.have_ref(), "reference not built: /root/reference must be mounted"
    arrays, index = {}, []

    def add(codec, fmt, img, h, w, padded=None, padding=0, strategy=2, note=""):
        nc = ck.ncomp(fmt)
        if padding:
            buf, _ = imagegen.with_row_padding(img, padding)
        else:
            buf = np.ascontiguousarray(img).reshape(-1)
        if codec == "dxt":
            out = ck.ref_dxt(fmt, buf, h, w, padded=padded, padding=padding)
        elif codec == "etc":
            out = ck.ref_etc(strategy, buf, h, w, padded=padded, padding=padding)
        else:
            out = ck.ref_pvrtc(buf, h, w)
        assert out is not None, (codec, fmt, h, w, padded, padding)
        i = len(index)
        arrays["in_%d" % i] = buf
        arrays["out_%d" % i] = out
        index.append(dict(codec=codec, format=fmt, h=h, w=w, padded=list(padded) if padded else None,
                          padding=padding, strategy=strategy, note=note, ncomp=nc))

    sizes = [(4, 4), (8, 8), (5, 7), (16, 12), (1, 1), (13, 3), (32, 32), (23, 41)]
    for fmt in (ck.RGB, ck.BGR, ck.RGBA, ck.BGRA):
        nc = ck.ncomp(fmt)
        for kind in imagegen.KINDS:
            for (h, w) in sizes[:5] if kind != "random" else sizes:
                add("dxt", fmt, imagegen.make(kind, h, w, nc, seed=fmt), h, w, note=kind)
        add("dxt", fmt, imagegen.make("random", 9, 10, nc, 7), 9, 10, padded=(16, 24), note="pad")
        add("dxt", fmt, imagegen.make("alpha_extremes", 6, 6, nc, 8), 6, 6, padded=(13, 6), note="pad rows only")
        add("dxt", fmt, imagegen.make("gradient", 8, 8, nc, 9), 8, 8, padded=(4, 20), note="pad cols only")
        add("dxt", fmt, imagegen.make("random", 12, 10, nc, 10), 12, 10, padding=5, note="row padding")
        add("dxt", fmt, imagegen.make("constant", 3, 3, nc, 11), 3, 3, padded=(12, 12), padding=3, note="pad + row padding")
    for strategy in range(4):
        for kind in imagegen.KINDS:
            for (h, w) in sizes[:4] if kind != "random" else sizes:
                add("etc", ck.RGB, imagegen.make(kind, h, w, 3, seed=strategy), h, w, strategy=strategy, note=kind)
        add("etc", ck.RGB, imagegen.make("random", 9, 10, 3, 7), 9, 10, padded=(16, 24), strategy=strategy, note="pad")
        add("etc", ck.RGB, imagegen.make("smooth_noise", 12, 10, 3, 10), 12, 10, padding=7, strategy=strategy, note="row padding")
    for s in (8, 16, 32, 64):
        for kind in imagegen.KINDS:
            add("pvrtc", ck.RGBA, imagegen.make(kind, s, s, 4, seed=s), s, s, note=kind)
        img = imagegen.make("random", s, s, 4, seed=99)
        img[..., 3] = 255
        add("pvrtc", ck.RGBA, img, s, s, note="opaque random")
        img = imagegen.make("random", s, s, 4, seed=98)
        img[0, 0] = (9, 200, 30, 77)
        img[..., 1] = 0
        img[0, 0, 1] = 200
        add("pvrtc", ck.RGBA, img, s, s, note="P1: zero axis, first pixel leaks")

    arrays["index_json"] = np.frombuffer(json.dumps(index).encode(), np.uint8)
    out_path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(out_path, **arrays)
    print("wrote %s: %d cases, %d bytes" % (out_path, len(index), os.path.getsize(out_path)))

    # Decoder fixture: CRC-32 of the REFERENCE's Decompress() of every unpadded DXT/ETC golden output.
    import zlib
    crcs = {}
    for i, meta in enumerate(index):
        if meta["codec"] == "pvrtc" or meta["padded"]:
            continue
        codec = 2 if meta["codec"] == "etc" else (0 if meta["ncomp"] == 3 else 1)
        px = ck.ref_decompress(codec, meta["format"], np.ascontiguousarray(arrays["out_%d" % i]), meta["h"], meta["w"])
        assert px is not None, meta
        crcs[str(i)] = zlib.crc32(px.tobytes())
    crc_path = os.path.join(ROOT, "tests", "golden", "decode_crc_v1.json")
    json.dump(crcs, open(crc_path, "w"))
    print("wrote %s: %d decode checksums" % (crc_path, len(crcs)))


if __name__ == "__main__":
    main()
