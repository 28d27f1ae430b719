#!/usr/bin/env python3
"""Generates image_compression_b200/csrc/dxt5_alpha_table.inc: the crossing-point table of the DXT5 alpha encoder, and
checks the whole table-driven classification against the reference's direct search (ComputeAlphaBits,
internal/dxtc_compressor.cc:427-479) for EVERY endpoint pair and EVERY pixel alpha.

Why a table.  The reference scores a pixel alpha against 8 candidate alphas and keeps the first strict minimum.
The candidates lie on a line, so that is a nearest-neighbour search whose answer changes only at crossing points
between neighbouring candidates (ties go to the smaller candidate index; candidates with equal value collapse onto
their smallest index).  Relative to the endpoint a0 the crossing points and the index change at each crossing
depend only on the mode (6- or 8-alpha) and on D = |a0 - a1|, so they are tabulated: 512 entries x 64 bytes.

The device code walks the candidates in ASCENDING alpha order in both modes: with d = alpha - a0 (16-bit lanes, two
pixels per instruction), crossing s is passed iff d + c_s >= 1, and the running index changes by step_s -- the true
signed difference of the two candidate indices, so every partial sum is itself a valid index 0..7 and several pixels'
3-bit fields can share one accumulator without borrows.

Entry layout, sixteen 32-bit words (index = D for the 6-alpha mode a0 <= a1, 256 + D for the 8-alpha mode a0 > a1):
  words 0..6   c_s in both 16-bit lanes (two's complement): 1 - (threshold_s - a0)
                 6-alpha: slots 1..5 are the line's crossings; slots 0 and 6 are the crossings against the explicit 0
                          and 255 candidates, which depend on a0 / a1 themselves and are patched in by the encoder
                 8-alpha: the seven crossings of the line from the a1 end (index 1) up to a0 (index 0)
  word  7      index the walk starts on (8-alpha; the 6-alpha start depends on a0 and is patched in)
  words 8..14  step_s (signed)
  word  15     6-alpha: index the walk is on when it reaches a1 (0 if D == 0 else 1)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "image_compression_b200", "csrc", "dxt5_alpha_table.inc")
HALF_HI = [0x00, 0x3C, 0x40, 0x42, 0x44, 0x45, 0x46, 0x47]  # high byte of fp16(0..7)


def candidates(a0, a1):
    t = [a0, a1]
    if a0 <= a1:
        t += [((5 - j) * a0 + j * a1) // 5 for j in range(1, 5)] + [0, 255]
    else:
        t += [((7 - j) * a0 + j * a1) // 7 for j in range(1, 7)]
    return t


def direct(a0, a1, a):
    t = candidates(a0, a1)
    best, pick = None, 0
    for c in range(8):
        e = (t[c] - a) ** 2
        if best is None or e < best:
            best, pick = e, c
    return pick


def run_reps(values, indices):
    """values: monotone list; indices: candidate index at each position.  Returns, per position, the smallest
    index among the positions that share its value."""
    reps = []
    for p, v in enumerate(values):
        reps.append(min(i for q, i in zip(values, indices) if q == v))
    return reps


def line_entry(mode8, D):
    """Relative offsets r_p along the line, crossing points and index changes."""
    if mode8:
        r = [-(-p * D // 7) for p in range(8)]            # ceil(p*D/7): offset BELOW a0 of line position p
        idx = [0, 2, 3, 4, 5, 6, 7, 1]
    else:
        r = [k * D // 5 for k in range(6)]                # floor(k*D/5): offset ABOVE a0 of line position k
        idx = [0, 2, 3, 4, 5, 1]
    reps = run_reps(r, idx)
    hrel, step = [], []
    for p in range(1, len(r)):
        tie = 0 if reps[p] < reps[p - 1] else 1
        hrel.append((r[p - 1] + r[p] + tie + 1) >> 1)
        step.append((reps[p] - reps[p - 1]) % 8)
    return hrel, step, reps


def lanes(v):
    return ((v & 0xFFFF) << 16) | (v & 0xFFFF)


def build():
    table = []
    for D in range(256):      # 6-alpha mode: ascending from a0 (index 0) to a1 (index 1)
        hrel, step, reps = line_entry(False, D)
        e = [0] * 16
        for k in range(5):
            e[1 + k] = lanes(1 - hrel[k])
            e[9 + k] = (reps[k + 1] - reps[k]) & 0xFFFFFFFF
        e[15] = reps[5]
        table.append(e)
    for D in range(256):      # 8-alpha mode (D == 0 never occurs: a0 > a1): ascending from a1 up to a0
        e = [0] * 16
        if D > 0:
            hrel, step, reps = line_entry(True, D)   # line positions 0..7 run DOWN from a0; crossing p lies hrel[p-1] below a0
            for s_ in range(7):
                q = 7 - s_                            # ascending slot s_ is the line's crossing q (between positions q-1 and q)
                e[s_] = lanes(hrel[q - 1])            # alpha <= a0 - hrel  <=>  NOT (alpha - a0 + hrel >= 1)
                e[8 + s_] = (reps[q - 1] - reps[q]) & 0xFFFFFFFF
            e[7] = reps[7]
        table.append(e)
    return table


def s16(v):
    v &= 0xFFFF
    return v - 0x10000 if v & 0x8000 else v


def s32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def classify_with_table(table, a0, a1, a):
    """Integer model of the device code (one 16-bit lane): seven crossings in ascending order, signed steps, every
    partial sum a valid index."""
    six = a0 <= a1
    e = list(table[(0 if six else 256) + abs(a0 - a1)])
    start = e[7]
    if six:
        last = e[15]
        e0 = 0 if a0 == 0 else 6
        e1 = last if a1 == 255 else 7
        start = e0
        thr0 = (a0 + (1 if a0 == 0 else 0) + 1) >> 1
        thr6 = (a1 + 257) >> 1
        e[0], e[8] = lanes(1 - (thr0 - a0)), (0 - e0) & 0xFFFFFFFF
        e[6], e[14] = lanes(1 - (thr6 - a0)), (e1 - last) & 0xFFFFFFFF
    code = start
    d = a - a0
    for s_ in range(7):
        t = max(min(d + s16(e[s_]), 1), 0)
        code += t * s32(e[8 + s_])
        assert 0 <= code <= 7, (a0, a1, a, s_, code)
    return code


def main():
    table = build()
    bad = 0
    for a0 in range(256):
        for a1 in range(256):
            lo, hi = min(a0, a1), max(a0, a1)
            # every alpha the encoder can meet: inside [lo, hi] plus the explicit extremes
            for a in set(range(lo, hi + 1)) | {0, 255}:
                if classify_with_table(table, a0, a1, a) != direct(a0, a1, a):
                    bad += 1
                    if bad < 10:
                        print("MISMATCH a0=%d a1=%d a=%d table=%d direct=%d" %
                              (a0, a1, a, classify_with_table(table, a0, a1, a), direct(a0, a1, a)), file=sys.stderr)
    if bad:
        sys.exit("table-driven classification differs from the direct search in %d cases" % bad)
    print("verified: all (a0, a1, alpha) combinations agree with the direct search")
    with open(OUT, "w") as f:
        f.write("// Generated by tools/gen_dxt5_alpha_table.py -- do not edit.  512 entries x 16 words.\n")
        for i, e in enumerate(table):
            f.write("/*%3d*/ %s,\n" % (i, ", ".join("0x%08xu" % w for w in e)))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
