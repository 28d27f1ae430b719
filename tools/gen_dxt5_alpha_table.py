#!/usr/bin/env python3
"""Generates image_compression_b200/csrc/dxt5_alpha_table.inc: the crossing-point table the DXT5 alpha encoder keeps in
shared memory, and checks the whole table-driven classification against the reference's direct search
(ComputeAlphaBits, internal/dxtc_compressor.cc:427-479) for EVERY endpoint pair and EVERY pixel alpha.

Why a table.  The reference scores a pixel alpha against 8 candidate alphas and keeps the first strict minimum.
The candidates lie on a line, so that is a nearest-neighbour search whose answer changes only at crossing points
between neighbouring candidates (ties go to the smaller candidate index; candidates with equal value collapse onto
their smallest index).  Relative to the endpoint a0 the crossing points and the index change at each crossing
depend only on the mode (6- or 8-alpha) and on D = |a0 - a1|, so they are tabulated: 512 entries x 16 bytes.

Entry layout (index = D for the 6-alpha mode a0 <= a1, 256 + D for the 8-alpha mode a0 > a1):
  bytes 0..6   slot value for crossings 1..7
                 8-alpha: 255 - hrel_p   (pixel crosses p iff a <= a0 - hrel_p)
                 6-alpha: bytes 1..5 = hrel_k, k = 1..5 (pixel crosses iff a >= a0 + hrel_k); bytes 0 and 6 are
                          the crossings against the explicit 0 and 255 candidates and are computed on the fly
  byte  7      6-alpha: index the walk is on when it reaches a1 (0 if D == 0 else 1); 8-alpha: 0
  bytes 8..14  index change (mod 8) at crossings 1..7, stored as the HIGH BYTE of the fp16 value (0, 1.0, 2.0 ...)
  byte  15     0
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


def build():
    table = []
    for D in range(256):      # 6-alpha mode
        hrel, step, reps = line_entry(False, D)
        e = [0] * 16
        for k in range(5):
            e[1 + k] = min(hrel[k], 255)
            e[9 + k] = HALF_HI[step[k]]
        e[7] = reps[5]
        table.append(e)
    for D in range(256):      # 8-alpha mode (D == 0 never occurs: a0 > a1)
        e = [0] * 16
        if D > 0:
            hrel, step, reps = line_entry(True, D)
            for p in range(7):
                e[p] = 255 - hrel[p]
                e[8 + p] = HALF_HI[step[p]]
        table.append(e)
    return table


HALF_VAL = {hb: i for i, hb in enumerate(HALF_HI)}


def classify_with_table(table, a0, a1, a):
    """Integer model of the device code: seven crossings, index change accumulated mod 8."""
    if a0 <= a1:
        e = table[a1 - a0]
        last = e[7]
        e0 = 0 if a0 == 0 else 6
        e1 = last if a1 == 255 else 7
        code = e0
        thr = [(a0 + (1 if a0 == 0 else 0) + 1) >> 1] + [a0 + e[1 + k] for k in range(5)] + [(a1 + 257) >> 1]
        step = [(0 - e0) % 8] + [HALF_VAL[e[9 + k]] for k in range(5)] + [(e1 - last) % 8]
        for p in range(7):
            if a >= thr[p]:
                code += step[p]
    else:
        e = table[256 + a0 - a1]
        code = 0
        for p in range(7):
            if a <= a0 - (255 - e[p]):
                code += HALF_VAL[e[8 + p]]
    return code % 8


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
        f.write("// Generated by tools/gen_dxt5_alpha_table.py -- do not edit.  512 entries x 16 bytes.\n")
        for i, e in enumerate(table):
            f.write("/*%3d*/ %s,\n" % (i, ", ".join("0x%02x" % b for b in e)))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
