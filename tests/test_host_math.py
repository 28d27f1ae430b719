"""Exhaustive checks of the integer identities the CUDA encoders rely on (pure Python; no GPU)."""


def blinn(v, bits):
    """Reference quantiser, internal/color_util.h:156-164."""
    m = (1 << bits) - 1
    i = v * m + 128
    return (i + (i >> 8)) >> 8


def test_quantiser_multiply_shift_forms():
    for v in range(256):
        assert (v * 249 + 1024) >> 11 == blinn(v, 5)
        assert (v * 253 + 512) >> 10 == blinn(v, 6)


def test_division_by_three_forms():
    for x in range(766):  # 2*255 + 255
        assert (x * 683) >> 11 == x // 3
        assert (x * (683 << 21)) >> 32 == x // 3


def test_third_of_the_high_lane():
    """dxt_encode.cuh: floor(hi / 3) of a word lo + 65536 * hi, lanes <= 765, as one multiply-high."""
    for hi in range(766):
        for lo in (0, 1, 2, 255, 509, 510, 763, 764, 765):
            assert (((lo + 65536 * hi) * (683 << 5)) >> 32) == hi // 3, (lo, hi)
    for lo in range(766):
        for hi in (0, 1, 2, 3, 254, 255, 511, 764, 765):
            assert (((lo + 65536 * hi) * (683 << 5)) >> 32) == hi // 3, (lo, hi)


def test_quantiser_on_lanes_does_not_carry():
    assert 255 * 249 + 1024 < 1 << 16 and 255 * 253 + 512 < 1 << 16


def direct_first_min(cands, l):
    best, pick = None, 0
    for c, L in enumerate(cands):
        e = (L - l) ** 2
        if best is None or e < best:
            best, pick = e, c
    return pick


def line_search(cands, l, i):
    """The crossing-point formulation used in dxt_encode.cuh (general path), in integers: keys 8*L + c for the
    candidates, 8*l + (lane index < 8) for the pixel."""
    keys = sorted(8 * L + c for c, L in enumerate(cands))
    rep, acc, v = keys[0], keys[0] & 3, 8 * l + (i & 7)
    for b in keys[1:]:
        same = (b - rep) < 4
        cb, cr = b & 3, rep & 3
        h = ((rep + b + (8 if cb < cr else 16)) >> 1) & ~7
        if not same:
            if v >= h:
                acc += (cb - cr) & 3
            rep = b
    return acc & 3


def monotone_search(cands, l, i):
    """The usual-case form (dxt_encode.cuh, all_regular): candidates (L0, L2, L3, L1) strictly ascending along the line,
    three crossings with index changes +2, +1, -2 on signed 16-bit lanes."""
    a0, a1, a2, a3 = (8 * L for L in (cands[0], cands[2], cands[3], cands[1]))
    h1 = ((a0 + a1 + 16) >> 1) & ~7
    h2 = ((a1 + a2 + 16) >> 1) & ~7
    h3 = ((a2 + a3 + 8) >> 1) & ~7
    key = 8 * l + (i & 7)
    assert key < 1 << 15
    t = [max(min(key + 1 - h, 1), 0) for h in (h1, h2, h3)]
    return 2 * t[0] + t[1] - 2 * t[2]


def test_monotone_search_equals_first_strict_minimum():
    import random
    rng = random.Random(11)
    for _ in range(60000):
        L0 = rng.randint(0, 3300)
        L = sorted({L0, L0 + rng.randint(1, 40), L0 + rng.randint(1, 900), rng.randint(L0, 3315)})
        if len(L) < 4:
            continue
        cands = [L[0], L[3], L[1], L[2]]  # reference order: base0, base1, interpolant next to base0, next to base1
        l = rng.randint(max(0, L[0] - 3), min(3315, L[3] + 3))
        assert monotone_search(cands, l, rng.randint(0, 15)) == direct_first_min(cands, l), (cands, l)


def test_dxt5_plain8_steps_are_constants():
    """dxt5_alpha_indices' warp-uniform form hard-codes what the crossing table says for every 8-alpha entry with
    eight distinct candidates (D = a0 - a1 >= 7): start index 1, index changes +6, -1, -1, -1, -1, -1, -2."""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "image_compression_b200", "csrc", "dxt5_alpha_table.inc")
    rows = {}
    for line in open(path):
        m = re.match(r"/\*\s*(\d+)\*/ (.*),$", line.strip())
        if m:
            rows[int(m.group(1))] = [int(x.rstrip("u"), 16) for x in m.group(2).split(", ")]
    assert len(rows) == 512
    for D in range(7, 256):
        e = rows[256 + D]
        steps = [v - (1 << 32) if v & 0x80000000 else v for v in e[8:15]]
        assert steps == [6, -1, -1, -1, -1, -1, -2] and e[7] == 1, D


def test_line_search_equals_first_strict_minimum():
    import random
    rng = random.Random(7)
    for _ in range(60000):
        mode = rng.random()
        if mode < 0.3:
            L = [rng.randint(0, 3315) for _ in range(4)]
        elif mode < 0.6:
            base = rng.randint(0, 3300)
            L = [base + rng.randint(0, 6) for _ in range(4)]
        else:
            base = rng.randint(0, 3000)
            L = [base + rng.choice([0, 0, 1, 2, 50, 51, 100]) for _ in range(4)]
        l = rng.randint(max(0, min(L) - 3), min(3315, max(L) + 3))
        assert line_search(L, l, rng.randint(0, 15)) == direct_first_min(L, l), (L, l)


def test_pvrtc_colour_reduction_on_the_packed_word():
    """pvrtc_encode.cuh:pv_reduce_colour works on the packed (r,g,b,a) word; this is the per-channel definition
    (ApplyBitDepthReduction, pvrtc_compressor.cc:93-106, 337-349) against it for every value of every channel."""
    def keep(v, n):
        kept = v & ((0xff << (8 - n)) & 0xff)
        out = kept | (kept >> n)
        if n <= 3:
            out |= kept >> (2 * n)
        return out

    def per_channel(c, is_b):
        r, g, b, a = c & 255, (c >> 8) & 255, (c >> 16) & 255, c >> 24
        if a == 255:
            r, g, b = keep(r, 5), keep(g, 5), keep(b, 5 if is_b else 4)
        else:
            r, g, b, a = keep(r, 4), keep(g, 4), keep(b, 4 if is_b else 3), keep(a, 3)
        return r | (g << 8) | (b << 16) | (a << 24)

    def packed(c, is_b):
        if is_b:
            opaque = (c & 0xfff8f8f8) | ((c >> 5) & 0x00070707)
            t = c & 0xe0f0f0f0
            translucent = t | ((t >> 4) & 0x000f0f0f) | ((t >> 3) & 0x1c000000) | ((t >> 6) & 0x03000000)
        else:
            opaque = (c & 0xfff0f8f8) | ((c >> 5) & 0x00000707) | ((c >> 4) & 0x000f0000)
            t = c & 0xe0e0f0f0
            translucent = t | ((t >> 4) & 0x00000f0f) | ((t >> 3) & 0x1c1c0000) | ((t >> 6) & 0x03030000)
        return opaque if c >= 0xff000000 else translucent

    import random
    rng = random.Random(5)
    for is_b in (False, True):
        for v in range(256):
            for shift in (0, 8, 16, 24):
                for others in (0x00000000, 0xffffffff, 0x5aa5c33c):
                    c = (others & ~(0xff << shift)) | (v << shift)
                    assert packed(c, is_b) == per_channel(c, is_b), (hex(c), is_b)
        for _ in range(20000):
            c = rng.getrandbits(32) | (0xff000000 if rng.random() < 0.3 else 0)
            assert packed(c, is_b) == per_channel(c, is_b), (hex(c), is_b)


# ---- addressing identities of the TMA drivers (block4x4_kernels.cuh); the host emulation does not run those kernels ----

def test_tile_window_offset_trick():
    """fetch(i) in an RGBA tile: (i >> 2) rows of 1024 bytes + (i & 3) pixels of 4 bytes == (i * 0x104) & 0xc0c."""
    for i in range(16):
        assert (i * 0x104) & 0xc0c == (i >> 2) * 1024 + (i & 3) * 4


def test_rgb888_row_words_unpack():
    """Four packed RGB pixels from three little-endian words with two funnel shifts (the TMA consumers' form)."""
    import random
    rng = random.Random(3)
    for _ in range(2000):
        b = [rng.randrange(256) for _ in range(12)]
        w0, w1, w2 = (int.from_bytes(bytes(b[4 * k:4 * k + 4]), "little") for k in range(3))
        funnel = lambda lo, hi, s: (((hi << 32) | lo) >> s) & 0xffffffff
        px = [w0 & 0xffffff, funnel(w0, w1, 24) & 0xffffff, funnel(w1, w2, 16) & 0xffffff, w2 >> 8]
        assert px == [b[3 * k] | b[3 * k + 1] << 8 | b[3 * k + 2] << 16 for k in range(4)]


def test_persistent_tile_walk_and_shifted_last_tile():
    """A CTA walks tiles blockIdx, blockIdx + grid, ... as (tx, ty) without dividing; the last tile of a row / column is
    shifted back to end at col1 / row1.  Every block of [row0,row1) x [col0,col1) must be covered, none outside."""
    import random
    rng = random.Random(5)
    BX, BY = 64, 4
    for _ in range(300):
        col0, row0 = 0, rng.randrange(0, 50)
        ncols, nrows = rng.randrange(BX, 700), rng.randrange(BY, 60)
        col1, row1 = col0 + ncols, row0 + nrows
        tiles_x, tiles_y = (ncols + BX - 1) // BX, (nrows + BY - 1) // BY
        num_tiles, grid = tiles_x * tiles_y, rng.randrange(1, 40)
        covered = set()
        step_y, step_x = divmod(grid, tiles_x)
        for cta in range(min(grid, num_tiles)):
            ty, tx = divmod(cta, tiles_x)
            tile = cta
            while tile < num_tiles:
                assert (ty, tx) == divmod(tile, tiles_x)
                bc, br = min(col0 + tx * BX, col1 - BX), min(row0 + ty * BY, row1 - BY)
                assert bc >= col0 and br >= row0
                covered.update((r, c) for r in range(br, br + BY) for c in range(bc, bc + BX))
                tile += grid
                tx += step_x
                ty += step_y
                if tx >= tiles_x:
                    tx -= tiles_x
                    ty += 1
        assert covered == {(r, c) for r in range(row0, row1) for c in range(col0, col1)}
