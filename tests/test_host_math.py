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


def direct_first_min(cands, l):
    best, pick = None, 0
    for c, L in enumerate(cands):
        e = (L - l) ** 2
        if best is None or e < best:
            best, pick = e, c
    return pick


def line_search(cands, l, i):
    """The crossing-point formulation used in dxt_encode.cuh, in integers."""
    keys = sorted(16 * L + c for c, L in enumerate(cands))
    rep, acc, v = keys[0], keys[0] & 3, 16 * l + i
    for b in keys[1:]:
        same = (b - rep) < 4
        cb, cr = b & 3, rep & 3
        h = ((rep + b + (16 if cb < cr else 32)) >> 1) & ~15
        if not same:
            if v >= h:
                acc += (cb - cr) & 3
            rep = b
    return acc & 3


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
