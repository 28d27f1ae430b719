// etc1_encode.cuh -- per-block ETC1 encoder (all four reference strategies), integer only.
//
// Byte-identical to the reference (paths relative to /root/reference/image_compression/internal/):
//   EncodeEtc1Block           etc_compressor.cc:545-586   (strategy switch; error_lr <= error_tb keeps unflipped)
//   FindBestSubblockEncoding  etc_compressor.cc:460-542   (sub-block means, 555-diff vs 444 rule, wire word)
//   FindBestCodeword          etc_compressor.cc:391-409   (first minimum over codewords 0..7)
//   ComputeCodewordError      etc_compressor.cc:350-385   (4 clamped candidates, SSD, first minimum)
//   FindCodewordHeuristic     etc_compressor.cc:415-455
//
// sm_100a mapping: a candidate colour is one VIADDMNMX.RELU per channel (add, min 255, clamp at 0); the "first strict
// minimum" rules are carried by keys (error*4 + index) and (cumulative*8 + codeword).  The exhaustive search only
// accumulates errors, in two forms that both work RELATIVE to the distance to the base colour: codewords that cannot
// clamp reduce to a piecewise-linear function of |sum of the pixel's channel differences| (line form: a dozen
// instructions per sub-block and codeword), the others to one IDP.4A per pixel and candidate (dot form); pixel
// indices are computed once, for the winning orientation and codewords, with the direct form VABSDIFF4.U8 + IDP.4A
// (the reference recomputes them inside every ComputeCodewordError call).
#pragma once
#include <cstdint>

namespace icb {

enum Etc1Strategy : int { kEtcSplitHorizontally = 0, kEtcSplitVertically = 1, kEtcSmallerError = 2, kEtcHeuristic = 3 };

// Codebook magnitudes {small, large} per codeword, one byte each; index order on the wire is
// +small, +large, -small, -large (etc_compressor.cc:101-110).
__device__ __forceinline__ int etc_small(int cw) { return static_cast<int>(0x2f2118120d090502ull >> (8 * cw)) & 0xff; }
__device__ __forceinline__ int etc_large(int cw) { return static_cast<int>(0xb76a503c2a1d1108ull >> (8 * cw)) & 0xff; }

// What the rolled codeword loop needs per codeword, the same for every block: the four modifiers (+s, +l, -s, -l) in
// both 16-bit lanes for the candidate builder, and for the line form -3(l+s) in both lanes, l-s in the two low bytes,
// 24 s^2 and -s.  Constant memory, indexed by the (warp-uniform) loop counter: one LDC per value instead of a dozen
// shift / mask / negate / replicate instructions per codeword, which is what made the rolled loop 4 % slower than the
// unrolled one.
struct EtcCodewordConsts {
  uint32_t m2[4];
  uint32_t neg_knee2, step_bytes;  // lanes(-3(l+s)); (l-s) | (l-s) << 8
  int s24, neg_s;                  // 8 pixels * 3 s^2; -s
  int large, pad[3];
};
constexpr int etc_small_c(int cw) { return static_cast<int>(0x2f2118120d090502ull >> (8 * cw)) & 0xff; }
constexpr int etc_large_c(int cw) { return static_cast<int>(0xb76a503c2a1d1108ull >> (8 * cw)) & 0xff; }
constexpr uint32_t etc_lanes(int m) { return (static_cast<uint32_t>(m) & 0xffffu) * 0x10001u; }
constexpr EtcCodewordConsts etc_codeword_consts(int cw) {
  return EtcCodewordConsts{{etc_lanes(etc_small_c(cw)), etc_lanes(etc_large_c(cw)), etc_lanes(-etc_small_c(cw)), etc_lanes(-etc_large_c(cw))},
                           etc_lanes(-3 * (etc_large_c(cw) + etc_small_c(cw))),
                           static_cast<uint32_t>(etc_large_c(cw) - etc_small_c(cw)) * 0x101u,
                           24 * etc_small_c(cw) * etc_small_c(cw), -etc_small_c(cw), etc_large_c(cw), {0, 0, 0}};
}
__constant__ EtcCodewordConsts
    c_etc_codewords[8] = {etc_codeword_consts(0), etc_codeword_consts(1), etc_codeword_consts(2), etc_codeword_consts(3),
                          etc_codeword_consts(4), etc_codeword_consts(5), etc_codeword_consts(6), etc_codeword_consts(7)};

// clamp255(base + modifier) on each of three channels, packed as bytes (r,g,b,0).
__device__ __forceinline__ uint32_t etc_candidate(int r, int g, int b, int modifier) {
  const uint32_t cr = static_cast<uint32_t>(__viaddmin_s32_relu(r, modifier, 255));
  const uint32_t cg = static_cast<uint32_t>(__viaddmin_s32_relu(g, modifier, 255));
  const uint32_t cb = static_cast<uint32_t>(__viaddmin_s32_relu(b, modifier, 255));
  return cr | (cg << 8) | (cb << 16);
}

__device__ __forceinline__ uint32_t etc_ssd(uint32_t px, uint32_t cand) {
  const uint32_t d = __vabsdiffu4(px, cand);
  return __dp4a(d, d, 0u);
}

struct EtcCandidates {
  uint32_t c[4];  // base + {small, large, -small, -large}, clamped, packed (r,g,b,0)
};

__device__ __forceinline__ EtcCandidates etc_candidates(uint32_t base_rgb, int ms, int ml) {
  const int r = base_rgb & 255u, g = (base_rgb >> 8) & 255u, b = (base_rgb >> 16) & 255u;
  EtcCandidates k;
  k.c[0] = etc_candidate(r, g, b, ms);
  k.c[1] = etc_candidate(r, g, b, ml);
  k.c[2] = etc_candidate(r, g, b, -ms);
  k.c[3] = etc_candidate(r, g, b, -ml);
  return k;
}

// Sum over the pixels selected by kMask of the distance to the nearest of the four candidates
// (ComputeCodewordError without the index bookkeeping, which only the winning codeword needs).
template <uint32_t kMask>
__device__ __forceinline__ uint32_t etc_codeword_error(const uint32_t (&px)[16], const EtcCandidates &k) {
  uint32_t total = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (kMask & (1u << i))
      total += min(min(etc_ssd(px[i], k.c[0]), etc_ssd(px[i], k.c[1])), min(etc_ssd(px[i], k.c[2]), etc_ssd(px[i], k.c[3])));
  }
  return total;
}

// Codewords whose four candidates do not clamp lie ON the line base + m*(1,1,1), m = +s, +l, -s, -l, and the distance
// of a pixel to such a candidate is |d|^2 - 2 m S + 3 m^2 with d = pixel - base, S = d_r + d_g + d_b: the nearest of the
// four is decided by |S| alone,
//     min_j |pixel - candidate_j|^2 = |d|^2 + min(3 s^2 - 2 s |S|, 3 l^2 - 2 l |S|)        (exact, no rounding anywhere)
// and the minimum of the two lines in |S| is the first one minus what the second gains beyond their crossing,
//     min(3 s^2 - 2 s u, 3 l^2 - 2 l u) = 3 s^2 - s (2u) - (l - s) max(0, 2u - 3 (l + s)),
// so the error of a sub-block's eight pixels is
//     sum |d|^2 + 24 s^2 - s W - (l - s) sum_i max(0, w_i - 3 (l + s)),      w_i = 2 |S_i|,  W = sum w_i.
// Only the last sum depends on both the codeword and the pixels, and w_i <= 1530 fits a 16-bit lane: with the eight
// w_i of a sub-block in four registers it is four VIADDMNMX.S16x2.RELU, three additions and one IDP.2A that adds the
// two lanes and multiplies by l - s -- a dozen instructions per sub-block and codeword where the general form (the dot
// form below: four IDP.4A, a complement, two minima and an addition per pixel, plus candidates, weights and
// accumulators) needs 84 and the direct form of rounds 1-2 (VABSDIFF4 + IDP.4A per candidate) needed a hundred.
// sum |d|^2, W and the w_i do not depend on the codeword: they are computed once per sub-block.  A codeword qualifies
// when l <= every base channel <= 255 - l, i.e. when l does not exceed the bases' margin (etc_noclamp_margin); the
// margin is made warp-uniform (minimum over the lanes: one REDUX) so that the choice of form never diverges.  The large
// magnitudes grow with the codeword: on uniform random bytes 3.9 of the 8 codewords qualify for a whole warp on average,
// on photo-like mid-tone structure 4.3, on regions that touch black or white in every 128 x 4-pixel strip none.

// The largest modifier magnitude that cannot clamp either of two base colours (bytes r,g,b,0), slightly conservative:
// 127 - max |channel - 128| (a channel of exactly l would still be safe on the low side; losing that case costs nothing
// but an unnecessary general-form evaluation).  Negative when a channel is 0 or 255.
__device__ __forceinline__ int etc_noclamp_margin(uint32_t base1, uint32_t base2) {
  const uint32_t a1 = __vabsdiffu4(base1, 0x00808080u), a2 = __vabsdiffu4(base2, 0x00808080u);  // the top bytes are zero
  const uint32_t m1 = max(max(a1 & 255u, (a1 >> 8) & 255u), a1 >> 16), m2 = max(max(a2 & 255u, (a2 >> 8) & 255u), a2 >> 16);
  return 127 - static_cast<int>(max(m1, m2));
}

// Sum of |pixel - base|^2 over the pixels selected by kMask: the part of a sub-block's error that no codeword changes
// (both the line form and the dot form below compute errors relative to it).  (From the quadrant sums of |p|^2 and p,
// sum |p|^2 - 2 b . sum p + 8 |b|^2, it is 20 instructions per block cheaper and a register more expensive: a spill.)
template <uint32_t kMask>
__device__ __forceinline__ uint32_t etc_d2(const uint32_t (&px)[16], uint32_t base) {
  uint32_t d2 = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (kMask & (1u << i)) {
      const uint32_t d = __vabsdiffu4(px[i], base);
      d2 = __dp4a(d, d, d2);
    }
  }
  return d2;
}

// The codeword-independent part of the line form for the pixels selected by kMask: w_i = 2 |S_i| of the eight pixels,
// two to a register, and their sum.
struct EtcLineTerms {
  uint32_t w[4];  // (w_even | w_odd << 16)
  uint32_t W;     // sum w_i
};
template <uint32_t kMask>
__device__ __forceinline__ EtcLineTerms etc_line_terms(const uint32_t (&px)[16], uint32_t base) {
  EtcLineTerms t;
  const uint32_t minus_2sum = 0u - __dp4a(base, 0x00020202u, 0u);
  int n = 0;  // (a constant in every copy of the unrolled body)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (kMask & (1u << i)) {
      const int s2 = static_cast<int>(__dp4a(px[i], 0x00020202u, minus_2sum));
      const uint32_t w = static_cast<uint32_t>(s2 < 0 ? -s2 : s2);
      if (n & 1)
        t.w[n >> 1] += w << 16;
      else
        t.w[n >> 1] = w;
      ++n;
    }
  }
  t.W = __dp2a_lo(t.w[0] + t.w[1] + t.w[2] + t.w[3], 0x0101u, 0u);  // (a lane holds at most 4 * 1530)
  return t;
}

// Error of the sub-block under a codeword none of whose candidates clamps, d2 = sum |pixel - base|^2.
__device__ __forceinline__ uint32_t etc_line_error(const EtcLineTerms &t, uint32_t d2, const EtcCodewordConsts &c) {
  // max(max(w - knee, -knee), 0) = max(w - knee, 0): the third operand is a register that is there anyway (a literal
  // zero costs a PRMT per instruction: ptxas 12.9 does not put RZ there)
  const uint32_t k = c.neg_knee2;
  const uint32_t beyond = __viaddmax_s16x2_relu(t.w[0], k, k) + __viaddmax_s16x2_relu(t.w[1], k, k) +
                          __viaddmax_s16x2_relu(t.w[2], k, k) + __viaddmax_s16x2_relu(t.w[3], k, k);
  return d2 + static_cast<uint32_t>(c.s24) + t.W * static_cast<uint32_t>(c.neg_s) - __dp2a_lo(beyond, c.step_bytes, 0u);
}

// The general form, for codewords some candidate of which clamps.  The distance of a pixel p to a candidate c, relative
// to its distance to the base b, is LINEAR in the pixel:
//     |p - c|^2 - |p - b|^2 = |c|^2 - |b|^2 - 2 (c - b) . p
// -- one IDP.4A of the pixel with per-candidate weights 2 (b - c) and per-candidate accumulator |c|^2 (the common
// -|b|^2 is added once per sub-block), where the direct form needs VABSDIFF4 + IDP.4A.  IDP.4A takes unsigned bytes:
// the candidates with negative modifiers have c <= b in every channel, weights 2 (b - c) >= 0; those with positive
// modifiers have c >= b and use the COMPLEMENTED pixel q = 255 - p, against which everything mirrors
// (|p - c| = |q - ~c|): weights 2 (c - b) >= 0, accumulator |~c|^2 + (|b|^2 - |~b|^2).  Per pixel and codeword: one
// LOP3 for q, four IDP.4A, two minima, an addition -- the integer pipe, which bounded the direct form (four VABSDIFF4 +
// two VIMNMX3 of its twelve instructions), is left with three.  Weights and accumulators cost 2.5 instructions per
// candidate and base, once per codeword.  Clamped magnitudes above 127 do not fit a doubled byte: they occur only for
// the large modifier of codeword 7 (183), whose two candidates take weights (c - b) and two chained IDP.4A instead.
// a * b + c as ONE multiply-add (left to itself the compiler forms 2 (c - b) as a subtraction and an addition)
__device__ __forceinline__ uint32_t etc_mad(uint32_t a, uint32_t b, uint32_t c) {
#ifdef ICB_HOST_EMULATION
  return a * b + c;
#else
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#endif
}

struct EtcDotBase {
  uint32_t two_b;  // 2 * base as an integer (the addend that turns 2c into 2 (c - b); bytes may carry, the sum does not)
  uint32_t delta;  // |b|^2 - |~b|^2
  uint32_t d2;     // sum |p - b|^2: what the line form adds to
  uint32_t rel;    // d2 - 8 |b|^2: what the sum of the dot form's per-pixel minima is added to
};
template <uint32_t kMask>
__device__ __forceinline__ EtcDotBase etc_dot_base(const uint32_t (&px)[16], uint32_t base) {
  const uint32_t nb = base ^ 0x00ffffffu, bb = __dp4a(base, base, 0u), d2 = etc_d2<kMask>(px, base);
  return EtcDotBase{2u * base, bb - __dp4a(nb, nb, 0u), d2, d2 - 8u * bb};
}

// Sum over the sub-block's pixels of min_j (|p - c_j|^2 - |p - b|^2 + |b|^2), candidates in the reference's order
// (+s, +l, -s, -l).  kWide: the large candidates' magnitudes may exceed 127 (codeword 7).
template <uint32_t kMask, bool kWide>
__device__ __forceinline__ uint32_t etc_dot_error(const uint32_t (&px)[16], const EtcCandidates &k, const EtcDotBase &b) {
  uint32_t w[4], a[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool wide = kWide && (j & 1);
    if (j < 2) {  // c >= b
      const uint32_t nc = k.c[j] ^ 0x00ffffffu;
      w[j] = wide ? k.c[j] - (b.two_b >> 1) : etc_mad(k.c[j], 2u, 0u - b.two_b);
      a[j] = __dp4a(nc, nc, b.delta);
    } else {  // c <= b
      w[j] = wide ? (b.two_b >> 1) - k.c[j] : etc_mad(k.c[j], 0u - 2u, b.two_b);
      a[j] = __dp4a(k.c[j], k.c[j], 0u);
    }
  }
  int total = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (kMask & (1u << i)) {
      const uint32_t p = px[i], q = px[i] ^ 0x00ffffffu;
      const int v0 = static_cast<int>(__dp4a(q, w[0], a[0]));
      const int v1 = static_cast<int>(kWide ? __dp4a(q, w[1], __dp4a(q, w[1], a[1])) : __dp4a(q, w[1], a[1]));
      const int v2 = static_cast<int>(__dp4a(p, w[2], a[2]));
      const int v3 = static_cast<int>(kWide ? __dp4a(p, w[3], __dp4a(p, w[3], a[3])) : __dp4a(p, w[3], a[3]));
      total += min(min(v0, v1), min(v2, v3));
    }
  }
  return b.rel + static_cast<uint32_t>(total);
}

// The same for BOTH sub-blocks of one orientation at once.  Building the clamped candidates is a quarter of the
// integer-pipe work of the exhaustive search when done channel by channel (three VIADDMNMX + two packing operations
// per candidate); the two sub-blocks meet the same modifier at the same time, so their channels ride in 16-bit lanes
// -- (r, b) of each base in one register, (g of base 1, g of base 2) in a third -- and one VIADDMNMX.S16x2.RELU adds,
// caps at 255 and floors at 0 in both lanes: three of those plus two byte permutes make a PAIR of candidates.
__device__ __forceinline__ void etc_candidate_pairs(uint32_t rb1, uint32_t rb2, uint32_t g12, const EtcCodewordConsts &cc,
                                                    EtcCandidates *k1, EtcCandidates *k2) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t m2 = cc.m2[j];
    const uint32_t c_rb1 = __viaddmin_s16x2_relu(rb1, m2, 0x00ff00ffu);
    const uint32_t c_rb2 = __viaddmin_s16x2_relu(rb2, m2, 0x00ff00ffu);
    const uint32_t c_g12 = __viaddmin_s16x2_relu(g12, m2, 0x00ff00ffu);
    k1->c[j] = __byte_perm(c_rb1, c_g12, 0x1240);  // (r, g1, b, 0)
    k2->c[j] = __byte_perm(c_rb2, c_g12, 0x1260);  // (r, g2, b, 0)
  }
}
template <uint32_t kMask1, uint32_t kMask2>
__device__ __forceinline__ void etc_best_codeword_keys(const uint32_t (&px)[16], uint32_t base1, uint32_t base2,
                                                       uint32_t *key1, uint32_t *key2) {
  const uint32_t rb1 = base1 & 0x00ff00ffu, rb2 = base2 & 0x00ff00ffu;
  const uint32_t g12 = __byte_perm(base1, base2, 0x3531);  // (g1, 0, g2, 0): the bases' top bytes are zero
  uint32_t best1 = 0xffffffffu, best2 = 0xffffffffu;
  const EtcDotBase dot1 = etc_dot_base<kMask1>(px, base1), dot2 = etc_dot_base<kMask2>(px, base2);
  // (REDUX.MIN on the margin + 128 >= 0, so that the unsigned minimum is the signed one)
  const int line_margin = static_cast<int>(__reduce_min_sync(__activemask(), static_cast<uint32_t>(etc_noclamp_margin(base1, base2) + 128))) - 128;
  EtcLineTerms line1 = {}, line2 = {};
  if (line_margin >= etc_large_c(0)) {  // (warp-uniform: dark and bright regions do not pay for sums they cannot use)
    line1 = etc_line_terms<kMask1>(px, base1);
    line2 = etc_line_terms<kMask2>(px, base2);
  }
  // The large magnitudes grow with the codeword, so the codewords that take the line form are a PREFIX 0 .. n-1 (n
  // warp-uniform), the dot form takes n .. 6, and codeword 7 (never on the line: 183 > 127) its wide variant with
  // literal constants: three pieces without a choice of form inside (one loop that chose per codeword: 117.8 us
  // against 111.2 on the benchmark input, 162 against 149 on dark content).  The loops are NOT unrolled: unrolled, the
  // exhaustive search is ~5000 instructions (80 KB) -- tolerable while every warp walks it in the same order, but with
  // two forms per codeword the warps of an SM spread over a 105 KB body and stall on instruction fetch (measured with
  // the round's first line form: 17 % fewer instructions executed, 158 -> 180-218 us on structured content), and the
  // dot loop unrolled by two or three spills (125 / 130 us).  Rolled, all forms of both orientations are ~1100
  // instructions; what a rolled loop needs per codeword comes from constant memory (c_etc_codewords).
  int cw = 0;
#pragma unroll 1
  for (; cw < 7 && c_etc_codewords[cw].large <= line_margin; ++cw) {
    const EtcCodewordConsts &cc = c_etc_codewords[cw];
    best1 = min(best1, etc_line_error(line1, dot1.d2, cc) * 8u + static_cast<uint32_t>(cw));
    best2 = min(best2, etc_line_error(line2, dot2.d2, cc) * 8u + static_cast<uint32_t>(cw));
  }
#pragma unroll 1
  for (; cw < 7; ++cw) {
    EtcCandidates k1, k2;
    etc_candidate_pairs(rb1, rb2, g12, c_etc_codewords[cw], &k1, &k2);
    best1 = min(best1, etc_dot_error<kMask1, false>(px, k1, dot1) * 8u + static_cast<uint32_t>(cw));
    best2 = min(best2, etc_dot_error<kMask2, false>(px, k2, dot2) * 8u + static_cast<uint32_t>(cw));
  }
  {
    constexpr EtcCodewordConsts cc7 = etc_codeword_consts(7);
    EtcCandidates k1, k2;
    etc_candidate_pairs(rb1, rb2, g12, cc7, &k1, &k2);
    best1 = min(best1, etc_dot_error<kMask1, true>(px, k1, dot1) * 8u + 7u);
    best2 = min(best2, etc_dot_error<kMask2, true>(px, k2, dot2) * 8u + 7u);
  }
  *key1 = best1;
  *key2 = best2;
}

// FindCodewordHeuristic: codeword from the largest per-channel mean absolute deviation; key as above.
template <uint32_t kMask>
__device__ __forceinline__ uint32_t etc_heuristic_codeword_key(const uint32_t (&px)[16], uint32_t base_rgb) {
  uint32_t dev_rb = 0, dev_g = 0;  // per-channel sums of |base - pixel| over the 8 pixels (each <= 2040)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (kMask & (1u << i)) {
      const uint32_t d = __vabsdiffu4(px[i], base_rgb);
      dev_rb += d & 0x00ff00ffu;
      dev_g += (d >> 8) & 0xffu;
    }
  }
  const uint32_t dev = max(max((dev_rb & 0xffffu) >> 3, dev_rb >> 19), dev_g >> 3);
  const uint32_t cw = dev > 144 ? 7 : dev > 93 ? 6 : dev > 70 ? 5 : dev > 51 ? 4 : dev > 35 ? 3 : dev > 23 ? 2 : dev > 12 ? 1 : 0;
  const EtcCandidates k = etc_candidates(base_rgb, etc_small(cw), etc_large(cw));
  return etc_codeword_error<kMask>(px, k) * 8u + cw;
}

// Base colours of one orientation.  sum1/sum2: per-channel sums of the two halves, (r | b<<16) and g.
// Returns the colour part of the hi word (flip and diff bits included) and the two decoded base colours.
struct EtcBases {
  uint32_t hi, base1, base2;
};
__device__ __forceinline__ EtcBases etc_bases(bool flip, uint32_t sum1_rb, uint32_t sum1_g, uint32_t sum2_rb,
                                              uint32_t sum2_g) {
  const uint32_t a1r = (sum1_rb & 0xffffu) >> 3, a1g = sum1_g >> 3, a1b = sum1_rb >> 19;  // truncating means
  const uint32_t a2r = (sum2_rb & 0xffffu) >> 3, a2g = sum2_g >> 3, a2b = sum2_rb >> 19;
  const uint32_t q1r = a1r >> 3, q1g = a1g >> 3, q1b = a1b >> 3, q2r = a2r >> 3, q2g = a2g >> 3, q2b = a2b >> 3;
  const int dr = static_cast<int>(q2r - q1r), dg = static_cast<int>(q2g - q1g), db = static_cast<int>(q2b - q1b);
  const bool diff_mode = dr >= -4 && dr <= 3 && dg >= -4 && dg <= 3 && db >= -4 && db <= 3;
  EtcBases out;
  out.hi = flip ? 1u : 0u;
  if (diff_mode) {  // 5-bit base + 3-bit two's-complement delta; decoder replicates the top three bits
    out.hi |= 2u | (q1r << 27) | (q1g << 19) | (q1b << 11) | ((dr & 7u) << 24) | ((dg & 7u) << 16) | ((db & 7u) << 8);
    out.base1 = ((q1r << 3) | (q1r >> 2)) | (((q1g << 3) | (q1g >> 2)) << 8) | (((q1b << 3) | (q1b >> 2)) << 16);
    out.base2 = ((q2r << 3) | (q2r >> 2)) | (((q2g << 3) | (q2g >> 2)) << 8) | (((q2b << 3) | (q2b >> 2)) << 16);
  } else {  // two 4-bit colours; decoder multiplies by 17
    const uint32_t n1r = a1r >> 4, n1g = a1g >> 4, n1b = a1b >> 4, n2r = a2r >> 4, n2g = a2g >> 4, n2b = a2b >> 4;
    out.hi |= (n1r << 28) | (n1g << 20) | (n1b << 12) | (n2r << 24) | (n2g << 16) | (n2b << 8);
    out.base1 = (n1r | (n1g << 8) | (n1b << 16)) * 17u;
    out.base2 = (n2r | (n2g << 8) | (n2b << 16)) * 17u;
  }
  return out;
}

// Pixel indices (the lo word) for the chosen orientation and codewords.  Pixels whose sub-block is the same in
// both orientations use a fixed candidate set; the top-right and bottom-left quadrants pick theirs by `flip`.
__device__ __forceinline__ uint32_t etc_pixel_indices(const uint32_t (&px)[16], bool flip, const EtcCandidates &k1,
                                                      const EtcCandidates &k2) {
  EtcCandidates top_right, bottom_left;  // flip: top-right belongs to sub-block 1 (top), bottom-left to 2
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    top_right.c[j] = flip ? k1.c[j] : k2.c[j];
    bottom_left.c[j] = flip ? k2.c[j] : k1.c[j];
  }
  uint32_t key[16];  // distance * 4 + index of the nearest candidate (first minimum)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int quadrant = ((i >> 3) << 1) | ((i >> 1) & 1);
    const EtcCandidates &k = quadrant == 0 ? k1 : quadrant == 3 ? k2 : quadrant == 1 ? top_right : bottom_left;
    key[i] = min(min(etc_ssd(px[i], k.c[0]) * 4u, etc_ssd(px[i], k.c[1]) * 4u + 1u),
                 min(etc_ssd(px[i], k.c[2]) * 4u + 2u, etc_ssd(px[i], k.c[3]) * 4u + 3u));
  }
  // Pixel (y, x) puts the low bit of its index at bit 4x + y and the high bit at 16 + 4x + y.  Column by column: the
  // low bytes of the column's four keys side by side (three PRMT), their bits 0 and bits 1 isolated for all four
  // at once, and one IDP.4A each with the weights 1,2,4,8 (even columns) or 16,32,64,128 (odd columns) drops them at
  // their places within a byte -- 8 instructions per column where shift-and-merge per pixel and bit took 16.
  uint32_t low[2] = {0, 0}, high[2] = {0, 0};  // [0]: columns 0 and 1, [1]: columns 2 and 3
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const uint32_t b = __byte_perm(__byte_perm(key[x], key[4 + x], 0x0040), __byte_perm(key[8 + x], key[12 + x], 0x0040), 0x5410);
    const uint32_t weights = (x & 1) ? 0x80402010u : 0x08040201u;
    low[x >> 1] = __dp4a(b & 0x01010101u, weights, low[x >> 1]);
    high[x >> 1] = __dp4a((b >> 1) & 0x01010101u, weights, high[x >> 1]);
  }
  return __byte_perm(__byte_perm(low[0], low[1], 0x0040), __byte_perm(high[0], high[1], 0x0040), 0x5410);
}

// px[i]: bytes (r,g,b,0) -- the top byte MUST be zero.  Returns the 8 wire bytes as two little-endian words
// (the format stores hi then lo, each big-endian: etc_compressor.cc:172-180).
__device__ __forceinline__ uint2 etc1_encode_block(const uint32_t (&px)[16], int strategy) {
  // 2x2 quadrant sums, (r | b<<16) and g: q[0]=top-left q[1]=top-right q[2]=bottom-left q[3]=bottom-right
  uint32_t q_rb[4] = {0, 0, 0, 0}, q_g[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int q = ((i >> 3) << 1) | ((i >> 1) & 1);
    q_rb[q] += px[i] & 0x00ff00ffu;
    q_g[q] = __dp4a(px[i], 0x00000100u, q_g[q]);  // (one IDP.4A where shift, mask and add are three)
  }
  const EtcBases lr = etc_bases(false, q_rb[0] + q_rb[2], q_g[0] + q_g[2], q_rb[1] + q_rb[3], q_g[1] + q_g[3]);
  const EtcBases tb = etc_bases(true, q_rb[0] + q_rb[1], q_g[0] + q_g[1], q_rb[2] + q_rb[3], q_g[2] + q_g[3]);
  constexpr uint32_t kLeft = 0x3333u, kRight = 0xccccu, kTop = 0x00ffu, kBottom = 0xff00u;

  bool flip;
  uint32_t key1, key2;  // cumulative_error * 8 + codeword of sub-blocks 1 and 2 of the chosen orientation
  if (strategy == kEtcHeuristic) {
    // Orientation from the colour difference of the halves.  The reference's bottom-right quadrant sum adds
    // pixel (2,2) twice and never (3,3) (etc_compressor.cc:563-564); the sub-block means below use the true sums.
    const uint32_t q3_rb = q_rb[3] - (px[15] & 0x00ff00ffu) + (px[10] & 0x00ff00ffu);
    const uint32_t q3_g = q_g[3] - ((px[15] >> 8) & 0xffu) + ((px[10] >> 8) & 0xffu);
    const uint32_t l_rb = q_rb[0] + q_rb[2], r_rb = q_rb[1] + q3_rb, t_rb = q_rb[0] + q_rb[1], b_rb = q_rb[2] + q3_rb;
    const int lr_ = (l_rb & 0xffffu) >> 3, lg = (q_g[0] + q_g[2]) >> 3, lb = l_rb >> 19;
    const int rr = (r_rb & 0xffffu) >> 3, rg = (q_g[1] + q3_g) >> 3, rb = r_rb >> 19;
    const int tr = (t_rb & 0xffffu) >> 3, tg = (q_g[0] + q_g[1]) >> 3, tb_ = t_rb >> 19;
    const int br = (b_rb & 0xffffu) >> 3, bg = (q_g[2] + q3_g) >> 3, bb = b_rb >> 19;
    const uint32_t e_lr = (rr - lr_) * (rr - lr_) + (rg - lg) * (rg - lg) + (rb - lb) * (rb - lb);
    const uint32_t e_tb = (br - tr) * (br - tr) + (bg - tg) * (bg - tg) + (bb - tb_) * (bb - tb_);
    flip = !(e_lr > e_tb);
    if (flip) {
      key1 = etc_heuristic_codeword_key<kTop>(px, tb.base1);
      key2 = etc_heuristic_codeword_key<kBottom>(px, tb.base2);
    } else {
      key1 = etc_heuristic_codeword_key<kLeft>(px, lr.base1);
      key2 = etc_heuristic_codeword_key<kRight>(px, lr.base2);
    }
  } else {
    uint32_t lr1 = 0, lr2 = 0, tb1 = 0, tb2 = 0;
    if (strategy != kEtcSplitHorizontally) etc_best_codeword_keys<kLeft, kRight>(px, lr.base1, lr.base2, &lr1, &lr2);
    if (strategy != kEtcSplitVertically) etc_best_codeword_keys<kTop, kBottom>(px, tb.base1, tb.base2, &tb1, &tb2);
    // kSmallerError keeps the unflipped block unless the flipped one is strictly better (etc_compressor.cc:583)
    flip = strategy == kEtcSplitHorizontally ||
           (strategy == kEtcSmallerError && (tb1 >> 3) + (tb2 >> 3) < (lr1 >> 3) + (lr2 >> 3));
    key1 = flip ? tb1 : lr1;
    key2 = flip ? tb2 : lr2;
  }
  const uint32_t cw1 = key1 & 7u, cw2 = key2 & 7u;
  const uint32_t base1 = flip ? tb.base1 : lr.base1, base2 = flip ? tb.base2 : lr.base2;
  const uint32_t hi = (flip ? tb.hi : lr.hi) | (cw1 << 5) | (cw2 << 2);
  const uint32_t lo = etc_pixel_indices(px, flip, etc_candidates(base1, etc_small(cw1), etc_large(cw1)),
                                        etc_candidates(base2, etc_small(cw2), etc_large(cw2)));
  return make_uint2(__byte_perm(hi, 0u, 0x0123), __byte_perm(lo, 0u, 0x0123));
}

}  // namespace icb
