// etc1_encode.cuh -- per-block ETC1 encoder (all four reference strategies), integer only.
//
// Byte-identical to the reference (paths relative to /root/reference/image_compression/internal/):
//   EncodeEtc1Block           etc_compressor.cc:545-586   (strategy switch; error_lr <= error_tb keeps unflipped)
//   FindBestSubblockEncoding  etc_compressor.cc:460-542   (sub-block means, 555-diff vs 444 rule, wire word)
//   FindBestCodeword          etc_compressor.cc:391-409   (first minimum over codewords 0..7)
//   ComputeCodewordError      etc_compressor.cc:350-385   (4 clamped candidates, SSD, first minimum)
//   FindCodewordHeuristic     etc_compressor.cc:415-455
//
// sm_100a mapping: a candidate colour is one VIADDMNMX.RELU per channel (add, min 255, clamp at 0); the squared
// distance of a pixel to a candidate is VABSDIFF4.U8 followed by IDP.4A of the difference with itself; the
// "first strict minimum" rules are carried by keys (error*4 + index) and (cumulative*8 + codeword).
#pragma once
#include <cstdint>

namespace icb {

enum Etc1Strategy : int { kEtcSplitHorizontally = 0, kEtcSplitVertically = 1, kEtcSmallerError = 2, kEtcHeuristic = 3 };

// Codebook magnitudes {small, large} per codeword, one byte each; index order on the wire is
// +small, +large, -small, -large (etc_compressor.cc:101-110).
__device__ __forceinline__ int etc_small(int cw) { return static_cast<int>(0x2f2118120d090502ull >> (8 * cw)) & 0xff; }
__device__ __forceinline__ int etc_large(int cw) { return static_cast<int>(0xb76a503c2a1d1108ull >> (8 * cw)) & 0xff; }

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

// Error of encoding the 8 pixels selected by `mask` (bit i = raster pixel i) with codeword cw around base
// (r,g,b).  Returns the cumulative error; *indices gets the 2-bit choices in wire positions.
__device__ __forceinline__ uint32_t etc_codeword_error(const uint32_t (&px)[16], uint32_t mask, int cw, int r, int g,
                                                       int b, uint32_t *indices) {
  const int ms = etc_small(cw), ml = etc_large(cw);
  const uint32_t c0 = etc_candidate(r, g, b, ms);
  const uint32_t c1 = etc_candidate(r, g, b, ml);
  const uint32_t c2 = etc_candidate(r, g, b, -ms);
  const uint32_t c3 = etc_candidate(r, g, b, -ml);
  uint32_t total = 0, idx = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (mask & (1u << i)) {  // mask is a compile-time constant at every call site after unrolling
      const uint32_t k = min(min(etc_ssd(px[i], c0) * 4u, etc_ssd(px[i], c1) * 4u + 1u),
                             min(etc_ssd(px[i], c2) * 4u + 2u, etc_ssd(px[i], c3) * 4u + 3u));
      total += k >> 2;
      const int p = 4 * (i & 3) + (i >> 2);  // column-major position: pixel (y,x) -> 4x + y
      idx |= (k & 1u) << p;
      idx |= ((k >> 1) & 1u) << (p + 16);
    }
  }
  *indices = idx;
  return total;
}

template <uint32_t kMask>
__device__ __forceinline__ uint32_t etc_pick_codeword(const uint32_t (&px)[16], int r, int g, int b, bool heuristic,
                                                 uint32_t *indices, uint32_t *error) {
  if (heuristic) {
    const uint32_t base = static_cast<uint32_t>(r | (g << 8) | (b << 16));
    uint32_t dev_rb = 0, dev_g = 0;  // per-channel sums of |base - pixel| over the 8 pixels (each <= 2040)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (kMask & (1u << i)) {
        const uint32_t d = __vabsdiffu4(px[i], base);
        dev_rb += d & 0x00ff00ffu;
        dev_g += (d >> 8) & 0xffu;
      }
    }
    const uint32_t dev = max(max((dev_rb & 0xffffu) >> 3, dev_rb >> 19), dev_g >> 3);
    const uint32_t cw = dev > 144 ? 7 : dev > 93 ? 6 : dev > 70 ? 5 : dev > 51 ? 4 : dev > 35 ? 3 : dev > 23 ? 2 : dev > 12 ? 1 : 0;
    *error = etc_codeword_error(px, kMask, cw, r, g, b, indices);
    return cw;
  }
  uint32_t best = 0xffffffffu, best_idx = 0;
  uint32_t best_cw = 0;
#pragma unroll
  for (int cw = 0; cw < 8; ++cw) {
    uint32_t idx;
    const uint32_t e = etc_codeword_error(px, kMask, cw, r, g, b, &idx);
    if (e < best) {
      best = e;
      best_idx = idx;
      best_cw = cw;
    }
  }
  *indices = best_idx;
  *error = best;
  return best_cw;
}

// One orientation: kFlip=false -> left/right 2x4 halves, kFlip=true -> top/bottom 4x2 halves.
// sum1/sum2: per-channel sums of the two halves, (r | b<<16) and g.  Returns hi word, *lo, *error.
template <bool kFlip>
__device__ __forceinline__ uint32_t etc_encode_split(const uint32_t (&px)[16], uint32_t sum1_rb, uint32_t sum1_g,
                                                     uint32_t sum2_rb, uint32_t sum2_g, bool heuristic, uint32_t *lo,
                                                     uint32_t *error) {
  constexpr uint32_t kMask1 = kFlip ? 0x00ffu : 0x3333u;
  constexpr uint32_t kMask2 = kFlip ? 0xff00u : 0xccccu;
  const uint32_t a1r = (sum1_rb & 0xffffu) >> 3, a1g = sum1_g >> 3, a1b = sum1_rb >> 19;  // truncating means
  const uint32_t a2r = (sum2_rb & 0xffffu) >> 3, a2g = sum2_g >> 3, a2b = sum2_rb >> 19;
  const uint32_t q1r = a1r >> 3, q1g = a1g >> 3, q1b = a1b >> 3, q2r = a2r >> 3, q2g = a2g >> 3, q2b = a2b >> 3;
  const int dr = static_cast<int>(q2r - q1r), dg = static_cast<int>(q2g - q1g), db = static_cast<int>(q2b - q1b);
  const bool diff_mode = dr >= -4 && dr <= 3 && dg >= -4 && dg <= 3 && db >= -4 && db <= 3;
  uint32_t hi = kFlip ? 1u : 0u;
  uint32_t b1r, b1g, b1b, b2r, b2g, b2b;  // base colours as a decoder will see them
  if (diff_mode) {
    hi |= 2u | (q1r << 27) | (q1g << 19) | (q1b << 11) | ((dr & 7u) << 24) | ((dg & 7u) << 16) | ((db & 7u) << 8);
    b1r = (q1r << 3) | (q1r >> 2); b1g = (q1g << 3) | (q1g >> 2); b1b = (q1b << 3) | (q1b >> 2);
    b2r = (q2r << 3) | (q2r >> 2); b2g = (q2g << 3) | (q2g >> 2); b2b = (q2b << 3) | (q2b >> 2);
  } else {
    const uint32_t n1r = a1r >> 4, n1g = a1g >> 4, n1b = a1b >> 4, n2r = a2r >> 4, n2g = a2g >> 4, n2b = a2b >> 4;
    hi |= (n1r << 28) | (n1g << 20) | (n1b << 12) | (n2r << 24) | (n2g << 16) | (n2b << 8);
    b1r = n1r * 17; b1g = n1g * 17; b1b = n1b * 17;
    b2r = n2r * 17; b2g = n2g * 17; b2b = n2b * 17;
  }
  uint32_t idx1, idx2, e1, e2;
  const uint32_t cw1 = etc_pick_codeword<kMask1>(px, b1r, b1g, b1b, heuristic, &idx1, &e1);
  const uint32_t cw2 = etc_pick_codeword<kMask2>(px, b2r, b2g, b2b, heuristic, &idx2, &e2);
  hi |= (cw1 << 5) | (cw2 << 2);
  *lo = idx1 | idx2;
  *error = e1 + e2;
  return hi;
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
    q_g[q] += (px[i] >> 8) & 0xffu;
  }
  const uint32_t left_rb = q_rb[0] + q_rb[2], left_g = q_g[0] + q_g[2];
  const uint32_t right_rb = q_rb[1] + q_rb[3], right_g = q_g[1] + q_g[3];
  const uint32_t top_rb = q_rb[0] + q_rb[1], top_g = q_g[0] + q_g[1];
  const uint32_t bottom_rb = q_rb[2] + q_rb[3], bottom_g = q_g[2] + q_g[3];
  uint32_t hi, lo, err;
  if (strategy == kEtcSplitHorizontally) {
    hi = etc_encode_split<true>(px, top_rb, top_g, bottom_rb, bottom_g, false, &lo, &err);
  } else if (strategy == kEtcSplitVertically) {
    hi = etc_encode_split<false>(px, left_rb, left_g, right_rb, right_g, false, &lo, &err);
  } else if (strategy == kEtcHeuristic) {
    // The reference's bottom-right quadrant sum adds pixel (2,2) twice and never (3,3) (etc_compressor.cc:563-564).
    const uint32_t q3_rb = q_rb[3] - (px[15] & 0x00ff00ffu) + (px[10] & 0x00ff00ffu);
    const uint32_t q3_g = q_g[3] - ((px[15] >> 8) & 0xffu) + ((px[10] >> 8) & 0xffu);
    const uint32_t l_rb = q_rb[0] + q_rb[2], r_rb = q_rb[1] + q3_rb, t_rb = q_rb[0] + q_rb[1], b_rb = q_rb[2] + q3_rb;
    const int lr = (l_rb & 0xffffu) >> 3, lg = (q_g[0] + q_g[2]) >> 3, lb = l_rb >> 19;
    const int rr = (r_rb & 0xffffu) >> 3, rg = (q_g[1] + q3_g) >> 3, rb = r_rb >> 19;
    const int tr = (t_rb & 0xffffu) >> 3, tg = (q_g[0] + q_g[1]) >> 3, tb = t_rb >> 19;
    const int br = (b_rb & 0xffffu) >> 3, bg = (q_g[2] + q3_g) >> 3, bb = b_rb >> 19;
    const uint32_t e_lr = (rr - lr) * (rr - lr) + (rg - lg) * (rg - lg) + (rb - lb) * (rb - lb);
    const uint32_t e_tb = (br - tr) * (br - tr) + (bg - tg) * (bg - tg) + (bb - tb) * (bb - tb);
    if (e_lr > e_tb)
      hi = etc_encode_split<false>(px, left_rb, left_g, right_rb, right_g, true, &lo, &err);
    else
      hi = etc_encode_split<true>(px, top_rb, top_g, bottom_rb, bottom_g, true, &lo, &err);
  } else {
    uint32_t lo2, err2;
    hi = etc_encode_split<false>(px, left_rb, left_g, right_rb, right_g, false, &lo, &err);
    const uint32_t hi2 = etc_encode_split<true>(px, top_rb, top_g, bottom_rb, bottom_g, false, &lo2, &err2);
    if (err2 < err) {
      hi = hi2;
      lo = lo2;
    }
  }
  return make_uint2(__byte_perm(hi, 0u, 0x0123), __byte_perm(lo, 0u, 0x0123));
}

}  // namespace icb
