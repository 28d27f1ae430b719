// dxt_encode.cuh -- per-block DXT1 / DXT5 encoders, one thread per 4x4 block, integer only.
//
// Results are byte-identical to the reference encoder (paths relative to /root/reference/image_compression/):
//   EncodeDxt1Block            internal/dxtc_compressor.cc:482-513
//   ComputeBaseColors          internal/dxtc_compressor.cc:284-311   (first min / first max of 4R+8G+B)
//   ComputeColorBits           internal/dxtc_compressor.cc:315-349   (1-D squared luminance distance, first min)
//   ComputeConstantColorBits   internal/dxtc_compressor.cc:353-369 + GetBestDxtcConstColors
//                              internal/dxtc_const_color_table.cc:322-392
//   ComputeBaseAlphas/AlphaBits internal/dxtc_compressor.cc:374-479
//
// How the scalar loops map to sm_100a instructions (checked with cuobjdump -sass):
//   * luminance and its raster index are produced together by one IDP.4A: key = 16*(4R+8G+B) + i.  The first
//     minimum in raster order is min(key); the first maximum is max(key ^ 15).  VIMNMX3 reduces three keys per
//     instruction.
//   * per-pixel index search: |16*L_c - 16*l| + c is one VABSDIFF.U32 with accumulate; the smallest such value
//     over c carries the reference's "first strict minimum" tie-break in its low bits.  A funnel shift
//     (SHF.R.W) moves those bits into the output word without masking.
#pragma once
#include <cstdint>

namespace icb {

// Table of optimal endpoint pairs for a constant channel value; regenerated, not copied
// (tools/gen_dxt_const_table.py re-runs the published search and checks it against the reference).
__device__ const uint8_t g_dxt_const_endpoints[256][8] = {
#include "dxt_const_table.inc"
};

// Blinn rounded quantiser: round(v * max / 255)  (internal/color_util.h:156-164).
__device__ __forceinline__ uint32_t quant_round(uint32_t v, uint32_t maxv) {
  const uint32_t i = v * maxv + 128u;
  return (i + (i >> 8)) >> 8;
}

__device__ __forceinline__ uint32_t to_565(uint32_t r, uint32_t g, uint32_t b) {
  return (quant_round(r, 31u) << 11) | (quant_round(g, 63u) << 5) | quant_round(b, 31u);
}

__device__ __forceinline__ uint32_t expand5(uint32_t v) { return (v << 3) | (v >> 2); }
__device__ __forceinline__ uint32_t expand6(uint32_t v) { return (v << 2) | (v >> 4); }

// floor(x / 3) for 0 <= x <= 765 (one IMAD + one shift).
__device__ __forceinline__ uint32_t div3_small(uint32_t x) { return (x * 683u) >> 11; }

// (4|dr| + 8|dg| + |db|)^2 between target t and colour (r,g,b)  (internal/color_util.h:410-417).
__device__ __forceinline__ uint32_t lum_of_diff_sq(uint32_t tr, uint32_t tg, uint32_t tb, uint32_t r, uint32_t g,
                                                   uint32_t b) {
  const uint32_t d = 4u * __usad(tr, r, 0u) + 8u * __usad(tg, g, 0u) + __usad(tb, b, 0u);
  return d * d;
}

// Constant-colour block (rare path, divergent by design).  `t` is the target colour as (r,g,b) bytes.
// Returns c0 | c1 << 16 | index << 32: the two 565 endpoints and the 2-bit index to replicate.  (Packed return
// value: pointer out-parameters of a non-inlined function would force the caller's registers through local
// memory on every block, not just on constant ones.)
__device__ __noinline__ uint64_t dxt_const_colour(uint32_t t, bool always4) {
  const uint32_t tr = t & 255u, tg = (t >> 8) & 255u, tb = (t >> 16) & 255u;
  const uint32_t qr = quant_round(tr, 31u), qg = quant_round(tg, 63u), qb = quant_round(tb, 31u);
  uint32_t c0 = (qr << 11) | (qg << 5) | qb, c1 = c0, which = 0;
  uint32_t best = lum_of_diff_sq(tr, tg, tb, expand5(qr), expand6(qg), expand5(qb));
  const uint8_t *row_r = g_dxt_const_endpoints[tr];
  const uint8_t *row_g = g_dxt_const_endpoints[tg];
  const uint8_t *row_b = g_dxt_const_endpoints[tb];
  if (!always4) {  // 1/2 blend of a three-colour block (DXT1 only)
    const uint32_t e0r = row_r[2], e1r = row_r[3], e0g = row_g[6], e1g = row_g[7], e0b = row_b[2], e1b = row_b[3];
    const uint32_t err = lum_of_diff_sq(tr, tg, tb, (expand5(e0r) + expand5(e1r)) >> 1,
                                        (expand6(e0g) + expand6(e1g)) >> 1, (expand5(e0b) + expand5(e1b)) >> 1);
    if (err < best) {
      const uint32_t p0 = (e0r << 11) | (e0g << 5) | e0b, p1 = (e1r << 11) | (e1g << 5) | e1b;
      which = 2;
      c0 = p0 < p1 ? p0 : p1;
      c1 = p0 < p1 ? p1 : p0;
      best = err;
    }
  }
  {  // 1/3 blend of a four-colour block
    const uint32_t e0r = row_r[0], e1r = row_r[1], e0g = row_g[4], e1g = row_g[5], e0b = row_b[0], e1b = row_b[1];
    const uint32_t err = lum_of_diff_sq(tr, tg, tb, div3_small(2u * expand5(e0r) + expand5(e1r)),
                                        div3_small(2u * expand6(e0g) + expand6(e1g)),
                                        div3_small(2u * expand5(e0b) + expand5(e1b)));
    if (err < best) {
      const uint32_t p0 = (e0r << 11) | (e0g << 5) | e0b, p1 = (e1r << 11) | (e1g << 5) | e1b;
      if (p0 > p1) {
        which = 2;
        c0 = p0;
        c1 = p1;
      } else {
        which = 3;
        c0 = p1;
        c1 = p0;
      }
    }
  }
  return c0 | (c1 << 16) | (static_cast<uint64_t>(which) << 32);
}

// Weights for IDP.4A: 16*(4,8,1) on the logical (r,g,b); the alpha byte always gets weight 0.
__device__ __forceinline__ uint32_t dxt_lum_weights(bool swap_rb) { return swap_rb ? 0x00408010u : 0x00108040u; }

// 565 quantisation straight from a packed pixel.  round(v*31/255) == (v*249 + 1024) >> 11 and
// round(v*63/255) == (v*253 + 512) >> 10 for every 8-bit v (checked exhaustively in tests/test_host_math.py), so
// one IDP.4A per channel -- weight in the byte that holds the channel, rounding term in the accumulator --
// replaces byte extraction, multiply and the two-step Blinn rounding.
__device__ __forceinline__ uint32_t dxt_to_565(uint32_t p, uint32_t w_red, uint32_t w_blue) {
  const uint32_t xr = __dp4a(p, w_red, 1024u), xg = __dp4a(p, 0x0000fd00u, 512u), xb = __dp4a(p, w_blue, 1024u);
  return (xr & 0xf800u) | ((xg >> 5) & 0x07e0u) | (xb >> 11);
}

__device__ __forceinline__ void sort2(uint32_t &a, uint32_t &b) {
  const uint32_t lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}

// Encodes the colour half.  px[i] = pixel i (raster order) as bytes (c0,c1,c2,x) in MEMORY order; the top byte
// is ignored.  fetch(i) must return px[i] (kept as a functor so callers can re-read shared memory instead of
// forcing a register-indexed array into local memory).  Returns {c0 | c1<<16, index bits}.
//
// Index search.  The reference scores pixel luminance l against the four candidate luminances L_c with
// (L_c - l)^2 and keeps the first strict minimum (dxtc_compressor.cc:334-345).  On a line that is a nearest-
// neighbour search, so the answer only changes where l crosses the midpoint of two neighbouring candidates:
//   * sort the candidates by (L_c, c) once per block (five min/max pairs);
//   * walking upwards, candidate b replaces the current one a iff 2l > L_a + L_b, or 2l == L_a + L_b and b has the
//     smaller index (that is what "first strict minimum" does with a tie); candidates with equal luminance are
//     represented by their smallest index;
//   * per pixel, each of the three crossings is one saturating float add (1.0 if crossed, else 0.0) and one
//     float multiply-add that accumulates the index change -- exact, since every value is an integer below 2^24,
//     and it runs on the FMA pipes while the integer pipe, which bounds this kernel, only does the final bit
//     insert.  The pixel value 2^23 + 16*l + i is produced directly in float format by the IDP.4A that computes
//     the luminance (accumulator 0x4B000000 + i), so no int->float conversion is needed.
template <typename Fetch>
__device__ __forceinline__ uint2 dxt1_encode_block(const uint32_t (&px)[16], bool swap_rb, bool always4, Fetch fetch) {
  const uint32_t w16 = dxt_lum_weights(swap_rb);
  uint32_t kf[16];  // 0x4B000000 + 16*lum + i: as an integer a (lum, index) key, as a float 2^23 + 16*lum + i
  uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    kf[i] = __dp4a(px[i], w16, 0x4b000000u + static_cast<uint32_t>(i));
    kmin = min(kmin, kf[i]);                                                // first minimum in raster order
    kmax = max(kmax, kf[i] ^ 15u);  // index field reversed: first maximum in raster order
  }
  uint32_t p0 = fetch(kmin & 15u), p1 = fetch((kmax & 15u) ^ 15u);  // base colours, memory byte order
  uint32_t lum0 = kmin & 0x000ffff0u, lum1 = kmax & 0x000ffff0u;  // 16 * luminance of p0 / p1
  const uint32_t w_red = swap_rb ? 0x00f90000u : 0x000000f9u, w_blue = swap_rb ? 0x000000f9u : 0x00f90000u;
  uint32_t c0 = dxt_to_565(p0, w_red, w_blue), c1 = dxt_to_565(p1, w_red, w_blue);
  uint32_t bits;
  if (c0 == c1) {
    // The reference swaps red and blue a second time here (dxtc_compressor.cc:360), i.e. it looks up the
    // memory-order colour.
    const uint64_t packed = dxt_const_colour(p0, always4);
    c0 = static_cast<uint32_t>(packed) & 0xffffu;
    c1 = (static_cast<uint32_t>(packed) >> 16) & 0xffffu;
    bits = static_cast<uint32_t>(packed >> 32) * 0x55555555u;
  } else {
    if (c0 < c1) {
      uint32_t t = p0; p0 = p1; p1 = t;
      t = c0; c0 = c1; c1 = t;
      t = lum0; lum0 = lum1; lum1 = t;
    }
    // Interpolants come from the UNQUANTISED base colours, channel by channel with truncation:
    // floor((2a+b)/3) = umulhi(2a+b, 683 << 21) for 2a+b <= 765.
    const uint32_t s_red = swap_rb ? 0x00010000u : 0x00000001u, s_blue = swap_rb ? 0x00000001u : 0x00010000u;
    const uint32_t r0 = __dp4a(p0, s_red, 0u), g0 = __dp4a(p0, 0x00000100u, 0u), b0 = __dp4a(p0, s_blue, 0u);
    const uint32_t r1 = __dp4a(p1, s_red, 0u), g1 = __dp4a(p1, 0x00000100u, 0u), b1 = __dp4a(p1, s_blue, 0u);
    constexpr uint32_t kThird = 683u << 21;
    const uint32_t lum2 = 64u * __umulhi(2u * r0 + r1, kThird) + 128u * __umulhi(2u * g0 + g1, kThird) +
                          16u * __umulhi(2u * b0 + b1, kThird);
    const uint32_t lum3 = 64u * __umulhi(r0 + 2u * r1, kThird) + 128u * __umulhi(g0 + 2u * g1, kThird) +
                          16u * __umulhi(b0 + 2u * b1, kThird);
    // Candidates as keys 16*L_c + c, sorted ascending.
    uint32_t s0 = lum0, s1 = lum1 + 1u, s2 = lum2 + 2u, s3 = lum3 + 3u;
    sort2(s0, s1); sort2(s2, s3); sort2(s0, s2); sort2(s1, s3); sort2(s1, s2);
    const uint32_t sorted[4] = {s0, s1, s2, s3};
    uint32_t rep = s0;                                  // lowest-index candidate of the current luminance
    float acc0 = __uint_as_float(0x4b000000u + (s0 & 3u));  // 2^23 + index of the lowest candidate
    float cross[3], step[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const uint32_t b = sorted[j + 1];
      const bool same_lum = (b - rep) < 4u;             // keys differ only in the index bits
      const uint32_t cb = b & 3u, cr = rep & 3u;
      // pixel key v = 16*l + i crosses iff v >= h, h = 16 * ceil((L_a + L_b + (cb < cr ? 0 : 1)) / 2)
      const uint32_t h = ((rep + b + (cb < cr ? 16u : 32u)) >> 1) & ~15u;
      cross[j] = __uint_as_float(same_lum ? 0x4b7fffffu : 0x4b000000u + h - 1u);
      step[j] = static_cast<float>((cb - cr) & 3u);  // irrelevant when same_lum: that crossing never fires
      rep = same_lum ? rep : b;
    }
    bits = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v = __uint_as_float(kf[i]);
      float acc = fmaf(__saturatef(v - cross[0]), step[0], acc0);
      acc = fmaf(__saturatef(v - cross[1]), step[1], acc);
      acc = fmaf(__saturatef(v - cross[2]), step[2], acc);
      bits = __funnelshift_r(bits, __float_as_uint(acc), 2);  // low two mantissa bits = chosen index (mod 4)
    }
  }
  return make_uint2(c0 | (c1 << 16), bits);
}

// Encodes the DXT5 alpha half from the top byte of each pixel.  Returns the 8 output bytes as two words.
__device__ __forceinline__ uint2 dxt5_encode_alpha(const uint32_t (&px)[16], bool one_pixel) {
  if (one_pixel) {  // window entirely outside the image: both endpoints = that alpha, all indices 0
    const uint32_t a = px[0] >> 24;
    return make_uint2(a | (a << 8), 0u);
  }
  uint32_t n0 = 0, n255 = 0, lo = 255, hi = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t a = px[i] >> 24;
    n0 += (a == 0u);
    n255 += (a == 255u);
    lo = min(lo, a == 0u ? 255u : a);   // 0 and 255 never tighten the interior range
    hi = max(hi, a == 255u ? 0u : a);
  }
  if (lo > hi) {  // every alpha is 0 or 255
    lo = 0;
    hi = 255;
  }
  uint32_t a0, a1;
  if (n0 > 1 || n255 > 1) {
    a0 = lo;
    a1 = hi;
  } else {
    if (n0 > 0) lo = 0;
    if (n255 > 0) hi = 255;
    a0 = hi;
    a1 = lo;
  }
  uint32_t t[8];  // candidate alphas scaled by 8 so the index fits below them
  t[0] = 8u * a0;
  t[1] = 8u * a1;
  if (a0 <= a1) {
    t[2] = 8u * ((4u * a0 + a1) / 5u);
    t[3] = 8u * ((3u * a0 + 2u * a1) / 5u);
    t[4] = 8u * ((2u * a0 + 3u * a1) / 5u);
    t[5] = 8u * ((a0 + 4u * a1) / 5u);
    t[6] = 0u;
    t[7] = 8u * 255u;
  } else {
    t[2] = 8u * ((6u * a0 + a1) / 7u);
    t[3] = 8u * ((5u * a0 + 2u * a1) / 7u);
    t[4] = 8u * ((4u * a0 + 3u * a1) / 7u);
    t[5] = 8u * ((3u * a0 + 4u * a1) / 7u);
    t[6] = 8u * ((2u * a0 + 5u * a1) / 7u);
    t[7] = 8u * ((a0 + 6u * a1) / 7u);
  }
  uint32_t acc_lo = 0, acc_hi = 0;  // 64-bit shift register; 3 bits enter at the top per pixel
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t a8 = (px[i] >> 21) & 0x7f8u;
    uint32_t k = min(__usad(t[0], a8, 0u), __usad(t[1], a8, 1u));
    k = min(k, min(__usad(t[2], a8, 2u), __usad(t[3], a8, 3u)));
    k = min(k, min(__usad(t[4], a8, 4u), __usad(t[5], a8, 5u)));
    k = min(k, min(__usad(t[6], a8, 6u), __usad(t[7], a8, 7u)));
    acc_lo = __funnelshift_r(acc_lo, acc_hi, 3);
    acc_hi = __funnelshift_r(acc_hi, k, 3);
  }
  // 48 code bits now sit in bits 16..63 of (acc_hi:acc_lo); bytes 0,1 are the endpoints.
  return make_uint2((acc_lo & 0xffff0000u) | a0 | (a1 << 8), acc_hi);
}

}  // namespace icb
