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
// Returns the 2-bit index to replicate; writes the two 565 endpoints.
__device__ __noinline__ uint32_t dxt_const_colour(uint32_t t, bool always4, uint32_t *c0_out, uint32_t *c1_out) {
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
  *c0_out = c0;
  *c1_out = c1;
  return which;
}

// Weights for IDP.4A: 16*(4,8,1) on the logical (r,g,b); the alpha byte always gets weight 0.
__device__ __forceinline__ uint32_t dxt_lum_weights(bool swap_rb) { return swap_rb ? 0x00408010u : 0x00108040u; }

// Encodes the colour half.  px[i] = pixel i (raster order) as bytes (c0,c1,c2,x) in MEMORY order; the top byte
// is ignored.  fetch(i) must return px[i] (kept as a functor so callers can re-read shared memory instead of
// forcing a register-indexed array into local memory).  Returns {c0 | c1<<16, index bits}.
template <typename Fetch>
__device__ __forceinline__ uint2 dxt1_encode_block(const uint32_t (&px)[16], bool swap_rb, bool always4, Fetch fetch) {
  const uint32_t w16 = dxt_lum_weights(swap_rb);
  uint32_t key[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) key[i] = __dp4a(px[i], w16, static_cast<uint32_t>(i));

  uint32_t kmin = key[0], kmax = key[0] ^ 15u;
#pragma unroll
  for (int i = 1; i < 16; ++i) {
    kmin = min(kmin, key[i]);
    kmax = max(kmax, key[i] ^ 15u);
  }
  // Base colours = first pixel of minimum / maximum luminance; brought into logical (r,g,b) byte order.
  uint32_t p0 = fetch(kmin & 15u), p1 = fetch((kmax & 15u) ^ 15u);
  if (swap_rb) {
    p0 = __byte_perm(p0, 0u, 0x3012);
    p1 = __byte_perm(p1, 0u, 0x3012);
  }
  uint32_t lum0 = kmin & ~15u, lum1 = kmax & ~15u;  // 16 * luminance of p0 / p1
  uint32_t c0 = to_565(p0 & 255u, (p0 >> 8) & 255u, (p0 >> 16) & 255u);
  uint32_t c1 = to_565(p1 & 255u, (p1 >> 8) & 255u, (p1 >> 16) & 255u);
  uint32_t bits;
  if (c0 == c1) {
    // The reference swaps red and blue a second time here (dxtc_compressor.cc:360), i.e. looks up the
    // memory-order colour.
    const uint32_t target = swap_rb ? __byte_perm(p0, 0u, 0x3012) : p0;
    bits = dxt_const_colour(target, always4, &c0, &c1) * 0x55555555u;
  } else {
    if (c0 < c1) {
      uint32_t t = p0; p0 = p1; p1 = t;
      t = c0; c0 = c1; c1 = t;
      t = lum0; lum0 = lum1; lum1 = t;
    }
    // Interpolants come from the UNQUANTISED base colours, channel by channel with truncation.
    const uint32_t r0 = p0 & 255u, g0 = (p0 >> 8) & 255u, b0 = (p0 >> 16) & 255u;
    const uint32_t r1 = p1 & 255u, g1 = (p1 >> 8) & 255u, b1 = (p1 >> 16) & 255u;
    const uint32_t lum2 = 64u * div3_small(2u * r0 + r1) + 128u * div3_small(2u * g0 + g1) + 16u * div3_small(2u * b0 + b1);
    const uint32_t lum3 = 64u * div3_small(r0 + 2u * r1) + 128u * div3_small(g0 + 2u * g1) + 16u * div3_small(b0 + 2u * b1);
    bits = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t l = key[i] & ~15u;
      const uint32_t k = min(min(__usad(lum0, l, 0u), __usad(lum1, l, 1u)), min(__usad(lum2, l, 2u), __usad(lum3, l, 3u)));
      bits = __funnelshift_r(bits, k, 2);  // low two bits of k = chosen index
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
