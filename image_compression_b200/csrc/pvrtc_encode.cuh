// pvrtc_encode.cuh -- PVRTC1 2bpp RGBA device functions (8x4-pixel blocks), integer only.
//
// Byte-identical to the reference (paths relative to /root/reference/image_compression/internal/):
//   GetExtremesFast             pvrtc_compressor.cc:255-329  (5 candidate axes, first min / first max, image[0] quirk)
//   ApplyColorChannelReduction  pvrtc_compressor.cc:337-349, ApplyBitDepthReduction :93-106
//   GetInterpolatedColor2BPP    pvrtc_compressor.cc:208-237, Interpolate4_2BPP :173-192  (toroidal bilinear upscale)
//   BestModulation              pvrtc_compressor.cc:148-166  (early-exit walk, NOT an argmin)
//   CalculateBlockModulationMode / Data  pvrtc_compressor.cc:395-496
//   EncodeColors                pvrtc_compressor.cc:356-388
//
// sm_100a mapping: the L1 colour distance is one VABSDIFF4.U8.ACC; blends run on two 16-bit lanes per register
// ((r,b) and (g,a)) so a bilinear tap is one IMAD per lane pair; extremes use (value*32 + raster index) keys.
#pragma once
#include <cstdint>

namespace icb {

__device__ __forceinline__ uint32_t pv_l1(uint32_t p, uint32_t q) { return __vsadu4(p, q); }

// Keep the top n bits of an 8-bit value and replicate them downwards (ApplyBitDepthReduction).
__device__ __forceinline__ uint32_t pv_keep_bits(uint32_t v, uint32_t n) {
  const uint32_t kept = v & ((0xffu << (8u - n)) & 0xffu);
  uint32_t out = kept | (kept >> n);
  if (n <= 3u) out |= kept >> (2u * n);
  return out;
}

// Colour as it will decode after being stored as the block's A (is_b=false) or B (is_b=true) colour.
__device__ __forceinline__ uint32_t pv_reduce_colour(uint32_t c, bool is_b) {
  uint32_t r = c & 255u, g = (c >> 8) & 255u, b = (c >> 16) & 255u, a = c >> 24;
  if (a == 255u) {
    r = pv_keep_bits(r, 5);
    g = pv_keep_bits(g, 5);
    b = pv_keep_bits(b, is_b ? 5 : 4);
  } else {
    r = pv_keep_bits(r, 4);
    g = pv_keep_bits(g, 4);
    b = pv_keep_bits(b, is_b ? 4 : 3);
    a = pv_keep_bits(a, 3);
  }
  return r | (g << 8) | (b << 16) | (a << 24);
}

// The block's two extreme colours.  px[j], j = 8*y + x, are the block's 32 pixels; first_pixel is the image's
// pixel (0,0), which the reference uses whenever an axis is all zero in the block (its "max" slot never moves
// off index 0).  Outputs are the already bit-reduced A and B colours.
__device__ __forceinline__ void pv_block_extremes(const uint32_t (&px)[32], uint32_t first_pixel, uint32_t *colour_a,
                                                  uint32_t *colour_b) {
  uint32_t kmin[5], kmax[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    kmin[k] = 0xffffffffu;
    kmax[k] = 0u;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    uint32_t v[5];
    v[0] = __dp4a(px[j], 0x001c964du, 0u) >> 8;  // (77r + 150g + 28b) / 256
    v[1] = px[j] & 255u;
    v[2] = (px[j] >> 8) & 255u;
    v[3] = (px[j] >> 16) & 255u;
    v[4] = px[j] >> 24;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      kmin[k] = min(kmin[k], v[k] * 32u + j);
      kmax[k] = max(kmax[k], v[k] * 32u + (31u - j));
    }
  }
  // Turn keys into colours.  Register-indexed lookup, written as a select chain over the 32 pixels.
  uint32_t cmin[5], cmax[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const uint32_t jmin = kmin[k] & 31u, jmax = 31u - (kmax[k] & 31u);
    uint32_t a = px[0], b = px[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) {
      a = (jmin == j) ? px[j] : a;
      b = (jmax == j) ? px[j] : b;
    }
    cmin[k] = a;
    cmax[k] = (kmax[k] >> 5) == 0u ? first_pixel : b;  // all-zero axis: stays at the image's first pixel
  }
  uint32_t best = 0, c0 = cmin[0], c1 = cmax[0];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const uint32_t d = pv_l1(cmin[k], cmax[k]);
    if (d > best) {
      best = d;
      c0 = cmin[k];
      c1 = cmax[k];
    }
  }
  if (__dp4a(c1, 0x01010101u, 0u) < __dp4a(c0, 0x01010101u, 0u)) {  // darker colour first
    const uint32_t t = c0;
    c0 = c1;
    c1 = t;
  }
  *colour_a = pv_reduce_colour(c0, false);
  *colour_b = pv_reduce_colour(c1, true);
}

struct PvLanes {
  uint32_t rb, ga;  // (r | b<<16), (g | a<<16)
};

__device__ __forceinline__ PvLanes pv_split(uint32_t c) { return PvLanes{c & 0x00ff00ffu, (c >> 8) & 0x00ff00ffu}; }
__device__ __forceinline__ uint32_t pv_join(PvLanes v) { return v.rb | (v.ga << 8); }

// (wa*a + wb*b) >> shift on both lanes; the per-lane sums never carry into the neighbouring lane.
__device__ __forceinline__ PvLanes pv_mix(PvLanes a, uint32_t wa, PvLanes b, uint32_t wb, uint32_t shift) {
  PvLanes out;
  out.rb = ((a.rb * wa + b.rb * wb) >> shift) & 0x00ff00ffu;
  out.ga = ((a.ga * wa + b.ga * wb) >> shift) & 0x00ff00ffu;
  return out;
}

// BestModulation: 0 = A, 1 = (5A+3B)/8, 2 = (3A+5B)/8, 3 = B; stops at the first step that does not improve.
__device__ __forceinline__ uint32_t pv_pick_modulation(uint32_t pixel, PvLanes a, PvLanes b) {
  const uint32_t d0 = pv_l1(pixel, pv_join(a));
  const uint32_t d1 = pv_l1(pixel, pv_join(pv_mix(a, 5u, b, 3u, 3u)));
  const uint32_t d2 = pv_l1(pixel, pv_join(pv_mix(a, 3u, b, 5u, 3u)));
  const uint32_t d3 = pv_l1(pixel, pv_join(b));
  uint32_t m = 0;
  if (d1 < d0) m = (d2 < d1) ? ((d3 < d2) ? 3u : 2u) : 1u;
  return m;
}

// EncodeColors: bit 0 = mode flag, A in bits 1..15, B in bits 16..31.
__device__ __forceinline__ uint32_t pv_pack_colours(uint32_t ca, uint32_t cb, bool one_bpp) {
  const uint32_t ar = ca & 255u, ag = (ca >> 8) & 255u, ab = (ca >> 16) & 255u, aa = ca >> 24;
  const uint32_t br = cb & 255u, bg = (cb >> 8) & 255u, bb = (cb >> 16) & 255u, ba = cb >> 24;
  uint32_t v = one_bpp ? 0u : 1u;
  if (aa == 255u)
    v |= (1u << 15) | ((ab >> 4) << 1) | ((ag >> 3) << 5) | ((ar >> 3) << 10);
  else
    v |= ((ab >> 5) << 1) | ((ag >> 4) << 4) | ((ar >> 4) << 8) | ((aa >> 5) << 12);
  if (ba == 255u)
    v |= (1u << 31) | ((bb >> 3) << 16) | ((bg >> 3) << 21) | ((br >> 3) << 26);
  else
    v |= ((bb >> 4) << 16) | ((bg >> 4) << 20) | ((br >> 4) << 24) | ((ba >> 5) << 28);
  return v;
}

// Modulation mode + data word for one block.  m[y][x]: the block's own 4x8 values in rows 0..3 / columns 0..7,
// the wrapped right-hand neighbour column in x = 8 and the wrapped row below in y = 4.
// Returns the 32 modulation bits; *one_bpp tells the colour packer which mode flag to write.
__device__ __forceinline__ uint32_t pv_pack_modulation(const uint32_t (&m)[5][9], bool *one_bpp) {
  uint32_t inter = 0, horizontal = 0, vertical = 0;
#pragma unroll
  for (int y = 0; y < 4; ++y)
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      inter += (m[y][x] == 1u || m[y][x] == 2u);
      horizontal += __usad(m[y][x], m[y + 1][x], 0u);  // (sic) "horizontal" looks at the row below
      vertical += __usad(m[y][x], m[y][x + 1], 0u);    // (sic) "vertical" looks at the column to the right
    }
  enum { k1Bpp, kAverage4, kVertical, kHorizontal } mode;
  if (inter <= 4u)
    mode = k1Bpp;
  else if (vertical > 10u && vertical > horizontal * 2u)
    mode = kVertical;
  else if (horizontal > 10u && horizontal > vertical * 2u)
    mode = kHorizontal;
  else
    mode = kAverage4;

  uint32_t bits = 0;
  if (mode == k1Bpp) {
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
      for (int x = 0; x < 8; ++x) bits |= (m[y][x] >> 1) << (8 * y + x);
  } else {
    int pos = 0;
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        if ((x ^ y) & 1) continue;  // checkerboard
        uint32_t v = m[y][x];
        if (pos == 0) v = (mode == kAverage4) ? (v & 2u) : (v | 1u);
        if (pos == 20) v = (mode == kVertical) ? (v | 1u) : (v & 2u);
        bits |= v << pos;
        pos += 2;
      }
  }
  *one_bpp = (mode == k1Bpp);
  return bits;
}

// Inverse of FromZOrder (pvrtc_compressor.cc:80-86): block y occupies the EVEN bits of the output index.
__device__ __forceinline__ uint32_t pv_spread_bits(uint32_t v) {
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
__device__ __forceinline__ uint32_t pv_z_index(uint32_t bx, uint32_t by) {
  return (pv_spread_bits(bx) << 1) | pv_spread_bits(by);
}

}  // namespace icb
