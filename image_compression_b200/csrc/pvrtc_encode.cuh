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

// value * 32 + index as ONE multiply-add on the FMA pipe (IMAD).  The multiplier arrives as a kernel parameter
// (PvrtcParams::key_scale, always 32): with a literal 32 -- in C or in PTX -- ptxas strength-reduces the multiply-add to
// LEA, which issues on the integer pipe that bounds Morph (measured: 130 LEA per block, integer pipe 82 % busy, FMA
// pipe 10 %); a multiplier it cannot see stays an IMAD with a constant-bank operand.
__device__ __forceinline__ uint32_t pv_key(uint32_t value, uint32_t key_scale, uint32_t index) { return value * key_scale + index; }

// Colour as it will decode after being stored as the block's A (is_b=false) or B (is_b=true) colour
// (ApplyColorChannelReduction: opaque colours keep 5,5,4|5 bits of r,g,b; translucent ones 4,4,3|4 and 3 bits of alpha;
// a channel keeps its top n bits and replicates them downwards, ApplyBitDepthReduction).  All four channels at once on the packed word: the kept bits are
// one mask, and "replicate downwards" is the masked word shifted by the channel's bit count (a channel's shifted bits
// are masked so that they do not run into its lower neighbour).  tests/test_host_math.py compares with the per-channel
// form for every value of every channel.
__device__ __forceinline__ uint32_t pv_reduce_colour(uint32_t c, bool is_b) {
  uint32_t opaque, translucent;
  if (is_b) {
    opaque = (c & 0xfff8f8f8u) | ((c >> 5) & 0x00070707u);
    const uint32_t t = c & 0xe0f0f0f0u;
    translucent = t | ((t >> 4) & 0x000f0f0fu) | ((t >> 3) & 0x1c000000u) | ((t >> 6) & 0x03000000u);
  } else {
    opaque = (c & 0xfff0f8f8u) | ((c >> 5) & 0x00000707u) | ((c >> 4) & 0x000f0000u);
    const uint32_t t = c & 0xe0e0f0f0u;
    translucent = t | ((t >> 4) & 0x00000f0fu) | ((t >> 3) & 0x1c1c0000u) | ((t >> 6) & 0x03030000u);
  }
  return c >= 0xff000000u ? opaque : translucent;
}

// The block's two extreme colours (GetExtremesFast + ApplyColorChannelReduction).  px[j], j = 8*y + x, are the
// block's 32 pixels; fetch(j) must return px[j] (callers re-read memory by index: a register-indexed lookup would
// be a 31-deep select chain per colour); first_pixel is the image's pixel (0,0), which the reference uses whenever
// an axis is all zero in the block (its "max" slot never moves off index 0); key_scale is 32 (see pv_key).
// Keys: value*32 + j for "first minimum", value*32 + (31 - j) for "first maximum" (ties resolve to the lowest j under
// max) -- each key its own IMAD rather than one key and an XOR, because the FMA pipe is idle here and the integer
// pipe is not.  A key needs 13 bits, so the four channel axes ride two to a register -- (r,b) and (g,a) in 16-bit
// lanes: one mask or byte permute plus two IMADs build four keys, VIMNMX3.U16x2 folds two pixels of two axes per
// instruction -- and so do the lightness keys of two neighbouring pixels.
template <typename Fetch>
__device__ __forceinline__ void pv_block_extremes(const uint32_t (&px)[32], uint32_t first_pixel, uint32_t key_scale,
                                                  Fetch fetch, uint32_t *colour_a, uint32_t *colour_b) {
  uint32_t min_l = 0xffffffffu, max_l = 0u;              // lightness axis: pixels j, j+1 in the two 16-bit lanes
  uint32_t min_rb = 0xffffffffu, max_rb = 0u, min_ga = 0xffffffffu, max_ga = 0u;  // packed channel keys
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    uint32_t nl[2], xl[2];  // n.. = key for the minimum, x.. = key for the maximum
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t nrb[2], xrb[2], nga[2], xga[2], dot[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t p = px[j + 2 * h + t], idx = static_cast<uint32_t>(j + 2 * h + t), rev = 31u - idx;
        dot[t] = __dp4a(p, 0x001c964du, 0u);  // 77r + 150g + 28b <= 65025: lightness is byte 1, bytes 2 and 3 are zero
        const uint32_t rb = p & 0x00ff00ffu, ga = __byte_perm(p, 0u, 0x4341);
        nrb[t] = pv_key(rb, key_scale, idx * 0x10001u);
        xrb[t] = pv_key(rb, key_scale, rev * 0x10001u);
        nga[t] = pv_key(ga, key_scale, idx * 0x10001u);
        xga[t] = pv_key(ga, key_scale, rev * 0x10001u);
      }
      min_rb = __vimin3_u16x2(min_rb, nrb[0], nrb[1]);
      max_rb = __vimax3_u16x2(max_rb, xrb[0], xrb[1]);
      min_ga = __vimin3_u16x2(min_ga, nga[0], nga[1]);
      max_ga = __vimax3_u16x2(max_ga, xga[0], xga[1]);
      // one byte permute divides both lightness sums by 256 and puts them side by side; one IMAD keys both pixels
      const uint32_t light2 = __byte_perm(dot[0], dot[1], 0x7531);
      const uint32_t i0 = static_cast<uint32_t>(j + 2 * h), i1 = i0 + 1u;
      nl[h] = pv_key(light2, key_scale, i0 | (i1 << 16));
      xl[h] = pv_key(light2, key_scale, (31u - i0) | ((31u - i1) << 16));
    }
    min_l = __vimin3_u16x2(min_l, nl[0], nl[1]);
    max_l = __vimax3_u16x2(max_l, xl[0], xl[1]);
  }
  min_l = min(min_l & 0xffffu, min_l >> 16);
  max_l = max(max_l & 0xffffu, max_l >> 16);
  // axis order of the reference: lightness, r, g, b, a
  const uint32_t kmin[5] = {min_l, min_rb & 0xffffu, min_ga & 0xffffu, min_rb >> 16, min_ga >> 16};
  const uint32_t kmax[5] = {max_l, max_rb & 0xffffu, max_ga & 0xffffu, max_rb >> 16, max_ga >> 16};
  uint32_t best = 0, c0 = 0, c1 = 0;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const uint32_t cmin = fetch(kmin[k] & 31u);
    const uint32_t cmax = (kmax[k] >> 5) == 0u ? first_pixel : fetch(31u - (kmax[k] & 31u));  // all-zero axis
    const uint32_t d = pv_l1(cmin, cmax);
    if (k == 0 || d > best) {  // strict '>' from 0: pair 0 unless a later one is strictly farther apart
      best = k == 0 ? d : max(best, d);
      c0 = cmin;
      c1 = cmax;
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

__device__ __forceinline__ PvLanes pv_split(uint32_t c) { return PvLanes{c & 0x00ff00ffu, __byte_perm(c, 0u, 0x4341)}; }
// Blend of two colours given as lane pairs, with the weights pre-scaled so that the divisor becomes 256:
// (wa*a + wb*b) / 256 per channel, wa + wb == 256.  Each 16-bit lane then holds at most 255*256 and the quotient
// is simply the lane's high byte, so ONE byte permute both divides and re-interleaves (r,g,b,a) -- no shifts or
// masks on the integer pipe, which is what bounds this kernel.
__device__ __forceinline__ uint32_t pv_blend256(PvLanes a, uint32_t wa, PvLanes b, uint32_t wb) {
  return __byte_perm(a.rb * wa + b.rb * wb, a.ga * wa + b.ga * wb, 0x7351);
}

// BestModulation: 0 = A, 1 = (5A+3B)/8, 2 = (3A+5B)/8, 3 = B; stops at the first step that does not improve.
// a, b: the interpolated A and B colours of this pixel as packed bytes.
__device__ __forceinline__ uint32_t pv_pick_modulation(uint32_t pixel, uint32_t a, uint32_t b) {
  const PvLanes la = pv_split(a), lb = pv_split(b);
  const uint32_t d0 = pv_l1(pixel, a);
  const uint32_t d1 = pv_l1(pixel, pv_blend256(la, 160u, lb, 96u));
  const uint32_t d2 = pv_l1(pixel, pv_blend256(la, 96u, lb, 160u));
  const uint32_t d3 = pv_l1(pixel, b);
  const bool s1 = d1 < d0, s2 = s1 && d2 < d1, s3 = s2 && d3 < d2;
  return (s1 ? 1u : 0u) + (s2 ? 1u : 0u) + (s3 ? 1u : 0u);
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

// Eight 2-bit fields of a row word -> eight bytes (two registers of four).
__device__ __forceinline__ uint32_t pv_fields_to_bytes(uint32_t four_fields) {  // input: bits 0..7
  uint32_t x = (four_fields | (four_fields << 12)) & 0x000f000fu;
  return (x | (x << 6)) & 0x03030303u;
}
// Fields 0,2,4,6 of a row word packed into 8 contiguous bits.
__device__ __forceinline__ uint32_t pv_even_fields(uint32_t row) {
  uint32_t x = row & 0x3333u;
  x = (x | (x >> 2)) & 0x0f0fu;
  return (x | (x >> 4)) & 0xffu;
}
// The high bit of each of the eight fields packed into 8 contiguous bits.
__device__ __forceinline__ uint32_t pv_high_bits(uint32_t row) {
  uint32_t x = (row >> 1) & 0x5555u;
  x = (x | (x >> 1)) & 0x3333u;
  x = (x | (x >> 2)) & 0x0f0fu;
  return (x | (x >> 4)) & 0xffu;
}

__device__ __noinline__ uint32_t pv_one_bpp_bits(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  return pv_high_bits(r0) | (pv_high_bits(r1) << 8) | (pv_high_bits(r2) << 16) | (pv_high_bits(r3) << 24);
}

// Modulation mode + data word for one block (CalculateBlockModulationMode / Data).  row[y], y = 0..3: the block's
// rows, eight 2-bit values each (pixel x in bits 2x..2x+1); row[4]: the wrapped row below; right[y]: the 2-bit value
// of the wrapped pixel to the right of row y.  Works on whole rows: value differences are summed four at a time
// with VABSDIFF4.ACC on byte-expanded rows, counts come from population counts.
// Returns the 32 modulation bits; *one_bpp tells the colour packer which mode flag to write.
__device__ __forceinline__ uint32_t pv_pack_modulation(const uint32_t (&row)[5], const uint32_t (&right)[4], bool *one_bpp) {
  uint32_t inter = 0, horizontal = 0, vertical = 0;
  uint32_t lo_bytes[5], hi_bytes[5];
#pragma unroll
  for (int y = 0; y < 5; ++y) {
    lo_bytes[y] = pv_fields_to_bytes(row[y] & 0xffu);
    hi_bytes[y] = pv_fields_to_bytes(row[y] >> 8);
  }
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    inter += __popc((row[y] ^ (row[y] >> 1)) & 0x5555u);  // values 1 and 2 have unequal bits
    // (sic) the reference's "horizontal" count compares with the row BELOW, "vertical" with the pixel to the RIGHT
    horizontal = __vsadu4(lo_bytes[y], lo_bytes[y + 1]) + __vsadu4(hi_bytes[y], hi_bytes[y + 1]) + horizontal;
    // pixel x+1 in the place of pixel x: the byte-expanded row moved down one byte, the right neighbour's value on top
    const uint32_t next_lo = __byte_perm(lo_bytes[y], hi_bytes[y], 0x4321), next_hi = __byte_perm(hi_bytes[y], right[y], 0x4321);
    vertical = __vsadu4(lo_bytes[y], next_lo) + __vsadu4(hi_bytes[y], next_hi) + vertical;
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

  // checkerboard (x ^ y even), two bits each in raster order: even rows keep fields 0,2,4,6, odd rows 1,3,5,7
  uint32_t bits = pv_even_fields(row[0]) | (pv_even_fields(row[1] >> 2) << 8) | (pv_even_fields(row[2]) << 16) |
                  (pv_even_fields(row[3] >> 2) << 24);
  // the low bit of the entries at bit positions 0 and 20 carries the sub-mode instead of data
  bits = (mode == kAverage4) ? (bits & ~1u) : (bits | 1u);
  bits = (mode == kVertical) ? (bits | (1u << 20)) : (bits & ~(1u << 20));
  // One bit per pixel (value / 2, raster order) for blocks that are nearly two-level: rare, and as a predicated tail of
  // this function its 28 instructions were issued for every block -- a call is only taken by the blocks that need it.
  if (mode == k1Bpp) bits = pv_one_bpp_bits(row[0], row[1], row[2], row[3]);
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
