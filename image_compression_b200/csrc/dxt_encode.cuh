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
// How the scalar loops map to sm_100a instructions (checked with cuobjdump -sass, rates measured with
// tools/microbench/pipe_rates.cu):
//   * luminance and its raster index are produced together by one IDP.4A: key = 16*(4R+8G+B) + i.  The first
//     minimum in raster order is min(key); the first maximum is max(key ^ 15).  VIMNMX3 reduces three keys per
//     instruction.
//   * both index searches (colour: 4 candidates on the luminance line; alpha: 8 candidates) are nearest-neighbour
//     searches on a line, evaluated as "how many crossing points has this pixel passed" with saturating
//     floating-point adds and multiply-adds on exactly representable integers -- fp32 for luminance (values
//     < 2^24), packed fp16 for alpha (values <= 256, two pixels per instruction) -- because the integer pipe
//     (LOP3/SHF/VIMNMX/VABSDIFF, one warp-instruction per two cycles per scheduler) is what bounds these kernels.
#pragma once
#include <cstdint>

namespace icb {

// Table of optimal endpoint pairs for a constant channel value; regenerated, not copied
// (tools/gen_dxt_const_table.py re-runs the published search and checks it against the reference).
__device__ __align__(8) const uint8_t g_dxt_const_endpoints[256][8] = {
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
  // one 8-byte load per channel row (the table rows are 8-byte aligned), bytes picked out of the two words
  const uint2 row_r = *reinterpret_cast<const uint2 *>(g_dxt_const_endpoints[tr]);
  const uint2 row_g = *reinterpret_cast<const uint2 *>(g_dxt_const_endpoints[tg]);
  const uint2 row_b = *reinterpret_cast<const uint2 *>(g_dxt_const_endpoints[tb]);
  auto byte_of = [](uint32_t word, int k) { return (word >> (8 * k)) & 255u; };
  if (!always4) {  // 1/2 blend of a three-colour block (DXT1 only): columns 2,3 (r, b) and 6,7 (g)
    const uint32_t e0r = byte_of(row_r.x, 2), e1r = byte_of(row_r.x, 3), e0g = byte_of(row_g.y, 2), e1g = byte_of(row_g.y, 3),
                   e0b = byte_of(row_b.x, 2), e1b = byte_of(row_b.x, 3);
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
  {  // 1/3 blend of a four-colour block: columns 0,1 (r, b) and 4,5 (g)
    const uint32_t e0r = byte_of(row_r.x, 0), e1r = byte_of(row_r.x, 1), e0g = byte_of(row_g.y, 0), e1g = byte_of(row_g.y, 1),
                   e0b = byte_of(row_b.x, 0), e1b = byte_of(row_b.x, 1);
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
//     insert.  The pixel value 2^23 + 16*l is produced directly in float format by the IDP.4A that computes the
//     luminance (accumulator 0x4B000000), so no int->float conversion is needed.
// kf[i] = 0x4B000000 + 16*lum(pixel i): read as a float it is 2^23 + 16*lum, which the index search consumes as is.
constexpr uint32_t kDxtLumBias = 0x4b000000u;

// A caller that stages pixels in a buffer it wants back early passes `release`; it is called once, right after the
// two base colours have been re-read through `fetch` (nothing reads the staged pixels after that).
struct NoRelease {
  __device__ __forceinline__ void operator()() const {}
};

template <bool kFullWarp, typename Fetch, typename Release = NoRelease>
__device__ __forceinline__ uint2 dxt1_encode_from_keys(const uint32_t (&kf)[16], bool swap_rb, bool always4, Fetch fetch,
                                                       Release release = Release()) {
  // First minimum / first maximum in raster order: 16-bit keys 16*lum + i (lum <= 3315), two pixels per register;
  // the maximum uses the index field reversed (^15) so that ties resolve to the lowest index.  VIMNMX3.U16x2
  // folds two more registers (four pixels) per instruction.
  uint32_t pk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) pk[k] = __byte_perm(kf[2 * k], kf[2 * k + 1], 0x5410) + ((2u * k) | ((2u * k + 1u) << 16));
  uint32_t mn = __vimin3_u16x2(pk[0], pk[1], pk[2]);
  mn = __vimin3_u16x2(mn, pk[3], pk[4]);
  mn = __vimin3_u16x2(mn, pk[5], pk[6]);
  mn = __vminu2(mn, pk[7]);
  uint32_t mx = __vimax3_u16x2(pk[0] ^ 0x000f000fu, pk[1] ^ 0x000f000fu, pk[2] ^ 0x000f000fu);
  mx = __vimax3_u16x2(mx, pk[3] ^ 0x000f000fu, pk[4] ^ 0x000f000fu);
  mx = __vimax3_u16x2(mx, pk[5] ^ 0x000f000fu, pk[6] ^ 0x000f000fu);
  mx = __vmaxu2(mx, pk[7] ^ 0x000f000fu);
  const uint32_t kmin = min(mn & 0xffffu, mn >> 16), kmax = max(mx & 0xffffu, mx >> 16);
  uint32_t p0 = fetch(kmin & 15u), p1 = fetch((kmax & 15u) ^ 15u);  // base colours, memory byte order
  release();
  uint32_t lum0 = kmin & 0xfff0u, lum1 = kmax & 0xfff0u;            // 16 * luminance of p0 / p1
  const uint32_t w_red = swap_rb ? 0x00f90000u : 0x000000f9u, w_blue = swap_rb ? 0x000000f9u : 0x00f90000u;
  uint32_t c0 = dxt_to_565(p0, w_red, w_blue), c1 = dxt_to_565(p1, w_red, w_blue);
  // Everything up to the warp vote below is computed for constant blocks too (and ignored): the vote has to sit
  // where the warp has not yet diverged on "is this block constant".
  const bool constant = c0 == c1;
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
  // Usual case, decided once per warp so the branch never diverges: the interpolants lie strictly between the
  // base colours, i.e. the candidates are already ordered 0,2,3,1 (or 1,3,2,0) along the luminance line with no
  // two equal.  Then the crossing order, the tie rules and the index changes are fixed and only the three
  // midpoints have to be computed.  (Constant blocks vote yes: they take neither path, and a no would send the
  // warp's other blocks down the slower general path -- flat image regions would pay for it.)
  const bool rising = lum0 < lum2 && lum2 < lum3 && lum3 < lum1;
  const bool falling = lum0 > lum2 && lum2 > lum3 && lum3 > lum1;
  const bool all_regular = __all_sync(kFullWarp ? 0xffffffffu : __activemask(), constant || rising || falling);
  uint32_t bits;
  if (constant) {
    // The reference swaps red and blue a second time here (dxtc_compressor.cc:360), i.e. it looks up the
    // memory-order colour.
    const uint64_t packed = dxt_const_colour(p0, always4);
    c0 = static_cast<uint32_t>(packed) & 0xffffu;
    c1 = (static_cast<uint32_t>(packed) >> 16) & 0xffffu;
    bits = static_cast<uint32_t>(packed >> 32) * 0x55555555u;
  } else if (all_regular) {
    // Ascending index sequence along the luminance line: rising 0,2,3,1, falling 1,3,2,0; a tie goes to the smaller
    // index.  Crossing points h1 < h2 < h3 (multiples of 16, "crossed iff 16*l >= h").  In both sequences the high
    // index bit is set exactly between the outer crossings and the low bit flips at the middle one:
    //   bit1 = [h1 <= v < h3] = sat(R + 1 - |v - mid|)      mid, R = centre and half-width of [h1, h3 - 16]
    //   bit0 = [v >= h2] (rising) / [v < h2] (falling) = sat(+-v -+ h2 ...)
    // and 2*bit1 + bit0 is added into the mantissa of 2^23 at the pixel's position, eight pixels per accumulator:
    // five exact FADD/FFMA per pixel on the FMA pipes and no per-pixel work on the integer pipe.
    const uint32_t a0 = rising ? lum0 : lum1, a1 = rising ? lum2 : lum3, a2 = rising ? lum3 : lum2, a3 = rising ? lum1 : lum0;
    const uint32_t h1 = ((a0 + a1 + 32u) >> 1) & ~15u;                     // 0->2 / 1->3: larger index, tie stays
    const uint32_t h2 = ((a1 + a2 + (rising ? 32u : 16u)) >> 1) & ~15u;    // 2->3 stays on tie, 3->2 moves
    const uint32_t h3 = ((a2 + a3 + 16u) >> 1) & ~15u;                     // 3->1 / 2->0: smaller index, tie moves
    const float mid = __uint_as_float(kDxtLumBias + ((h1 + h3 - 16u) >> 1));
    const float rp1 = __uint_as_float(kDxtLumBias + ((h3 - h1 - 16u) >> 1) + 1u) - 8388608.0f;  // R + 1
    const float sgn = rising ? 1.0f : -1.0f;
    const float k2 = __uint_as_float(rising ? 0xcb000000u + h2 - 1u : kDxtLumBias + h2);
    float acc_lo = 8388608.0f, acc_hi = 8388608.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v = __uint_as_float(kf[i]);
      const float u = __saturatef(rp1 - fabsf(v - mid));
      const float t = __saturatef(fmaf(v, sgn, k2));
      const float z = fmaf(u, 2.0f, t);
      const float scale = static_cast<float>(1u << (2 * (i & 7)));
      if (i < 8)
        acc_lo = fmaf(z, scale, acc_lo);
      else
        acc_hi = fmaf(z, scale, acc_hi);
    }
    bits = __byte_perm(__float_as_uint(acc_lo), __float_as_uint(acc_hi), 0x5410);
  } else {
    float acc0, cross[3], step[3];
    {
      // General case (flat blocks, crossed or equal candidates): sort the candidates as keys 16*L_c + c.
      uint32_t s0 = lum0, s1 = lum1 + 1u, s2 = lum2 + 2u, s3 = lum3 + 3u;
      sort2(s0, s1); sort2(s2, s3); sort2(s0, s2); sort2(s1, s3); sort2(s1, s2);
      const uint32_t sorted[4] = {s0, s1, s2, s3};
      uint32_t rep = s0;                                  // lowest-index candidate of the current luminance
      acc0 = __uint_as_float(kDxtLumBias + (s0 & 3u));    // 2^23 + index of the lowest candidate
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint32_t b = sorted[j + 1];
        const bool same_lum = (b - rep) < 4u;             // keys differ only in the index bits
        const uint32_t cb = b & 3u, cr = rep & 3u;
        // pixel key v = 16*l crosses iff v >= h, h = 16 * ceil((L_a + L_b + (cb < cr ? 0 : 1)) / 2)
        const uint32_t h = ((rep + b + (cb < cr ? 16u : 32u)) >> 1) & ~15u;
        cross[j] = __uint_as_float(same_lum ? 0x4b7fffffu : kDxtLumBias + h - 1u);
        step[j] = static_cast<float>((cb - cr) & 3u);  // irrelevant when same_lum: that crossing never fires
        rep = same_lum ? rep : b;
      }
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

// Keys from 16 packed pixels (bytes c0,c1,c2,x in memory order; x ignored).
template <bool kFullWarp = false, typename Fetch, typename Release = NoRelease>
__device__ __forceinline__ uint2 dxt1_encode_block(const uint32_t (&px)[16], bool swap_rb, bool always4, Fetch fetch,
                                                   Release release = Release()) {
  const uint32_t w16 = dxt_lum_weights(swap_rb);
  uint32_t kf[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i] = __dp4a(px[i], w16, kDxtLumBias);
  return dxt1_encode_from_keys<kFullWarp>(kf, swap_rb, always4, fetch, release);
}

// Keys straight from four rows of packed RGB888 (three 32-bit words = four pixels per row): the byte weights of
// IDP.4A do the unpacking, a pixel that straddles two words is two chained IDPs.  rows[y][0..2] = the 12 bytes.
template <bool kFullWarp = false, typename Fetch, typename Release = NoRelease>
__device__ __forceinline__ uint2 dxt1_encode_rgb888_rows(const uint32_t (&rows)[4][3], bool swap_rb, bool always4, Fetch fetch,
                                                         Release release = Release()) {
  const uint32_t w = dxt_lum_weights(swap_rb);  // bytes (w0, w1, w2, 0) for memory-order channels 0,1,2
  const uint32_t w0 = w & 0xffu, w1 = (w >> 8) & 0xffu, w2 = (w >> 16) & 0xffu;
  const uint32_t wa = w;                                   // pixel 0: word 0 bytes 0,1,2
  const uint32_t wb_lo = w0 << 24, wb_hi = w1 | (w2 << 8);  // pixel 1: word 0 byte 3, word 1 bytes 0,1
  const uint32_t wc_lo = (w0 << 16) | (w1 << 24), wc_hi = w2;  // pixel 2: word 1 bytes 2,3, word 2 byte 0
  const uint32_t wd = w << 8;                              // pixel 3: word 2 bytes 1,2,3
  uint32_t kf[16];
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    kf[4 * y + 0] = __dp4a(rows[y][0], wa, kDxtLumBias);
    kf[4 * y + 1] = __dp4a(rows[y][1], wb_hi, __dp4a(rows[y][0], wb_lo, kDxtLumBias));
    kf[4 * y + 2] = __dp4a(rows[y][2], wc_hi, __dp4a(rows[y][1], wc_lo, kDxtLumBias));
    kf[4 * y + 3] = __dp4a(rows[y][2], wd, kDxtLumBias);
  }
  return dxt1_encode_from_keys<kFullWarp>(kf, swap_rb, always4, fetch, release);
}

// ---------------------------------------------------------------------------------------------------------
// DXT5 alpha block
// ---------------------------------------------------------------------------------------------------------

// Crossing-point table, 512 entries x 16 bytes; layout and derivation in tools/gen_dxt5_alpha_table.py, which also
// verifies the table-driven search against the reference's direct search for every (a0, a1, alpha).
__device__ __align__(16) const uint8_t g_dxt5_alpha_table[512 * 16] = {
#include "dxt5_alpha_table.inc"
};
constexpr int kDxt5AlphaTableBytes = 512 * 16;

// Packed fp16 helpers on raw 32-bit registers (two lanes per instruction; HFMA2 / HADD2 in SASS).
// ICB_HOST_EMULATION is defined only by tests/hostemu (the encoders compiled for the CPU to be checked against the
// oracle without a GPU); the library itself is never built that way.
#ifdef ICB_HOST_EMULATION
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) { return icb_emu::fma_f16x2(a, b, c, false); }
__device__ __forceinline__ uint32_t h2_fma_sat(uint32_t a, uint32_t b, uint32_t c) { return icb_emu::fma_f16x2(a, b, c, true); }
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) { return icb_emu::add_f16x2(a, b, false); }
__device__ __forceinline__ uint32_t h2_add_sat(uint32_t a, uint32_t b) { return icb_emu::add_f16x2(a, b, true); }
#else
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t h2_fma_sat(uint32_t a, uint32_t b, uint32_t c) {  // clamps each lane to [0, 1]
  uint32_t d;
  asm("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2_add_sat(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.sat.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
#endif

// Encodes the DXT5 alpha half from the top byte of each pixel (ComputeBaseAlphas, dxtc_compressor.cc:374-424;
// ComputeAlphaBits :427-479; bit layout Dxt5AlphaBits :103-158).  Returns the 8 output bytes as two words.
// `table` points at the crossing-point table (shared memory in the TMA kernel, global memory otherwise).
//
// Alphas and every threshold are integers <= 256, exact in fp16, so the block is processed two pixels per
// instruction: pixel pair (i, i+8) lives in one register as half2(1280 + a_i, 1280 + a_{i+8}) -- the bit pattern
// 0x6500 | a, so building it costs a byte permute and a mask, no conversion.
//   statistics  [a >= 1] and [a == 255] are saturating adds; "min over alphas that are not 0" and "max over
//               alphas that are not 255" become min/max over x - 255*[..] on the raw bit patterns (VIMNMX.U16x2);
//               the counts accumulate as 1024 + n so they can be read back from the mantissa.
//   indices     nearest-candidate search as seven crossings: t = sat(+-x + K_p) is 1 once the pixel has crossed,
//               acc += t * step_p accumulates the candidate index modulo 8 in the mantissa of 1024 + index.
__device__ __forceinline__ uint2 dxt5_encode_alpha(const uint32_t (&px)[16], bool one_pixel, const uint4 *table) {
  if (one_pixel) {  // window entirely outside the image: both endpoints = that alpha, all indices 0
    const uint32_t a = px[0] >> 24;
    return make_uint2(a | (a << 8), 0u);
  }
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (__byte_perm(px[i], px[i + 8], 0x7733) & 0x00ff00ffu) | 0x65006500u;

  // ---- statistics
  uint32_t fmin = 0xffffffffu, gmax = 0u, nz = 0x64006400u, n255 = 0x64006400u;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t z = h2_add_sat(x[i], 0xe500e500u);        // [a >= 1]      (x - 1280, clamped to [0,1])
    const uint32_t y = h2_add_sat(x[i], 0xe5fee5feu);        // [a == 255]    (x - 1534)
    fmin = __vminu2(fmin, h2_fma(z, 0xdbf8dbf8u, x[i]));     // 1025 + (a == 0 ? 255 : a)
    gmax = __vmaxu2(gmax, h2_fma(y, 0xdbf8dbf8u, x[i]));     // 1280 + (a == 255 ? 0 : a)
    nz = h2_add(nz, z);
    n255 = h2_add(n255, y);
  }
  int lo = static_cast<int>(min(fmin & 0xffffu, fmin >> 16) & 0x3ffu) - 1;
  int hi = static_cast<int>(max(gmax & 0xffffu, gmax >> 16) & 0x3ffu) - 256;
  const uint32_t num_nonzero = (nz & 0x3ffu) + ((nz >> 16) & 0x3ffu);
  const uint32_t num_opaque = (n255 & 0x3ffu) + ((n255 >> 16) & 0x3ffu);
  if (lo > hi) {  // every alpha is 0 or 255
    lo = 0;
    hi = 255;
  }
  uint32_t a0, a1;
  if (num_nonzero < 15u || num_opaque > 1u) {  // more than one fully transparent or fully opaque pixel
    a0 = lo;
    a1 = hi;
  } else {
    if (num_nonzero < 16u) lo = 0;
    if (num_opaque > 0u) hi = 255;
    a0 = hi;
    a1 = lo;
  }

  // ---- crossings for this (mode, |a0 - a1|)
  const bool six = a0 <= a1;  // 6-alpha mode: candidates 0 and 255 are explicit
  const uint32_t dist = __usad(a0, a1, 0u);
  const uint4 e = table[(six ? 0u : 256u) + dist];
  const uint32_t base = (six ? 0xe4ffu : 0x6402u) + a0;
  uint32_t K[7], S[7];
  K[0] = __dp4a(e.x, 0x00000001u, base); K[1] = __dp4a(e.x, 0x00000100u, base);
  K[2] = __dp4a(e.x, 0x00010000u, base); K[3] = __dp4a(e.x, 0x01000000u, base);
  K[4] = __dp4a(e.y, 0x00000001u, base); K[5] = __dp4a(e.y, 0x00000100u, base);
  K[6] = __dp4a(e.y, 0x00010000u, base);
  S[0] = __byte_perm(e.z, 0u, 0x0404); S[1] = __byte_perm(e.z, 0u, 0x1414);
  S[2] = __byte_perm(e.z, 0u, 0x2424); S[3] = __byte_perm(e.z, 0u, 0x3434);
  S[4] = __byte_perm(e.w, 0u, 0x0404); S[5] = __byte_perm(e.w, 0u, 0x1414);
  S[6] = __byte_perm(e.w, 0u, 0x2424);
  uint32_t start = 0x6400u;
  if (six) {
    // crossings against the explicit candidates: 0 (index 6, or 0 when a0 is itself 0) below the line and
    // 255 (index 7, or the line's last index when a1 is itself 255) above it
    const uint32_t a0_zero = a0 == 0u ? 1u : 0u, last = e.y >> 24;
    K[0] = 0xe4ffu + ((a0 + a0_zero + 1u) >> 1);
    S[0] = a0_zero ? 0u : 0x4000u;                               // 6 -> 0 is +2 (mod 8)
    K[6] = 0xe4ffu + ((a1 + 257u) >> 1);
    S[6] = a1 == 255u ? 0u : 0x4700u - (last << 8);              // last -> 7
    start += a0_zero ? 0u : 6u;
  }
  const uint32_t sign = six ? 0x3c003c00u : 0xbc00bc00u;         // +1 / -1 in both lanes
  start *= 0x10001u;
#pragma unroll
  for (int p = 0; p < 7; ++p) K[p] *= 0x10001u;
  S[0] *= six ? 0x10001u : 1u;  // table steps are already in both lanes; the on-the-fly ones are not
  S[6] *= six ? 0x10001u : 1u;

  // ---- indices: words 0..3 and 4..7 accumulate 3-bit codes at bit 3*(w & 3) of each lane
  uint32_t acc_a = 0, acc_b = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    uint32_t acc = start;
#pragma unroll
    for (int p = 0; p < 7; ++p) acc = h2_fma(h2_fma_sat(x[w], sign, K[p]), S[p], acc);
    const uint32_t code = acc & 0x00070007u;
    if (w < 4)
      acc_a += code << (3 * w);
    else
      acc_b += code << (3 * (w - 4));
  }
  // lanes: low = pixels 0..7, high = pixels 8..15; pixel n's code goes to bit 16 + 3n of the 64-bit block half
  const uint32_t word0 = a0 | (a1 << 8) | ((acc_a & 0xfffu) << 16) | (acc_b << 28);
  const uint32_t word1 = ((acc_b & 0xfffu) >> 4) | ((acc_a >> 16) << 8) | ((acc_b >> 16) << 20);
  return make_uint2(word0, word1);
}

}  // namespace icb
