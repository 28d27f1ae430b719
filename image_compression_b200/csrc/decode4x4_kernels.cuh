// decode4x4_kernels.cuh -- block decoders for DXT1, DXT5 and ETC1 (the first "next" row after the compress path,
// SURVEY.md section 8f): Compressor4x4Helper::Decompress (internal/compressor4x4_helper.h:218-262) over
// DecodeDxt1Block / DecodeDxt5Block (internal/dxtc_compressor.cc:167-267) and Etc1BlockDecoder
// (internal/etc_compressor.cc:198-289).  One thread per block; a warp reads 256/512 contiguous bytes of blocks and
// writes four pixel rows of 384/512 contiguous bytes.  Byte-identical to the reference for every bit pattern,
// including blocks no encoder produces (3-colour DXT1 blocks, ETC1 differential overflow).
#pragma once
#include <cstdint>

#include "dxt_encode.cuh"
#include "etc1_encode.cuh"

namespace icb {

struct Decode4x4Params {
  const uint8_t *blocks;  // block_rows * block_cols blocks, raster order
  uint8_t *dst;           // pixel (0,0)
  uint32_t height, width; // pixels written: rows/columns beyond are dropped
  uint32_t pitch;         // destination bytes per row
  uint32_t block_cols;    // blocks per block row IN THE STREAM (the reference uses ceil(width / 4))
  uint32_t block_rows;
  int swap_rb;
};

// Four palette entries of a DXT colour block as (r,g,b) bytes in DESTINATION order, top byte 0.
__device__ __forceinline__ void dxt_decode_palette(uint2 colour, bool swap_rb, bool always4, uint32_t (&pal)[4]) {
  const uint32_t c0 = colour.x & 0xffffu, c1 = colour.x >> 16;
  uint32_t ch[2][3];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t c = k ? c1 : c0;
    const uint32_t r = expand5(c >> 11), g = expand6((c >> 5) & 63u), b = expand5(c & 31u);
    ch[k][0] = swap_rb ? b : r;
    ch[k][1] = g;
    ch[k][2] = swap_rb ? r : b;
  }
  uint32_t p2 = 0, p3 = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    uint32_t v2, v3;
    if (c0 == c1) {
      v2 = v3 = ch[1][i];
    } else if (always4 || c0 > c1) {
      v2 = div3_small(2u * ch[0][i] + ch[1][i]);
      v3 = div3_small(ch[0][i] + 2u * ch[1][i]);
    } else {  // three-colour mode: midpoint and black
      v2 = (ch[0][i] + ch[1][i]) >> 1;
      v3 = 0;
    }
    p2 |= v2 << (8 * i);
    p3 |= v3 << (8 * i);
  }
  pal[0] = ch[0][0] | (ch[0][1] << 8) | (ch[0][2] << 16);
  pal[1] = ch[1][0] | (ch[1][1] << 8) | (ch[1][2] << 16);
  pal[2] = p2;
  pal[3] = p3;
}

// One pixel row of a decoded block, channel-planar: byte x of c0 / c1 / c2 / a is that channel of pixel x (channels in
// DESTINATION byte order).  The decoders select four pixels per instruction -- the palette of a channel is one
// register (two for the eight DXT5 alphas or the two ETC1 sub-blocks), the row's codes spread to nibbles are a PRMT
// selector -- instead of walking a select chain per pixel.
struct RowPlanes {
  uint32_t c0, c1, c2, a;
};

// 2-bit codes -> nibbles: 16 bits (8 codes) in, 32 bits out.
__device__ __forceinline__ uint32_t spread2_to_nibbles(uint32_t x) {
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  return (x | (x << 2)) & 0x33333333u;
}
// 3-bit codes -> nibbles: 24 bits (8 codes) in, 32 bits out.
__device__ __forceinline__ uint32_t spread3_to_nibbles(uint32_t x) {
  x = (x & 0x00000fffu) | ((x & 0x00fff000u) << 4);
  x = (x & 0x003f003fu) | ((x & 0x0fc00fc0u) << 2);
  return (x & 0x07070707u) | ((x & 0x38383838u) << 1);
}

template <int kCodec>
__device__ __forceinline__ void decode_block_rows(const uint8_t *blk, bool swap_rb, RowPlanes (&rows)[4]) {
  if constexpr (kCodec == 2) {
    const uint2 raw = *reinterpret_cast<const uint2 *>(blk);
    const uint32_t hi = __byte_perm(raw.x, 0u, 0x0123), lo = __byte_perm(raw.y, 0u, 0x0123);
    const bool flip = hi & 1u, diff = hi & 2u;
    const int cw[2] = {static_cast<int>((hi >> 5) & 7u), static_cast<int>((hi >> 2) & 7u)};
    uint32_t cand[2][3];  // [sub-block][channel]: bytes = base + {small, large, -small, -large}, clamped
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int base[2];
      if (diff) {
        const int b5 = static_cast<int>((hi >> (27 - 8 * k)) & 31u);
        int d3 = static_cast<int>((hi >> (24 - 8 * k)) & 7u);
        d3 = (d3 & 4) ? d3 - 8 : d3;
        const int second = b5 + d3;  // may leave 0..31 for blocks no encoder produces; same arithmetic as the reference
        base[0] = (b5 << 3) | (b5 >> 2);
        base[1] = (second * 8) | ((second >> 2) & 7);
      } else {
        base[0] = static_cast<int>((hi >> (28 - 8 * k)) & 15u) * 17;
        base[1] = static_cast<int>((hi >> (24 - 8 * k)) & 15u) * 17;
      }
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int ms = etc_small(cw[s2]), ml = etc_large(cw[s2]);
        cand[s2][k] = static_cast<uint32_t>(__viaddmin_s32_relu(base[s2], ms, 255)) |
                      (static_cast<uint32_t>(__viaddmin_s32_relu(base[s2], ml, 255)) << 8) |
                      (static_cast<uint32_t>(__viaddmin_s32_relu(base[s2], -ms, 255)) << 16) |
                      (static_cast<uint32_t>(__viaddmin_s32_relu(base[s2], -ml, 255)) << 24);
      }
    }
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      // pixel (y,x) sits at bit 4x+y (index LSB: large magnitude) and 16+4x+y (MSB: negative);
      // selector nibble = index + 4 * sub-block
      const uint32_t second = flip ? (y >= 2 ? 0x4444u : 0u) : 0x4400u;
      const uint32_t sel = ((lo >> y) & 0x1111u) | (((lo >> (16 + y)) & 0x1111u) << 1) | second;
      rows[y].c0 = __byte_perm(cand[0][0], cand[1][0], sel);
      rows[y].c1 = __byte_perm(cand[0][1], cand[1][1], sel);
      rows[y].c2 = __byte_perm(cand[0][2], cand[1][2], sel);
      rows[y].a = 0xffffffffu;
    }
  } else {
    constexpr bool kDxt5 = kCodec == 1;
    const uint2 colour = *reinterpret_cast<const uint2 *>(blk + (kDxt5 ? 8 : 0));
    uint32_t pal[4];
    dxt_decode_palette(colour, swap_rb, kDxt5, pal);
    // channel-planar palette: byte j of plane i = channel i of palette entry j
    const uint32_t lo01 = __byte_perm(pal[0], pal[1], 0x5140), hi01 = __byte_perm(pal[0], pal[1], 0x7362);
    const uint32_t lo23 = __byte_perm(pal[2], pal[3], 0x5140), hi23 = __byte_perm(pal[2], pal[3], 0x7362);
    // lo01 = [c0 of entry 0, c0 of 1, c1 of 0, c1 of 1], hi01 = [c2 of 0, c2 of 1, -, -]; likewise for entries 2, 3
    const uint32_t plane0 = __byte_perm(lo01, lo23, 0x5410), plane1 = __byte_perm(lo01, lo23, 0x7632);
    const uint32_t plane2 = __byte_perm(hi01, hi23, 0x5410);
    const uint32_t sel01 = spread2_to_nibbles(colour.y & 0xffffu), sel23 = spread2_to_nibbles(colour.y >> 16);
    uint32_t asel01 = 0, asel23 = 0, alpha_lo = 0, alpha_hi = 0;
    if constexpr (kDxt5) {
      const uint2 a = *reinterpret_cast<const uint2 *>(blk);
      const uint32_t a0 = a.x & 255u, a1 = (a.x >> 8) & 255u;
      uint32_t alpha[8];
      alpha[0] = a0;
      alpha[1] = a1;
      if (a0 > a1) {
#pragma unroll
        for (int k = 1; k <= 6; ++k) alpha[1 + k] = ((7 - k) * a0 + k * a1) / 7u;
      } else {
#pragma unroll
        for (int k = 1; k <= 4; ++k) alpha[1 + k] = ((5 - k) * a0 + k * a1) / 5u;
        alpha[6] = 0;
        alpha[7] = 255;
      }
      alpha_lo = alpha[0] | (alpha[1] << 8) | (alpha[2] << 16) | (alpha[3] << 24);
      alpha_hi = alpha[4] | (alpha[5] << 8) | (alpha[6] << 16) | (alpha[7] << 24);
      // 48 bits of 3-bit codes start at byte 2 of the block: pixels 0..7 in the first 24, 8..15 in the rest
      asel01 = spread3_to_nibbles((a.x >> 16) | ((a.y & 0xffu) << 16));
      asel23 = spread3_to_nibbles(a.y >> 8);
    }
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const uint32_t sel = (y < 2 ? sel01 : sel23) >> (16 * (y & 1));  // PRMT reads the low four nibbles
      rows[y].c0 = __byte_perm(plane0, 0u, sel);
      rows[y].c1 = __byte_perm(plane1, 0u, sel);
      rows[y].c2 = __byte_perm(plane2, 0u, sel);
      if constexpr (kDxt5) {
        rows[y].a = __byte_perm(alpha_lo, alpha_hi, (y < 2 ? asel01 : asel23) >> (16 * (y & 1)));
      } else {
        rows[y].a = 0xffffffffu;
      }
    }
  }
}

// Decodes one block to 16 pixels, bytes (c0,c1,c2,alpha); alpha is 255 for the 3-component codecs.
template <int kCodec>
__device__ __forceinline__ void decode_block(const uint8_t *blk, bool swap_rb, uint32_t (&px)[16]) {
  RowPlanes rows[4];
  decode_block_rows<kCodec>(blk, swap_rb, rows);
#pragma unroll
  for (int y = 0; y < 4; ++y) {  // 4x4 byte transpose: planes -> packed pixels
    const uint32_t t0 = __byte_perm(rows[y].c0, rows[y].c1, 0x5140), t1 = __byte_perm(rows[y].c0, rows[y].c1, 0x7362);
    const uint32_t u0 = __byte_perm(rows[y].c2, rows[y].a, 0x5140), u1 = __byte_perm(rows[y].c2, rows[y].a, 0x7362);
    px[4 * y + 0] = __byte_perm(t0, u0, 0x5410);
    px[4 * y + 1] = __byte_perm(t0, u0, 0x7632);
    px[4 * y + 2] = __byte_perm(t1, u1, 0x5410);
    px[4 * y + 3] = __byte_perm(t1, u1, 0x7632);
  }
}

template <int kCodec>
__global__ void __launch_bounds__(128) decode4x4_kernel(const Decode4x4Params p) {
  constexpr int kBlockBytes = kCodec == 1 ? 16 : 8;
  constexpr int kNcomp = kCodec == 1 ? 4 : 3;
  // x: 128 consecutive block columns per CTA; y: block rows, strided by the grid
  const uint32_t bc = blockIdx.x * blockDim.x + threadIdx.x;
  if (bc >= p.block_cols || 4u * bc >= p.width) return;  // column entirely outside the destination
  const uint32_t cols = min(4u, p.width - 4u * bc);
  const bool aligned = (reinterpret_cast<uintptr_t>(p.dst) % (kNcomp == 4 ? 16 : 4) == 0) && p.pitch % (kNcomp == 4 ? 16 : 4) == 0;
  const bool wide = cols == 4u && aligned;  // 4 * bc * kNcomp is a multiple of 16 (RGBA) or 12 (RGB)
  for (uint32_t br = blockIdx.y; br < p.block_rows && 4u * br < p.height; br += gridDim.y) {
    RowPlanes rows[4];
    decode_block_rows<kCodec>(p.blocks + (static_cast<size_t>(br) * p.block_cols + bc) * kBlockBytes, p.swap_rb != 0, rows);
    const uint32_t nrows = min(4u, p.height - 4u * br);
    uint8_t *origin = p.dst + static_cast<size_t>(4u * br) * p.pitch + static_cast<size_t>(4u * bc) * kNcomp;
#pragma unroll
    for (uint32_t y = 0; y < 4; ++y) {
      if (y >= nrows) break;
      uint8_t *row = origin + static_cast<size_t>(y) * p.pitch;
      const uint32_t t0 = __byte_perm(rows[y].c0, rows[y].c1, 0x5140), t1 = __byte_perm(rows[y].c0, rows[y].c1, 0x7362);
      if (wide) {
        if constexpr (kNcomp == 4) {
          const uint32_t u0 = __byte_perm(rows[y].c2, rows[y].a, 0x5140), u1 = __byte_perm(rows[y].c2, rows[y].a, 0x7362);
          *reinterpret_cast<uint4 *>(row) = make_uint4(__byte_perm(t0, u0, 0x5410), __byte_perm(t0, u0, 0x7632),
                                                       __byte_perm(t1, u1, 0x5410), __byte_perm(t1, u1, 0x7632));
        } else {  // 12 bytes, pixels P Q R S: [P.c0 P.c1 P.c2 Q.c0] [Q.c1 Q.c2 R.c0 R.c1] [R.c2 S.c0 S.c1 S.c2]
          uint32_t *w = reinterpret_cast<uint32_t *>(row);
          w[0] = __byte_perm(t0, rows[y].c2, 0x2410);
          w[1] = __byte_perm(__byte_perm(t0, rows[y].c2, 0x0053), t1, 0x5410);
          w[2] = __byte_perm(t1, rows[y].c2, 0x7326);
        }
      } else {
        const uint32_t planes[4] = {rows[y].c0, rows[y].c1, rows[y].c2, rows[y].a};
        for (uint32_t x = 0; x < cols; ++x)
          for (int k = 0; k < kNcomp; ++k) row[x * kNcomp + k] = static_cast<uint8_t>(planes[k] >> (8 * x));
      }
    }
  }
}

}  // namespace icb
