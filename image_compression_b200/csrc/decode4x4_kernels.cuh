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

// Decodes one block to 16 pixels, bytes (c0,c1,c2,alpha); alpha is 255 for the 3-component codecs.
template <int kCodec>
__device__ __forceinline__ void decode_block(const uint8_t *blk, bool swap_rb, uint32_t (&px)[16]) {
  if constexpr (kCodec == 2) {
    const uint2 raw = *reinterpret_cast<const uint2 *>(blk);
    const uint32_t hi = __byte_perm(raw.x, 0u, 0x0123), lo = __byte_perm(raw.y, 0u, 0x0123);
    const bool flip = hi & 1u, diff = hi & 2u;
    const int cw[2] = {static_cast<int>((hi >> 5) & 7u), static_cast<int>((hi >> 2) & 7u)};
    int base[2][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (diff) {
        const int b5 = static_cast<int>((hi >> (27 - 8 * k)) & 31u);
        int d3 = static_cast<int>((hi >> (24 - 8 * k)) & 7u);
        d3 = (d3 & 4) ? d3 - 8 : d3;
        const int second = b5 + d3;  // may leave 0..31 for blocks no encoder produces; same arithmetic as the reference
        base[0][k] = (b5 << 3) | (b5 >> 2);
        base[1][k] = (second * 8) | ((second >> 2) & 7);
      } else {
        base[0][k] = static_cast<int>((hi >> (28 - 8 * k)) & 15u) * 17;
        base[1][k] = static_cast<int>((hi >> (24 - 8 * k)) & 15u) * 17;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int y = i >> 2, x = i & 3, p = 4 * x + y;
      const uint32_t idx = ((lo >> p) & 1u) | (((lo >> (p + 16)) & 1u) << 1);
      const int second = flip ? (y >= 2) : (x >= 2);
      const int mag = (idx & 1u) ? etc_large(cw[second]) : etc_small(cw[second]);
      const int m = (idx & 2u) ? -mag : mag;
      px[i] = etc_candidate(base[second][0], base[second][1], base[second][2], m) | 0xff000000u;
    }
  } else {
    constexpr bool kDxt5 = kCodec == 1;
    const uint2 colour = *reinterpret_cast<const uint2 *>(blk + (kDxt5 ? 8 : 0));
    uint32_t pal[4];
    dxt_decode_palette(colour, swap_rb, kDxt5, pal);
    uint32_t alpha[8];
    uint64_t abits = 0;
    if constexpr (kDxt5) {
      const uint2 a = *reinterpret_cast<const uint2 *>(blk);
      const uint32_t a0 = a.x & 255u, a1 = (a.x >> 8) & 255u;
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
      abits = (static_cast<uint64_t>(a.y) << 16) | (a.x >> 16);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t code = (colour.y >> (2 * i)) & 3u;
      uint32_t c = code == 0 ? pal[0] : code == 1 ? pal[1] : code == 2 ? pal[2] : pal[3];
      if constexpr (kDxt5) {
        const uint32_t ac = static_cast<uint32_t>(abits >> (3 * i)) & 7u;
        uint32_t av = alpha[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) av = ac == k ? alpha[k] : av;
        c |= av << 24;
      } else {
        c |= 0xff000000u;
      }
      px[i] = c;
    }
  }
}

template <int kCodec>
__global__ void __launch_bounds__(128) decode4x4_kernel(const Decode4x4Params p) {
  constexpr int kBlockBytes = kCodec == 1 ? 16 : 8;
  constexpr int kNcomp = kCodec == 1 ? 4 : 3;
  const uint64_t total = static_cast<uint64_t>(p.block_rows) * p.block_cols;
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t br = static_cast<uint32_t>(t / p.block_cols), bc = static_cast<uint32_t>(t % p.block_cols);
    if (4u * br >= p.height || 4u * bc >= p.width) continue;  // block entirely outside the destination
    uint32_t px[16];
    decode_block<kCodec>(p.blocks + t * kBlockBytes, p.swap_rb != 0, px);
    const uint32_t rows = min(4u, p.height - 4u * br), cols = min(4u, p.width - 4u * bc);
    uint8_t *origin = p.dst + static_cast<size_t>(4u * br) * p.pitch + static_cast<size_t>(4u * bc) * kNcomp;
    const bool wide = cols == 4u && (reinterpret_cast<uintptr_t>(origin) % (kNcomp == 4 ? 16 : 4) == 0) &&
                      p.pitch % (kNcomp == 4 ? 16 : 4) == 0;
    for (uint32_t y = 0; y < rows; ++y) {
      uint8_t *row = origin + static_cast<size_t>(y) * p.pitch;
      const uint32_t a = px[4 * y], b = px[4 * y + 1], c = px[4 * y + 2], d = px[4 * y + 3];
      if (wide) {
        if constexpr (kNcomp == 4) {
          *reinterpret_cast<uint4 *>(row) = make_uint4(a, b, c, d);
        } else {  // 12 bytes = three words
          uint32_t *w = reinterpret_cast<uint32_t *>(row);
          w[0] = (a & 0x00ffffffu) | (b << 24);
          w[1] = ((b >> 8) & 0xffffu) | (c << 16);
          w[2] = ((c >> 16) & 0xffu) | (d << 8);
        }
      } else {
        const uint32_t v[4] = {a, b, c, d};
        for (uint32_t x = 0; x < cols; ++x)
          for (int k = 0; k < kNcomp; ++k) row[x * kNcomp + k] = static_cast<uint8_t>(v[x] >> (8 * k));
      }
    }
  }
}

}  // namespace icb
