// blockops_kernels.cuh -- compressed-domain operations on DXT1 / DXT5 / ETC1 block streams (SURVEY.md section 8f,
// ranks 3-4): the callers on the far side of the compress path, all re-using the hot-path encoders.
//
//   downsample4x4_kernel   Compressor4x4Helper::Downsample + DownsampleBlocks2x2/2x1/1x2
//                          (internal/compressor4x4_helper.h:264-391, 594-636): decode 2x2 source blocks, average
//                          2x2 pixels with truncation (internal/color_util.h:335-380), re-encode one block
//   pad4x4_kernel          Compressor4x4Helper::Pad (internal/compressor4x4_helper.h:393-477) with the codecs'
//                          column / row / corner pad blocks (internal/dxtc_compressor.cc:594-696,
//                          internal/etc_compressor.cc:645-698)
//   fill_solid4x4_kernel   CreateSolidImage (internal/dxtc_compressor.cc:820-840, internal/etc_compressor.cc:595-617,
//                          772-785 -> compressor4x4_helper.h:522-545)
//   transcode_dxt1_to_etc1_kernel  TranscodeDxt1ToEtc1 (internal/dxtc_to_etc_transcoder.cc:29-40), in place
//
// One thread per OUTPUT block everywhere; a warp reads and writes contiguous runs of blocks.  These are not
// roofline kernels (each is dominated by the encoder it calls); they exist so that a block stream that lives in
// HBM never has to visit the host to get its mip chain, its power-of-two padding or its ETC1 twin.
#pragma once
#include <cstdint>
#include <type_traits>

#include "block4x4_generic.cuh"
#include "decode4x4_kernels.cuh"

namespace icb {

constexpr int kBlockOpThreads = 128;

// Encodes a 16-pixel window held in registers.  The DXT encoder looks two pixels up by index; a register-indexed
// array would spill, so the window is parked in this thread's row of shared memory (stride 17 words: conflict-free).
template <int kCodec>
__device__ __forceinline__ void encode_window(uint32_t (&px)[16], int etc_strategy, uint32_t (*park)[17], uint8_t *out) {
  if constexpr (kCodec == kCodecEtc1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) px[i] &= 0x00ffffffu;  // the ETC encoder wants a zero top byte
  }
  uint32_t *mine = park[threadIdx.x];
#pragma unroll
  for (int i = 0; i < 16; ++i) mine[i] = px[i];
  auto fetch = [&](uint32_t i) { return mine[i]; };
  encode_and_store<kCodec>(px, fetch, false, 0, etc_strategy, reinterpret_cast<const uint4 *>(g_dxt5_alpha_table), out);
}

// floor((a + b + c + d) / 4) on each of the four bytes.
__device__ __forceinline__ uint32_t average4_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  const uint32_t even = (a & 0x00ff00ffu) + (b & 0x00ff00ffu) + (c & 0x00ff00ffu) + (d & 0x00ff00ffu);
  const uint32_t odd = ((a >> 8) & 0x00ff00ffu) + ((b >> 8) & 0x00ff00ffu) + ((c >> 8) & 0x00ff00ffu) + ((d >> 8) & 0x00ff00ffu);
  return ((even >> 2) & 0x00ff00ffu) | (((odd >> 2) & 0x00ff00ffu) << 8);
}

struct Downsample4x4Params {
  const uint8_t *in;          // in_rows * in_cols blocks
  uint8_t *out;               // out_rows * out_cols blocks
  uint32_t in_rows, in_cols;  // ceil(uncompressed / 4) -- the grid the reference walks
  uint32_t out_rows, out_cols;
  uint32_t height, width;     // uncompressed size (only read by the single-block case)
  int etc_strategy;
};

template <int kCodec>
__global__ void __launch_bounds__(kBlockOpThreads) downsample4x4_kernel(const Downsample4x4Params p) {
  constexpr int kBlockBytes = CodecTraits<kCodec>::kBlockBytes;
  __shared__ uint32_t park[kBlockOpThreads][17];
  const bool many_rows = p.in_rows > 1, many_cols = p.in_cols > 1;
  const uint64_t total = static_cast<uint64_t>(p.out_rows) * p.out_cols;
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t orow = static_cast<uint32_t>(t / p.out_cols), ocol = static_cast<uint32_t>(t % p.out_cols);
    uint32_t win[16];
#pragma unroll
    for (int qy = 0; qy < 2; ++qy) {
#pragma unroll
      for (int qx = 0; qx < 2; ++qx) {
        // Quadrant (qy,qx) of the output window comes from source block (2*orow + qy, 2*ocol + qx); with a single
        // row (column) of source blocks the same block feeds both halves (DownsampleBlocks1x2 / 2x1).
        const uint32_t srow = many_rows ? 2u * orow + qy : 0u, scol = many_cols ? 2u * ocol + qx : 0u;
        uint32_t px[16];
        decode_block<kCodec>(p.in + (static_cast<size_t>(srow) * p.in_cols + scol) * kBlockBytes, false, px);
        if (!many_rows && !many_cols) {
          // One block in all: an image 1, 2 or 4 pixels wide/high is first stretched to 4x4 by replication
          // (compressor4x4_helper.h:340-366); 3 is refused on the host.
          if (p.width == 1u) {
#pragma unroll
            for (int r = 0; r < 4; ++r) px[4 * r + 1] = px[4 * r + 2] = px[4 * r + 3] = px[4 * r];
          } else if (p.width == 2u) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              px[4 * r + 2] = px[4 * r];
              px[4 * r + 3] = px[4 * r + 1];
            }
          }
          if (p.height == 1u) {
#pragma unroll
            for (int c = 0; c < 4; ++c) px[4 + c] = px[8 + c] = px[12 + c] = px[c];
          } else if (p.height == 2u) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              px[8 + c] = px[c];
              px[12 + c] = px[4 + c];
            }
          }
        }
#pragma unroll
        for (int sy = 0; sy < 2; ++sy) {
#pragma unroll
          for (int sx = 0; sx < 2; ++sx) {
            const int s = 8 * sy + 2 * sx;
            win[4 * (2 * qy + sy) + 2 * qx + sx] = average4_bytes(px[s], px[s + 1], px[s + 4], px[s + 5]);
          }
        }
      }
    }
    encode_window<kCodec>(win, p.etc_strategy, park, p.out + t * kBlockBytes);
  }
}

// ---------------------------------------------------------------------------------------------------------

// kind: 0 = replicate the block's last column, 1 = its last row, 2 = its bottom-right pixel.
template <int kCodec>
__device__ __forceinline__ void make_pad_block(const uint8_t *src, int kind, int etc_strategy, uint32_t (*park)[17],
                                               uint8_t *out) {
  if constexpr (kCodec == kCodecEtc1) {
    uint32_t px[16];
    decode_block<kCodecEtc1>(src, false, px);
    if (kind == 2) {
      // CreateSolidBlock of the corner pixel: differential mode, zero delta, codeword 0, all indices 0.
      // (The reference computes a colour adjusted by the smallest codebook entry and then does not use it:
      // etc_compressor.cc:600-611.)
      const uint32_t c = px[15];
      const uint32_t hi = 2u | (((c & 0xffu) >> 3) << 27) | ((((c >> 8) & 0xffu) >> 3) << 19) | ((((c >> 16) & 0xffu) >> 3) << 11);
      *reinterpret_cast<uint2 *>(out) = make_uint2(__byte_perm(hi, 0u, 0x0123), 0u);
      return;
    }
    uint32_t win[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) win[i] = kind == 0 ? px[(i & 12) + 3] : px[12 + (i & 3)];
    encode_window<kCodecEtc1>(win, etc_strategy, park, out);
  } else {
    constexpr bool kDxt5 = kCodec == kCodecDxt5;
    const uint2 colour = *reinterpret_cast<const uint2 *>(src + (kDxt5 ? 8 : 0));
    // 2-bit codes, byte y = row y, pixel x at bit 2x (dxtc_compressor.cc:546-554, 594-696)
    uint32_t bits = colour.y;
    if (kind == 0) bits = ((bits >> 6) & 0x03030303u) * 0x55u;
    if (kind == 1) bits = (bits >> 24) * 0x01010101u;
    if (kind == 2) bits = ((bits >> 30) & 3u) * 0x55555555u;
    if constexpr (kDxt5) {
      const uint2 a = *reinterpret_cast<const uint2 *>(src);
      const uint64_t codes = (static_cast<uint64_t>(a.y) << 16) | (a.x >> 16);  // 16 x 3 bits
      uint64_t outc = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int from = kind == 0 ? (i & 12) + 3 : kind == 1 ? 12 + (i & 3) : 15;
        outc |= ((codes >> (3 * from)) & 7ull) << (3 * i);
      }
      *reinterpret_cast<uint4 *>(out) = make_uint4((a.x & 0xffffu) | (static_cast<uint32_t>(outc) << 16),
                                                   static_cast<uint32_t>(outc >> 16), colour.x, bits);
    } else {
      *reinterpret_cast<uint2 *>(out) = make_uint2(colour.x, bits);
    }
  }
}

struct Pad4x4Params {
  const uint8_t *in;  // in_rows * in_cols blocks
  uint8_t *out;       // out_rows * out_cols blocks, out_rows >= in_rows, out_cols >= in_cols
  uint32_t in_rows, in_cols, out_rows, out_cols;
  int etc_strategy;
};

template <int kCodec>
__global__ void __launch_bounds__(kBlockOpThreads) pad4x4_kernel(const Pad4x4Params p) {
  constexpr int kBlockBytes = CodecTraits<kCodec>::kBlockBytes;
  using BlockWord = typename std::conditional<kBlockBytes == 16, uint4, uint2>::type;
  __shared__ uint32_t park[kBlockOpThreads][17];
  const uint64_t total = static_cast<uint64_t>(p.out_rows) * p.out_cols;
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t r = static_cast<uint32_t>(t / p.out_cols), c = static_cast<uint32_t>(t % p.out_cols);
    const bool below = r >= p.in_rows, right = c >= p.in_cols;
    const uint32_t sr = below ? p.in_rows - 1u : r, sc = right ? p.in_cols - 1u : c;
    const uint8_t *src = p.in + (static_cast<size_t>(sr) * p.in_cols + sc) * kBlockBytes;
    uint8_t *out = p.out + t * kBlockBytes;
    if (!below && !right)
      *reinterpret_cast<BlockWord *>(out) = *reinterpret_cast<const BlockWord *>(src);
    else
      make_pad_block<kCodec>(src, below ? (right ? 2 : 1) : 0, p.etc_strategy, park, out);
  }
}

// ---------------------------------------------------------------------------------------------------------

// The block CreateSolidImage replicates.  colour bytes (c0,c1,c2,alpha) as passed by the caller: the reference
// quantises them in the order given, whatever the format says (dxtc_compressor.cc:826-837).
template <int kCodec>
__global__ void __launch_bounds__(256) fill_solid4x4_kernel(uint8_t *out, uint64_t num_blocks, uint32_t colour) {
  const uint32_t c0 = colour & 0xffu, c1 = (colour >> 8) & 0xffu, c2 = (colour >> 16) & 0xffu, a = colour >> 24;
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < num_blocks;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if constexpr (kCodec == kCodecEtc1) {
      const uint32_t hi = 2u | ((c0 >> 3) << 27) | ((c1 >> 3) << 19) | ((c2 >> 3) << 11);
      reinterpret_cast<uint2 *>(out)[t] = make_uint2(__byte_perm(hi, 0u, 0x0123), 0u);
    } else {
      const uint32_t q = to_565(c0, c1, c2);
      if constexpr (kCodec == kCodecDxt5)
        reinterpret_cast<uint4 *>(out)[t] = make_uint4(a | (a << 8), 0u, q | (q << 16), 0u);
      else
        reinterpret_cast<uint2 *>(out)[t] = make_uint2(q | (q << 16), 0u);
    }
  }
}

// DXT1 block -> 16 RGB pixels (no channel swap) -> ETC1 block with the heuristic strategy, written over the input.
__global__ void __launch_bounds__(kBlockOpThreads) transcode_dxt1_to_etc1_kernel(uint8_t *blocks, uint64_t num_blocks) {
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < num_blocks;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    uint32_t px[16];
    decode_block<kCodecDxt1>(blocks + t * 8, false, px);
#pragma unroll
    for (int i = 0; i < 16; ++i) px[i] &= 0x00ffffffu;
    reinterpret_cast<uint2 *>(blocks)[t] = etc1_encode_block(px, kEtcHeuristic);
  }
}

}  // namespace icb
