// block4x4_generic.cuh -- the 4x4-block codecs' common pieces (codec ids, launch parameters, "encode 16 gathered
// pixels and store the block") and the generic image driver: any pointer / pitch / size, clamp-to-edge window gather
// as in Pixel4x4 (/root/reference/image_compression/internal/pixel4x4.h:45-67, internal/pixel4x4.cc:24-59), one
// block per thread in the reference's raster block order (internal/compressor4x4_helper.h:175-216, 479-520).
// The TMA drivers (block4x4_kernels.cuh) and the compressed-domain operations (blockops_kernels.cuh) build on this.
// Free of TMA / mbarrier code so that tests/hostemu can step through it on the CPU.
#pragma once
#include <cstdint>

#include "dxt_encode.cuh"
#include "etc1_encode.cuh"

namespace icb {

enum Codec4x4 : int { kCodecDxt1 = 0, kCodecDxt5 = 1, kCodecEtc1 = 2 };

struct Encode4x4Params {
  const uint8_t *src;     // pixel (0,0)
  uint8_t *dst;           // block (0,0) of the full output grid
  uint32_t height, width; // source image size in pixels
  uint32_t pitch;         // source bytes per row
  uint32_t grid_cols;     // blocks per output row (ceil(coded_w / 4))
  uint32_t row0, row1;    // block-row range [row0, row1) this launch encodes
  uint32_t col0, col1;    // block-column range [col0, col1)
  int swap_rb;            // kBGR / kBGRA
  int etc_strategy;
};

template <int kCodec>
struct CodecTraits;
template <>
struct CodecTraits<kCodecDxt1> {
  static constexpr int kBlockBytes = 8;
};
template <>
struct CodecTraits<kCodecDxt5> {
  static constexpr int kBlockBytes = 16;
};
template <>
struct CodecTraits<kCodecEtc1> {
  static constexpr int kBlockBytes = 8;
};

// Encodes 16 gathered pixels and stores the block.  px bytes are (c0,c1,c2,c3) in memory order; for 3-component
// sources c3 is zero.
// kFullWarp: the caller guarantees that all 32 lanes of the warp are here (lets warp votes skip the active-mask query).
// release(): called once, as soon as the encoder no longer needs `fetch` (see dxt_encode.cuh).
template <int kCodec, bool kFullWarp = false, typename Fetch, typename Release = NoRelease>
__device__ __forceinline__ void encode_and_store(const uint32_t (&px)[16], Fetch fetch, bool one_pixel, int swap_rb,
                                                 int etc_strategy, const uint4 *alpha_table, uint8_t *out,
                                                 Release release = Release()) {
  if constexpr (kCodec == kCodecDxt1) {
#ifdef ICB_PROBE_ENCODER
    // Measurement builds only (tools/build_variants.sh): no encoder, every pixel still loaded and 8 bytes stored, to
    // find the streaming ceiling of the TMA driver itself.
    uint32_t a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a ^= px[i];
      b += px[i + 8];
    }
    release();
    *reinterpret_cast<uint2 *>(out) = make_uint2(a, b);
#else
    const uint2 c = dxt1_encode_block<kFullWarp>(px, swap_rb != 0, false, fetch, release);
    *reinterpret_cast<uint2 *>(out) = c;
#endif
  } else if constexpr (kCodec == kCodecDxt5) {
    // Colour first: it releases the staged pixels early.  (Tried in round 2: alpha endpoints first with a
    // prefetch.global.L1 of the crossing-table row behind the colour half -- CCTL.E.PF1 with 32 different addresses per
    // warp doubled the kernel time, 90 -> 189 us; the row loads hit L1 99 % of the time anyway.  Also tried: half of the
    // CTA's warps running the alpha half BEFORE the colour half, to stagger the integer-heavy and FP32-heavy phases
    // across a scheduler's warps -- 88 -> 104 us, the late release of the staged pixels starves the two-stage ring.)
    // (Tried in round 2, second half: extracting the eight alpha lane pairs BEFORE the colour half, so that it keeps
    // eight registers alive instead of the sixteen pixel words -- 83.4 against 82.6 us, with or without a fifth CTA.)
    const uint2 c = dxt1_encode_block<kFullWarp>(px, swap_rb != 0, true, fetch, release);
    const uint2 a = dxt5_encode_alpha<kFullWarp>(px, one_pixel, alpha_table);
    *reinterpret_cast<uint4 *>(out) = make_uint4(a.x, a.y, c.x, c.y);
  } else {
    release();  // every pixel is already in registers
    const uint2 e = etc1_encode_block(px, etc_strategy);
    *reinterpret_cast<uint2 *>(out) = e;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Generic driver
// ---------------------------------------------------------------------------------------------------------

template <int kNcomp>
__device__ __forceinline__ uint32_t load_pixel_clamped(const Encode4x4Params &p, uint32_t y, uint32_t x) {
  y = min(y, p.height - 1u);
  x = min(x, p.width - 1u);
  const uint8_t *q = p.src + static_cast<size_t>(y) * p.pitch + static_cast<size_t>(x) * kNcomp;
  uint32_t v = q[0] | (static_cast<uint32_t>(q[1]) << 8) | (static_cast<uint32_t>(q[2]) << 16);
  if (kNcomp == 4) v |= static_cast<uint32_t>(q[3]) << 24;
  return v;
}

template <int kCodec, int kNcomp>
__global__ void __launch_bounds__(128) encode4x4_generic_kernel(const Encode4x4Params p) {
  const uint32_t ncols = p.col1 - p.col0;
  const uint64_t total = static_cast<uint64_t>(p.row1 - p.row0) * ncols;
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t br = p.row0 + static_cast<uint32_t>(t / ncols);
    const uint32_t bc = p.col0 + static_cast<uint32_t>(t % ncols);
    uint32_t px[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) px[i] = load_pixel_clamped<kNcomp>(p, 4u * br + (i >> 2), 4u * bc + (i & 3));
    auto fetch = [&](uint32_t i) { return load_pixel_clamped<kNcomp>(p, 4u * br + (i >> 2), 4u * bc + (i & 3u)); };
    const bool one_pixel = 4u * br >= p.height && 4u * bc >= p.width;  // pixel4x4.cc:58
    uint8_t *out = p.dst + (static_cast<size_t>(br) * p.grid_cols + bc) * CodecTraits<kCodec>::kBlockBytes;
    encode_and_store<kCodec>(px, fetch, one_pixel, p.swap_rb, p.etc_strategy,
                             reinterpret_cast<const uint4 *>(g_dxt5_alpha_table), out);
  }
}

}  // namespace icb
