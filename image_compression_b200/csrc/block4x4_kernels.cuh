// block4x4_kernels.cuh -- image-level drivers for the 4x4-block codecs (DXT1, DXT5, ETC1).
//
// These kernels ARE the reference's block loop, Compressor4x4Helper::Compress / CompressAndPad
// (/root/reference/image_compression/internal/compressor4x4_helper.h:175-216, 479-520): one 4x4 window per
// block in raster order, window gather with clamp-to-edge replication as in Pixel4x4
// (internal/pixel4x4.h:45-67, internal/pixel4x4.cc:24-59).
//
// Two drivers:
//   encode4x4_tma_kernel      the fast path.  Persistent CTAs; one elected producer thread streams 2-D pixel
//                             tiles HBM -> shared memory with TMA (cp.async.bulk.tensor) through a ring of
//                             mbarrier-guarded stages; consumer warps read their block's four rows with 128-bit
//                             (or 3x32-bit for RGB888) conflict-free shared loads, encode in registers and write
//                             8/16 B per lane, coalesced.  Needs a 16-byte aligned base and pitch.
//   encode4x4_generic_kernel  any pointer / pitch / size, and the pad region of CompressAndPad: byte loads with
//                             clamped coordinates straight from global memory.
#pragma once
#include <cuda.h>
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
template <int kCodec, typename Fetch>
__device__ __forceinline__ void encode_and_store(const uint32_t (&px)[16], Fetch fetch, bool one_pixel, int swap_rb,
                                                 int etc_strategy, const uint4 *alpha_table, uint8_t *out) {
  if constexpr (kCodec == kCodecDxt1) {
    const uint2 c = dxt1_encode_block(px, swap_rb != 0, false, fetch);
    *reinterpret_cast<uint2 *>(out) = c;
  } else if constexpr (kCodec == kCodecDxt5) {
    const uint2 a = dxt5_encode_alpha(px, one_pixel, alpha_table);
    const uint2 c = dxt1_encode_block(px, swap_rb != 0, true, fetch);
    *reinterpret_cast<uint4 *>(out) = make_uint4(a.x, a.y, c.x, c.y);
  } else {
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

// ---------------------------------------------------------------------------------------------------------
// TMA driver
// ---------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// mbarrier / TMA primitives; every address is a 32-bit shared-window address computed once per kernel.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Blocks until the barrier's phase with the given parity has completed.  The suspend-time hint lets the hardware
// park the warp instead of spinning through the issue slots the encoder warps need.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680)
      : "memory");
}
// 2-D tiled bulk tensor load, global -> shared, completion on an mbarrier; streaming (evict-first) L2 policy.
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap *map, uint32_t bar, int32_t x, int32_t y,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_dst),
      "l"(map), "r"(bar), "r"(x), "r"(y), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}

// Tile geometry.  A tile is kTileBlocksX x kTileBlocksY blocks; the TMA box is (kTileBlocksX*4*kNcomp/4) 32-bit
// words wide (the image row is described to TMA as 32-bit words so that RGB888 rows fit the 256-element box
// limit) and kTileBlocksY*4 rows tall.
template <int kNcomp>
struct TileShape {
  static constexpr int kBlocksX = 64;                       // 256 pixels
  static constexpr int kBlocksY = 4;                        // 16 pixel rows
  static constexpr int kRowWords = kBlocksX * kNcomp;       // 32-bit words per tile row (256 or 192)
  static constexpr int kRows = kBlocksY * 4;
  static constexpr int kBytes = kRowWords * 4 * kRows;      // 16384 or 12288
  static constexpr int kConsumerThreads = kBlocksX * kBlocksY;  // one block per consumer thread per tile
};

template <int kCodec, int kNcomp, int kTmaStages>
__global__ void __launch_bounds__(TileShape<kNcomp>::kConsumerThreads + 32)
    encode4x4_tma_kernel(const __grid_constant__ CUtensorMap src_map, const Encode4x4Params p, uint32_t tiles_x,
                         uint32_t num_tiles) {
  using Shape = TileShape<kNcomp>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // layout: kTmaStages tiles, then kTmaStages "full" barriers, then kTmaStages "empty" barriers
  // (kTmaStages trades bytes in flight per CTA against resident CTAs per SM; the launcher picks)
  const uint32_t tiles_s = smem_u32(smem_raw);
  const uint32_t full_s = tiles_s + kTmaStages * Shape::kBytes, empty_s = full_s + kTmaStages * 8;
  constexpr uint32_t kConsumerWarps = Shape::kConsumerThreads / 32;
  constexpr uint32_t kBlockBytes = CodecTraits<kCodec>::kBlockBytes;
  const uint32_t step_y = gridDim.x / tiles_x, step_x = gridDim.x - step_y * tiles_x;
  // DXT5 keeps its 8 KB crossing-point table behind the ring
  const uint4 *alpha_table = reinterpret_cast<const uint4 *>(smem_raw + kTmaStages * (Shape::kBytes + 16));
  if constexpr (kCodec == kCodecDxt5) {
    uint4 *dst = reinterpret_cast<uint4 *>(smem_raw + kTmaStages * (Shape::kBytes + 16));
    const uint4 *src = reinterpret_cast<const uint4 *>(g_dxt5_alpha_table);
    for (uint32_t i = threadIdx.x; i < kDxt5AlphaTableBytes / 16; i += blockDim.x) dst[i] = src[i];
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTmaStages; ++s) {
      mbar_init(full_s + 8 * s, 1);
      mbar_init(empty_s + 8 * s, kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t stage = 0, phase = 0;
  uint32_t ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;  // this CTA's current tile
  const uint32_t warp = threadIdx.x >> 5;
  if (warp == kConsumerWarps) {
    // ---- producer warp: one lane streams this CTA's tiles through the ring
    if ((threadIdx.x & 31) == 0) {
      const uint64_t policy = l2_evict_first_policy();
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(empty_s + 8 * stage, phase ^ 1u);
        mbar_arrive_expect_tx(full_s + 8 * stage, Shape::kBytes);
        tma_load_2d(tiles_s + stage * Shape::kBytes, &src_map, full_s + 8 * stage,
                    static_cast<int32_t>((p.col0 + tx * Shape::kBlocksX) * kNcomp),
                    static_cast<int32_t>((p.row0 + ty * Shape::kBlocksY) * 4u), policy);
        if (++stage == kTmaStages) {
          stage = 0;
          phase ^= 1u;
        }
        tx += step_x;  // next tile of this CTA: tile + gridDim.x, kept as (tx, ty) without dividing
        ty += step_y;
        if (tx >= tiles_x) {
          tx -= tiles_x;
          ++ty;
        }
      }
    }
    return;
  }

  // ---- consumer warps: thread t owns block (t / kBlocksX, t % kBlocksX) of every tile
  const uint32_t lbx = threadIdx.x % Shape::kBlocksX, lby = threadIdx.x / Shape::kBlocksX;
  const uint32_t own_off = (lby * 4u * Shape::kRowWords + lbx * kNcomp) * 4u;  // byte offset of the window in a tile
  for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const uint32_t tile_br = p.row0 + ty * Shape::kBlocksY, tile_bc = p.col0 + tx * Shape::kBlocksX;
    const uint8_t *tile_bytes = smem_raw + stage * Shape::kBytes;
    // Whole tile inside the image and inside this launch's block range: no clamping, no bounds checks.
    const bool tile_inside = (tile_br + Shape::kBlocksY) * 4u <= p.height && (tile_bc + Shape::kBlocksX) * 4u <= p.width &&
                             tile_br + Shape::kBlocksY <= p.row1 && tile_bc + Shape::kBlocksX <= p.col1;
    // Pixel i of this thread's window, re-read from the tile (the encoders look two pixels up by index).
    auto fetch_inside = [&](uint32_t i) {
      if constexpr (kNcomp == 4) {
        return *reinterpret_cast<const uint32_t *>(tile_bytes + own_off + (i >> 2) * (Shape::kRowWords * 4) + (i & 3u) * 4u);
      } else {
        const uint8_t *q = tile_bytes + own_off + (i >> 2) * (Shape::kRowWords * 4) + (i & 3u) * 3u;
        return static_cast<uint32_t>(q[0]) | (static_cast<uint32_t>(q[1]) << 8) | (static_cast<uint32_t>(q[2]) << 16);
      }
    };
    // Same with clamp-to-edge replication for tiles that the image edge cuts through.  The clamped coordinate
    // never leaves the tile because every encoded window starts inside the image.
    auto fetch_edge = [&](uint32_t i) {
      const uint32_t ymax = min(p.height - 1u - tile_br * 4u, static_cast<uint32_t>(Shape::kRows - 1));
      const uint32_t xmax = min(p.width - 1u - tile_bc * 4u, static_cast<uint32_t>(Shape::kBlocksX * 4 - 1));
      const uint32_t y = min(lby * 4u + (i >> 2), ymax), x = min(lbx * 4u + (i & 3u), xmax);
      if constexpr (kNcomp == 4) {
        return *reinterpret_cast<const uint32_t *>(tile_bytes + (y * Shape::kRowWords + x) * 4u);
      } else {
        const uint8_t *q = tile_bytes + y * (Shape::kRowWords * 4) + x * 3u;
        return static_cast<uint32_t>(q[0]) | (static_cast<uint32_t>(q[1]) << 8) | (static_cast<uint32_t>(q[2]) << 16);
      }
    };
    mbar_wait(full_s + 8 * stage, phase);

    uint8_t *out = p.dst + (static_cast<size_t>(tile_br + lby) * p.grid_cols + tile_bc + lbx) * kBlockBytes;
    if (tile_inside) {
      if constexpr (kCodec == kCodecDxt1 && kNcomp == 3) {
        // RGB888 -> DXT1 never needs the unpacked pixels: luminance keys come straight from the row words
        uint32_t rows[4][3];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const uint32_t *w = reinterpret_cast<const uint32_t *>(tile_bytes + own_off + y * (Shape::kRowWords * 4));
          rows[y][0] = w[0]; rows[y][1] = w[1]; rows[y][2] = w[2];
        }
        *reinterpret_cast<uint2 *>(out) = dxt1_encode_rgb888_rows(rows, p.swap_rb != 0, false, fetch_inside);
      } else {
        uint32_t px[16];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const uint8_t *row = tile_bytes + own_off + y * (Shape::kRowWords * 4);
          if constexpr (kNcomp == 4) {
            const uint4 v = *reinterpret_cast<const uint4 *>(row);
            px[4 * y + 0] = v.x; px[4 * y + 1] = v.y; px[4 * y + 2] = v.z; px[4 * y + 3] = v.w;
          } else {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(row);  // 12 bytes = four packed RGB pixels
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            px[4 * y + 0] = w0 & 0x00ffffffu;
            px[4 * y + 1] = __funnelshift_r(w0, w1, 24) & 0x00ffffffu;
            px[4 * y + 2] = __funnelshift_r(w1, w2, 16) & 0x00ffffffu;
            px[4 * y + 3] = w2 >> 8;
          }
        }
        encode_and_store<kCodec>(px, fetch_inside, false, p.swap_rb, p.etc_strategy, alpha_table, out);
      }
    } else if (tile_br + lby < p.row1 && tile_bc + lbx < p.col1) {
      uint32_t px[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) px[i] = fetch_edge(i);
      encode_and_store<kCodec>(px, fetch_edge, false, p.swap_rb, p.etc_strategy, alpha_table, out);
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(empty_s + 8 * stage);
    if (++stage == kTmaStages) {
      stage = 0;
      phase ^= 1u;
    }
    tx += step_x;
    ty += step_y;
    if (tx >= tiles_x) {
      tx -= tiles_x;
      ++ty;
    }
  }
}

}  // namespace icb
