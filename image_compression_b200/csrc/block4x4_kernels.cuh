// block4x4_kernels.cuh -- the TMA image drivers for the 4x4-block codecs (DXT1, DXT5, ETC1).
//
// These kernels ARE the reference's block loop, Compressor4x4Helper::Compress / CompressAndPad
// (/root/reference/image_compression/internal/compressor4x4_helper.h:175-216, 479-520): one 4x4 window per
// block in raster order.
//
// Three drivers:
//   encode4x4_ring_kernel     same tiles, ring and consumers as encode4x4_tma_kernel but without a producer warp: the
//                             consumer warps count themselves off a slot and the last one refills it (DXT5's default;
//                             see the comment above that kernel for the measured trade-off).
//   encode4x4_tma_kernel      the fast path, for every 64 x 4-block tile that lies entirely inside the image.
//                             Persistent CTAs; one elected producer thread streams 2-D pixel tiles HBM -> shared
//                             memory with TMA (cp.async.bulk.tensor) through a ring of mbarrier-guarded stages;
//                             consumer warps read their block's four rows with 128-bit (or 3x32-bit for RGB888)
//                             conflict-free shared loads, encode in registers and write 8/16 B per lane, coalesced.
//                             Needs a 16-byte aligned base and pitch.
//   encode4x4_generic_kernel  (block4x4_generic.cuh) any pointer / pitch / size: the ragged right and bottom strips
//                             the tiles do not cover, unaligned sources and the pad region of CompressAndPad.  Byte
//                             loads with clamped coordinates straight from global memory.
#pragma once
#include <cuda.h>
#include <cstdint>

#include "block4x4_generic.cuh"

namespace icb {

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
// The producer's wait for a free ring slot.  It is nearly always early (the ring is full while the encoders work).
// Neither the suspend-hinted wait above nor this nanosleep really parks the warp -- ncu's source view shows the loop
// coming round every ~7 ns whatever the sleep argument (50 ns .. 1 us measured the same; 5 us was slower) -- but
// this form measured fastest, and the polling warp mostly fills issue slots the encoders leave idle (see the ring
// kernel below for the variant without any polling).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(200);
  }
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
// L2 prefetch of a tile (no shared-memory destination, no completion to wait for).  The ring kernel issues it for a
// CTA's first tiles BEFORE griddepcontrol.wait: a prefetch only moves lines from DRAM into L2, which every later write of
// the preceding kernel goes through as well, so it cannot make stale data visible -- it merely lets the DRAM latency
// of the first loads overlap the tail of the kernel before (CTAs of this launch become resident as soon as CTAs of
// that one exit).  Measured on B200, 8192^2, launches back to back: DXT1 from RGBA8 48.9-49.1 -> 47.95 us, DXT5
// 76.0 -> 74.7 us.  The producer-warp kernel does not do it: there it COSTS (DXT1 from RGB888 41.1 -> 42.5 us, ETC1,
// which launches without programmatic dependent launch, 156 -> 166 us; profiles/r02b_driver_ab.txt, visit v8).
#ifndef ICB_RING_PREFETCH_TILES
#define ICB_RING_PREFETCH_TILES kTmaStages  // how many of a CTA's first tiles (A/B knob)
#endif
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap *map, int32_t x, int32_t y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}

// Tile geometry.  A tile is kBlocksX x kBlocksY blocks; the TMA box is (kBlocksX*4*kNcomp/4) 32-bit words wide (the
// image row is described to TMA as 32-bit words so that RGB888 rows fit the 256-element box limit) and kBlocksY*4
// rows tall.  One consumer thread per block of the tile.
// Tile height per codec, measured on B200 (profiles/r02_tile_shapes.txt, 8192^2, us per launch):
//   DXT1   64 x 8 blocks (32 KB tiles, 16 consumer warps, 2 CTAs per SM): 50.2 (RGBA8) / 49.6 (RGB888) against 51.6 / 51.7
//          with 64 x 4 and 52.2 / 56.6 with 64 x 2 -- fewer, larger hand-overs suit the kernel that is bound by them;
//   DXT5   64 x 4 (110.9 us with 64 x 8, 91.8 with 64 x 2 against 87.2);
//   ETC1   64 x 4, 3 CTAs of 288 threads and 72 registers per SM (end of round 2, 4096^2: 110.9 us; 64 x 2 / 64 x 1 on the
//          producer-less driver with the same 768 encoder threads per SM 111.1 / 111.9, with a producer warp 115.0 / 117.4;
//          2 CTAs per SM with 96 registers 121.1 against 118.0 for the kernel of that visit).
// ICB_DXT1_TILE_BLOCKS_Y overrides DXT1's for A/B builds (tools/build_variants.sh).
#ifndef ICB_DXT1_TILE_BLOCKS_Y
#define ICB_DXT1_TILE_BLOCKS_Y 8
#endif
template <int kCodec>
constexpr int tile_blocks_y() { return kCodec == kCodecDxt1 ? ICB_DXT1_TILE_BLOCKS_Y : 4; }

template <int kCodec, int kNcomp>
struct TileShape {
  static constexpr int kBlocksX = 64;                       // 256 pixels
  static constexpr int kBlocksY = tile_blocks_y<kCodec>();  // 16 or 32 pixel rows
  static constexpr int kRowWords = kBlocksX * kNcomp;       // 32-bit words per tile row (256 or 192)
  static constexpr int kRows = kBlocksY * 4;
  static constexpr int kBytes = kRowWords * 4 * kRows;      // 16384 / 12288 (x2 for DXT1)
  static constexpr int kConsumerThreads = kBlocksX * kBlocksY;  // one block per consumer thread per tile
  // resident CTAs per SM the kernels are compiled for (register budget = 65536 / threads / CTAs)
  static constexpr int kProducerMinCtas =
      kCodec == kCodecDxt1 ? (kBlocksY >= 16 ? 1 : (kBlocksY >= 8 ? 2 : 4)) : 3;
};

// Shared-memory loads by 32-bit shared-window address (no generic pointers: the tile base stays one register and the
// stage / row offsets fold into the instruction's immediate).
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint64_t lds_u64(uint32_t addr) {
  uint64_t v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u64(uint32_t addr, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// The TMA kernel only ever sees blocks whose windows lie entirely inside the image (the launcher hands ragged edge
// blocks and pad regions to the generic kernel), so it has no clamping and no bounds checks.  It covers block rows
// [row0,row1) x columns [col0,col1), each extent at least one tile; tiles are numbered row-major, tiles_x per row,
// and the last tile of a row / column is shifted back so that it ends exactly at col1 / row1 -- the blocks it
// shares with its neighbour are simply encoded twice, to the same bytes.
// kSwapRb: red and blue exchanged (kBGR / kBGRA sources) as a COMPILE-TIME constant -- the luminance, quantiser and
// channel-extraction weights of the DXT encoders depend on it, and as a run-time parameter it cost about ten
// uniform-datapath selects per block, which share the warp's issue slots (ETC1 does not look at it).
template <int kCodec, int kNcomp, int kTmaStages, bool kSwapRb>
__global__ void __launch_bounds__(TileShape<kCodec, kNcomp>::kConsumerThreads + 32, TileShape<kCodec, kNcomp>::kProducerMinCtas)
    encode4x4_tma_kernel(const __grid_constant__ CUtensorMap src_map, const Encode4x4Params p, uint32_t tiles_x,
                         uint32_t num_tiles) {
  using Shape = TileShape<kCodec, kNcomp>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // layout: kTmaStages tiles, then kTmaStages "full" barriers, kTmaStages "empty" barriers, kTmaStages tile origins
  // (kTmaStages trades bytes in flight per CTA against resident CTAs per SM; the launcher picks)
  // (volatile: keeps the shared-window base in a register instead of re-deriving it from SR_CgaCtaId per tile)
  uint32_t tiles_s;
  asm volatile("mov.u32 %0, %1;" : "=r"(tiles_s) : "r"(smem_u32(smem_raw)));
  const uint32_t full_s = tiles_s + kTmaStages * Shape::kBytes, empty_s = full_s + kTmaStages * 8;
  // ... then per stage the byte offset of the tile's first block in the output (8 bytes): whoever loads a tile works
  // its position out ONCE and leaves it here; the consumers pick it up with one 64-bit shared load instead of each
  // thread tracking (tx, ty), clamping and multiplying on its own (about fifteen instructions per block).
  const uint32_t origin_s = empty_s + kTmaStages * 8;
  constexpr uint32_t kConsumerWarps = Shape::kConsumerThreads / 32;
  constexpr uint32_t kBlockBytes = CodecTraits<kCodec>::kBlockBytes;
  constexpr uint32_t kRowBytes = Shape::kRowWords * 4;
  const uint32_t step_y = gridDim.x / tiles_x, step_x = gridDim.x - step_y * tiles_x;
  // DXT5's 32 KB crossing table is read from global memory (four 16-byte loads per block, L1-resident)
  const uint4 *alpha_table = reinterpret_cast<const uint4 *>(g_dxt5_alpha_table);

  uint32_t ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;  // this CTA's current tile
  const uint32_t warp = threadIdx.x >> 5;
  // ---- producer: one lane of the last warp streams this CTA's tiles through the ring.  It also sets the barriers
  // up and fills the ring BEFORE the CTA-wide sync, so the first tiles are already in flight while the other warps
  // are still starting.
  const bool is_producer = threadIdx.x == Shape::kConsumerThreads;
  uint32_t p_stage = 0, p_phase = 0, p_tile = blockIdx.x;
  const uint64_t policy = is_producer ? l2_evict_first_policy() : 0ull;  // the source is read exactly once
  auto produce_one = [&]() {
    mbar_wait_relaxed(empty_s + 8 * p_stage, p_phase ^ 1u);  // passes at once on the first trip round the ring
    const uint32_t bc = min(p.col0 + tx * Shape::kBlocksX, p.col1 - Shape::kBlocksX);
    const uint32_t br = min(p.row0 + ty * Shape::kBlocksY, p.row1 - Shape::kBlocksY);
    // (ordered before the consumers' reads by the arrive below and the tile's completion on the same barrier)
    sts_u64(origin_s + 8 * p_stage, (static_cast<uint64_t>(br) * p.grid_cols + bc) * kBlockBytes);
    mbar_arrive_expect_tx(full_s + 8 * p_stage, Shape::kBytes);
    tma_load_2d(tiles_s + p_stage * Shape::kBytes, &src_map, full_s + 8 * p_stage, static_cast<int32_t>(bc * kNcomp),
                static_cast<int32_t>(br * 4u), policy);
    if (++p_stage == kTmaStages) {
      p_stage = 0;
      p_phase ^= 1u;
    }
    p_tile += gridDim.x;
    tx += step_x;  // next tile of this CTA: tile + gridDim.x, kept as (tx, ty) without dividing
    ty += step_y;
    if (tx >= tiles_x) {
      tx -= tiles_x;
      ++ty;
    }
  };
  // Programmatic dependent launch: let the next kernel in the stream start its own launch and prologue while this
  // one is still running, and do not touch global memory before the previous kernel has completed and flushed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (is_producer) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&src_map) : "memory");
    for (int s = 0; s < kTmaStages; ++s) {
      mbar_init(full_s + 8 * s, 1);
      mbar_init(empty_s + 8 * s, kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // Only this thread waits for the previous kernel: every other thread's first global access (a block store)
    // comes after it has consumed a tile, i.e. after this thread has passed the wait and issued the load.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int s = 0; s < kTmaStages && p_tile < num_tiles; ++s) produce_one();
  }
  __syncthreads();
  if (warp == kConsumerWarps) {
    if (is_producer)
      while (p_tile < num_tiles) produce_one();
    return;
  }

  // ---- consumer warps: thread t owns block (t / kBlocksX, t % kBlocksX) of every tile
  const uint32_t lbx = threadIdx.x % Shape::kBlocksX, lby = threadIdx.x / Shape::kBlocksX;
  // shared-window address of this thread's window in stage 0
  const uint32_t win0 = tiles_s + (lby * 4u * Shape::kRowWords + lbx * kNcomp) * 4u;
  // This thread's block in a tile whose first block is (0,0); the launcher checks that a row of blocks fits 32 bits.
  uint8_t *const out_origin = p.dst + (static_cast<size_t>(lby) * p.grid_cols + lbx) * kBlockBytes;
  constexpr bool swap_rb = kSwapRb;
  // The ring is walked with the stage as a compile-time constant (barrier and tile addresses become immediates)
  // where the encoder is small; ETC1's exhaustive search is ~5000 instructions, three copies of which would not
  // fit the instruction cache, so it keeps a run-time stage.
  constexpr int kUnroll = kCodec == kCodecEtc1 ? 1 : kTmaStages;
  uint32_t phase = 0, tile = blockIdx.x, rt_stage = 0;
  while (true) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t stage = kUnroll == 1 ? rt_stage : static_cast<uint32_t>(u);
      if (tile >= num_tiles) return;  // uniform over the CTA's consumers; nothing after the loop needs them
      const uint32_t win = win0 + stage * Shape::kBytes;
      // Pixel i (< 16) of this thread's window, re-read from the tile (the encoders look two pixels up by index).
      auto fetch = [&](uint32_t i) {
        if constexpr (kNcomp == 4) {
          // (i >> 2) rows of 1024 bytes + (i & 3) pixels of 4 bytes, as one multiply and one mask
          static_assert(kRowBytes == 1024, "offset trick assumes 1 KB tile rows");
          return lds_u32(win + ((i * 0x104u) & 0xc0cu));
        } else {
          const uint32_t q = win + (i >> 2) * kRowBytes + (i & 3u) * 3u;
          return lds_u8(q) | (lds_u8(q + 1) << 8) | (lds_u8(q + 2) << 16);
        }
      };
      mbar_wait(full_s + 8 * stage, phase);
      uint8_t *out = out_origin + lds_u64(origin_s + 8 * stage);

      if constexpr (kCodec == kCodecDxt1 && kNcomp == 3) {
        // RGB888 -> DXT1 never needs the unpacked pixels: luminance keys come straight from the row words
        uint32_t rows[4][3];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          rows[y][0] = lds_u32(win + y * kRowBytes);
          rows[y][1] = lds_u32(win + y * kRowBytes + 4);
          rows[y][2] = lds_u32(win + y * kRowBytes + 8);
        }
        *reinterpret_cast<uint2 *>(out) = dxt1_encode_rgb888_rows<true>(rows, swap_rb, false, fetch);
      } else {
        uint32_t px[16];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          if constexpr (kNcomp == 4) {
            const uint4 v = lds_v4(win + y * kRowBytes);
            px[4 * y + 0] = v.x; px[4 * y + 1] = v.y; px[4 * y + 2] = v.z; px[4 * y + 3] = v.w;
          } else {  // 12 bytes = four packed RGB pixels
            const uint32_t w0 = lds_u32(win + y * kRowBytes), w1 = lds_u32(win + y * kRowBytes + 4),
                           w2 = lds_u32(win + y * kRowBytes + 8);
            px[4 * y + 0] = w0 & 0x00ffffffu;
            px[4 * y + 1] = __funnelshift_r(w0, w1, 24) & 0x00ffffffu;
            px[4 * y + 2] = __funnelshift_r(w1, w2, 16) & 0x00ffffffu;
            px[4 * y + 3] = w2 >> 8;
          }
        }
        encode_and_store<kCodec, true>(px, fetch, false, swap_rb ? 1 : 0, p.etc_strategy, alpha_table, out);
      }
      // Hands the slot back to the producer.  (Handing it back from inside the encoder, right after its last read of the
      // staged pixels, measured the same: 43.6 against 43.0-44.8 us for RGB888, 51.6 against 50.9 for RGBA8.)
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(empty_s + 8 * stage);
      tile += gridDim.x;
      if constexpr (kUnroll == 1) {
        if (++rt_stage == kTmaStages) {
          rt_stage = 0;
          phase ^= 1u;
        }
      }
    }
    if constexpr (kUnroll != 1) phase ^= 1u;
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMA driver without a producer warp
// ---------------------------------------------------------------------------------------------------------
//
// Same tiles, ring and consumer code as encode4x4_tma_kernel, but nobody polls for free ring slots: every warp
// counts itself off a per-stage "released" mbarrier as soon as it has read its last pixel from a tile (one arrive per
// warp per tile, issued from inside the encoder through the `release` hook), and the warp whose arrival completes the
// barrier's phase refills the slot on the spot -- expect_tx + one TMA load of tile (current + stages * grid).  Dropping
// the producer warp frees its register share (256 instead of 288 threads per CTA: 64 registers per thread at four CTAs
// per SM), which is what DXT5 needs: with the producer warp it is limited to three CTAs per SM (72 registers), here it
// runs four, with a two-stage ring (early release makes two stages enough) and its crossing table read through L1.
// Measured on B200 (8192^2): DXT5 103.5 -> 94.2 us in round 1.  DXT1 from RGBA8 moved here in round 2, once the
// integer-lane index search had shortened its encoder (48.8-49.1 us against 50.9-51.2 with the producer warp; ncu times
// the kernel alone at 46.6 us, the rest is the gap between back-to-back launches); DXT1 from RGB888 (43.0-44.8 against
// 46.6) and ETC1 stay with the producer warp (ICB_DRIVER=ring|producer overrides for experiments).
// Until the second half of round 2 the warps counted off with an acq_rel atomic add on a counter word, which ptxas
// expands into its warp-aggregated form (vote, FLO, POPC, MEMBAR, ATOMS, SHFL: eighteen instructions in lane 0's path);
// the mbarrier form is six (DXT5 78.3 -> 77.6 us).

// Arrives on an mbarrier and returns the number of arrivals its phase was still waiting for BEFORE this one
// (mbarrier.pending_count of the state the arrive returns; tools/microbench/mbar_pending.cu prints 4 3 2 1 4 3 2 1 4 for
// a barrier of four): 1 = this arrival completed the phase.
__device__ __forceinline__ uint32_t mbar_arrive_pending(uint32_t bar) {
  uint32_t pending;
  asm volatile(
      "{\n"
      ".reg .b64 st;\n"
      "mbarrier.arrive.shared::cta.b64 st, [%1];\n"
      "mbarrier.pending_count.b64 %0, st;\n"
      "}\n"
      : "=r"(pending)
      : "r"(bar)
      : "memory");
  return pending;
}

// Resident CTAs per SM the DXT5 build of the ring kernel is compiled for.  4 = 64 registers (the measured default).
// 5 = 48 registers: no faster (79.5 against 79.0 us) and, with CUDA 12.9's ptxas, WRONG -- in that build the first
// statistics key of the alpha half, __vadd2(x0, 0xffffffff), comes out as VIADD.16x2 R8,R4,0x0 / VIADD.16x2 R5,R4,
// 0xffffffff / PRMT R14,R8,0x7610,R5, i.e. the low lane keeps alpha instead of alpha - 1; the bench's in-run parity
// flag caught it (profiles/r02b_driver_ab.txt).  Kept as a knob only so that the finding can be reproduced.
#ifndef ICB_DXT5_RING_MIN_CTAS
#define ICB_DXT5_RING_MIN_CTAS 4
#endif

template <int kCodec, int kNcomp, int kTmaStages, bool kSwapRb>
__global__ void __launch_bounds__(TileShape<kCodec, kNcomp>::kConsumerThreads,
                                  kCodec == kCodecEtc1 ? 3 : (kCodec == kCodecDxt5 ? ICB_DXT5_RING_MIN_CTAS : TileShape<kCodec, kNcomp>::kProducerMinCtas))
    encode4x4_ring_kernel(const __grid_constant__ CUtensorMap src_map, const Encode4x4Params p, uint32_t tiles_x,
                          uint32_t num_tiles) {
  using Shape = TileShape<kCodec, kNcomp>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // layout: kTmaStages tiles, then per stage a "full" barrier (8 bytes), a "released" barrier of kWarps arrivals
  // (8 bytes), a tile origin (8 bytes)
  uint32_t tiles_s;
  asm volatile("mov.u32 %0, %1;" : "=r"(tiles_s) : "r"(smem_u32(smem_raw)));
  // ... then per stage the byte offset of the tile's first block in the output (see encode4x4_tma_kernel)
  const uint32_t full_s = tiles_s + kTmaStages * Shape::kBytes, count_s = full_s + kTmaStages * 8;
  const uint32_t origin_s = count_s + kTmaStages * 8;
  constexpr uint32_t kWarps = Shape::kConsumerThreads / 32;
  constexpr uint32_t kBlockBytes = CodecTraits<kCodec>::kBlockBytes;
  constexpr uint32_t kRowBytes = Shape::kRowWords * 4;
  // DXT5's 32 KB crossing table is read from global memory (four 16-byte loads per block, L1-resident): a
  // shared-memory copy would cost resident CTAs.
  const uint4 *alpha_table = reinterpret_cast<const uint4 *>(g_dxt5_alpha_table);
  const uint32_t last_bc = p.col1 - Shape::kBlocksX, last_br = p.row1 - Shape::kBlocksY;

  // Loads tile number `t` into ring slot `stage` (one thread).  Tiles are numbered row-major, tiles_x per row; the last
  // tile of a row / column is shifted back so that it ends exactly at col1 / row1.
  auto load_tile = [&](uint32_t t, uint32_t stage) {
    const uint32_t ty = t / tiles_x, tx = t - ty * tiles_x;
    const uint32_t bc = min(p.col0 + tx * Shape::kBlocksX, last_bc), br = min(p.row0 + ty * Shape::kBlocksY, last_br);
    // (every warp reads the slot's previous origin before it counts itself off, i.e. before the refill that calls this)
    sts_u64(origin_s + 8 * stage, (static_cast<uint64_t>(br) * p.grid_cols + bc) * kBlockBytes);
    mbar_arrive_expect_tx(full_s + 8 * stage, Shape::kBytes);
    tma_load_2d(tiles_s + stage * Shape::kBytes, &src_map, full_s + 8 * stage, static_cast<int32_t>(bc * kNcomp),
                static_cast<int32_t>(br * 4u), l2_evict_first_policy());  // the source is read exactly once
  };

  // Programmatic dependent launch: the next kernel in the stream may start its prologue now; this one touches global
  // memory only after the previous kernel has completed (thread 0 waits before the first load, and every other
  // thread's first global access is a store that follows a tile thread 0 loaded).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&src_map) : "memory");
    for (int s = 0; s < kTmaStages; ++s) {
      mbar_init(full_s + 8 * s, 1);
      mbar_init(count_s + 8 * s, kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // the first tiles into L2 while the previous kernel is still finishing (see tma_prefetch_l2_2d)
    for (uint32_t s = 0, t = blockIdx.x; s < ICB_RING_PREFETCH_TILES && t < num_tiles; ++s, t += gridDim.x) {
      const uint32_t pty = t / tiles_x, ptx = t - pty * tiles_x;
      tma_prefetch_l2_2d(&src_map, static_cast<int32_t>(min(p.col0 + ptx * Shape::kBlocksX, last_bc) * kNcomp),
                         static_cast<int32_t>(min(p.row0 + pty * Shape::kBlocksY, last_br) * 4u));
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (uint32_t s = 0, t = blockIdx.x; s < kTmaStages && t < num_tiles; ++s, t += gridDim.x) load_tile(t, s);
  }
  __syncthreads();

  // thread t owns block (t / kBlocksX, t % kBlocksX) of every tile
  const uint32_t lbx = threadIdx.x % Shape::kBlocksX, lby = threadIdx.x / Shape::kBlocksX;
  const uint32_t win0 = tiles_s + (lby * 4u * Shape::kRowWords + lbx * kNcomp) * 4u;
  uint8_t *const out_origin = p.dst + (static_cast<size_t>(lby) * p.grid_cols + lbx) * kBlockBytes;
  constexpr bool swap_rb = kSwapRb;
  const uint32_t refill_stride = kTmaStages * gridDim.x;
  constexpr int kUnroll = kCodec == kCodecEtc1 ? 1 : kTmaStages;  // see encode4x4_tma_kernel
  uint32_t phase = 0, tile = blockIdx.x, rt_stage = 0;
  while (true) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t stage = kUnroll == 1 ? rt_stage : static_cast<uint32_t>(u);
      if (tile >= num_tiles) return;
      const uint32_t win = win0 + stage * Shape::kBytes;
      auto fetch = [&](uint32_t i) {
        if constexpr (kNcomp == 4) {
          static_assert(kRowBytes == 1024 || kNcomp != 4, "offset trick assumes 1 KB tile rows");
          return lds_u32(win + ((i * 0x104u) & 0xc0cu));
        } else {
          const uint32_t q = win + (i >> 2) * kRowBytes + (i & 3u) * 3u;
          return lds_u8(q) | (lds_u8(q + 1) << 8) | (lds_u8(q + 2) << 16);
        }
      };
      // Done with this slot (called by the encoder as soon as it has read its last pixel from the tile): count this
      // warp off; the last warp to do so refills the slot.
      auto release = [&]() {
#ifdef ICB_RING_SKEW_TEST
        // Test builds only (tools/build_variants.sh skew "-DICB_RING_SKEW_TEST"): some warps hand their slot back
        // microseconds late, a different set on every tile; the output must not change.
        if ((((threadIdx.x >> 5) * 5u + tile) & 3u) == 0u) __nanosleep(3000);
#endif
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
          // The arrive releases this warp's reads of the slot; the warp that completes the phase acquires everybody's
          // (a wait on the phase it has just completed: passes at once) before it lets TMA overwrite the slot.
          if (mbar_arrive_pending(count_s + 8 * stage) == 1u && tile + refill_stride < num_tiles) {
            mbar_wait(count_s + 8 * stage, phase);
            load_tile(tile + refill_stride, stage);
          }
        }
        __syncwarp();
      };
#ifdef ICB_RING_SKEW_TEST
      if ((((threadIdx.x >> 5) * 3u + tile) & 7u) == 1u) __nanosleep(5000);  // ... and some arrive late at the next tile
#endif
      mbar_wait(full_s + 8 * stage, phase);
      uint8_t *out = out_origin + lds_u64(origin_s + 8 * stage);  // before release(): the refill overwrites it

      if constexpr (kCodec == kCodecDxt1 && kNcomp == 3) {
        uint32_t rows[4][3];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          rows[y][0] = lds_u32(win + y * kRowBytes);
          rows[y][1] = lds_u32(win + y * kRowBytes + 4);
          rows[y][2] = lds_u32(win + y * kRowBytes + 8);
        }
        *reinterpret_cast<uint2 *>(out) = dxt1_encode_rgb888_rows<true>(rows, swap_rb, false, fetch, release);
      } else {
        uint32_t px[16];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          if constexpr (kNcomp == 4) {
            const uint4 v = lds_v4(win + y * kRowBytes);
            px[4 * y + 0] = v.x; px[4 * y + 1] = v.y; px[4 * y + 2] = v.z; px[4 * y + 3] = v.w;
          } else {
            const uint32_t w0 = lds_u32(win + y * kRowBytes), w1 = lds_u32(win + y * kRowBytes + 4),
                           w2 = lds_u32(win + y * kRowBytes + 8);
            px[4 * y + 0] = w0 & 0x00ffffffu;
            px[4 * y + 1] = __funnelshift_r(w0, w1, 24) & 0x00ffffffu;
            px[4 * y + 2] = __funnelshift_r(w1, w2, 16) & 0x00ffffffu;
            px[4 * y + 3] = w2 >> 8;
          }
        }
        encode_and_store<kCodec, true>(px, fetch, false, swap_rb ? 1 : 0, p.etc_strategy, alpha_table, out, release);
      }
      tile += gridDim.x;
      if constexpr (kUnroll == 1) {
        if (++rt_stage == kTmaStages) {
          rt_stage = 0;
          phase ^= 1u;
        }
      }
    }
    if constexpr (kUnroll != 1) phase ^= 1u;
  }
}


}  // namespace icb
