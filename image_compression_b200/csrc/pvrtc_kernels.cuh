// pvrtc_kernels.cuh -- image-level PVRTC1 2bpp pipeline (reference: CompressPVRTC_RGBA_2BPP,
// /root/reference/image_compression/internal/pvrtc_compressor.cc:586-597 = Morph :506-521, Modulate :527-540,
// Encode :551-580).
//
// Three kernels, all coalesced and small enough to stay in the instruction cache (a two-kernel form with Modulate and
// Pack fused through shared memory is at the end of this file: bit-exact, but slower on B200, so opt-in):
//   pvrtc_morph_kernel     one thread per 8x4 block -> bit-reduced A and B colours (one (A, B) pair per block: the
//                          reference's two w/8 x h/4 images, interleaved, in scratch)
//   pvrtc_modulate_kernel  one thread per 8-pixel segment of kModRows consecutive rows: bilinear upscale of A and B
//                          (toroidal), modulation choice per pixel, eight 2-bit values packed into one uint16 of
//                          scratch.  The reference's byte-per-pixel modulation image becomes 2 bits per pixel and
//                          stays in L2.
//   pvrtc_pack_kernel      one thread per block: its 4 row words plus the wrapped column to the right and row
//                          below (all the mode decision needs), mode + bit packing, store at the Z-order slot.
#pragma once
#include <cstdint>

#include "pvrtc_encode.cuh"

namespace icb {

// Pixel rows per Modulate thread: kModRows consecutive rows that interpolate between the same two rows of the
// low-resolution images (rows 4g+2 .. 4g+5 lie between block rows g and g+1), so the six colour-pair loads and their
// lane splits are done once per thread.  Must divide 4.
#ifndef ICB_PVRTC_MOD_ROWS
#define ICB_PVRTC_MOD_ROWS 4
#endif
constexpr uint32_t kModRows = ICB_PVRTC_MOD_ROWS;
static_assert(kModRows == 1 || kModRows == 2 || kModRows == 4, "a thread's rows must share their low-resolution rows");
constexpr uint32_t kModThreads = 128;

// Whole image: src holds all rows, src_row0 = 0, every range covers the image.  Row stripe (one rank of a sharded
// image): src holds only the pixel rows of block rows [r0 - 1, r1 + 1) -- the stripe plus one
// block row of halo above and below, wrapped round the torus -- starting at image row src_row0 = 4 * (r0 - 1) mod h;
// Morph runs over those r1 - r0 + 2 block rows (the halo blocks are recomputed, not exchanged), Modulate over pixel
// rows [4 r0, 4 r1] (the extra row is the one the last block row's mode decision looks at) rounded outwards to whole
// thread units, Pack over [r0, r1).
// Scratch images are indexed by absolute position in both cases, and so is the Z-ordered output.
struct PvrtcParams {
  const uint32_t *src;          // RGBA8 pixels, row-major, no padding; row 0 of this buffer is image row src_row0
  const uint32_t *first_pixel;  // image pixel (0,0): quirk P1 (pvrtc_compressor.cc:255-329) needs it in every block
  uint2 *low;                   // (w/8) x (h/4) colour pairs: .x = the block's A colour, .y = its B colour (one 8-byte
                                // load/store per block; the two low-resolution images of the reference, interleaved)
  uint16_t *mod;                // h x (w/8) words: 2-bit modulation of pixels 8*bx .. 8*bx+7 of row y
  uint2 *dst;                   // w*h/32 blocks in Z-order
  uint32_t width, height;
  uint32_t src_row0;            // image row held in row 0 of src
  uint32_t morph_row0, morph_rows;  // block rows Morph covers: morph_row0 .. +morph_rows (mod h/4)
  uint32_t mod_unit0, mod_units;    // Modulate covers pixel rows 2 + kModRows * [mod_unit0, mod_unit0 + mod_units) (mod h)
  uint32_t pack_row0, pack_rows;    // block rows Pack covers
  uint32_t key_scale;               // the constant 32, kept out of the compiler's sight (pvrtc_encode.cuh:pv_key)
  uint32_t lw_shift;                // log2(width / 8): images are powers of two, so thread -> (block column, row) is
                                    // a mask and a shift instead of a 20-instruction division by a run-time value
};

// Host side: the parameter block for block rows [r0, r1) of an h x w image.  `whole`: src holds the whole image
// (src_row0 = 0) and every kernel covers everything; otherwise src holds the stripe and its two halo block rows.
// scratch: icb_pvrtc2_scratch_size(h, w) bytes (8-byte aligned) -- the A/B colour pairs, then the 2-bit modulation words.
inline PvrtcParams pvrtc_make_params(const void *src, const void *first_pixel, void *scratch, void *dst, uint32_t h,
                                     uint32_t w, uint32_t src_row0, uint32_t r0, uint32_t r1, bool whole) {
  PvrtcParams p;
  const uint32_t lw = w / 8, lh = h / 4, nblocks = lw * lh;
  p.src = static_cast<const uint32_t *>(src);
  p.first_pixel = static_cast<const uint32_t *>(first_pixel);
  p.low = static_cast<uint2 *>(scratch);
  p.mod = reinterpret_cast<uint16_t *>(p.low + nblocks);
  p.dst = static_cast<uint2 *>(dst);
  p.width = w;
  p.height = h;
  p.src_row0 = src_row0;
  if (whole) {
    p.morph_row0 = 0; p.morph_rows = lh;
    p.mod_unit0 = 0; p.mod_units = h / kModRows;
  } else {
    p.morph_row0 = (r0 + lh - 1) & (lh - 1); p.morph_rows = r1 - r0 + 2;
    // pixel rows 4 r0 .. 4 r1 inclusive, in units of kModRows rows that start at row 2 (the first unit may begin up to
    // two rows early and the last one end a row late: those rows lie in the resident halo block rows)
    const uint32_t units = h / kModRows;
    const uint32_t first = ((4 * r0 + h - 2) & (h - 1)) / kModRows, last = ((4 * r1 + h - 2) & (h - 1)) / kModRows;
    p.mod_unit0 = first; p.mod_units = ((last + units - first) & (units - 1)) + 1;
  }
  p.pack_row0 = r0; p.pack_rows = r1 - r0;
  p.key_scale = 32;
  p.lw_shift = 0;
  while ((1u << p.lw_shift) < lw) ++p.lw_shift;
  return p;
}

// Address of image row y (absolute, < height) in the resident rows.
__device__ __forceinline__ const uint32_t *pv_src_row(const PvrtcParams &p, uint32_t y) {
  return p.src + static_cast<size_t>((y - p.src_row0) & (p.height - 1u)) * p.width;
}

// All three kernels are launched with programmatic dependent launch: each lets the next kernel in the stream start
// its launch and prologue at once (launch_dependents) and waits for the kernel before it to be complete, its memory
// flushed, before its own first global access (griddepcontrol.wait) -- only the launch latency overlaps, which is what
// separates three 10-35 us kernels.  Morph waits too: it may follow the kernel that produced the image, and it
// overwrites the A/B colours the previous encode's Pack is still reading.  Because a programmatically launched kernel
// is alive before its predecessor has finished writing, every global load in these kernels is a plain load: the
// non-coherent path (__ldg) is defined only for data that stays read-only for the kernel's whole lifetime, and both the
// image (written by whatever precedes Morph in the stream) and the scratch images (written by the kernel before) are not.
#ifdef ICB_HOST_EMULATION  // tests/hostemu only: kernels run one after another on the CPU, nothing to order
__device__ __forceinline__ void pv_launch_dependents() {}
__device__ __forceinline__ void pv_wait_for_previous() {}
#else
__device__ __forceinline__ void pv_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pv_wait_for_previous() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// Eight CTAs per SM = 64 registers per thread: with that budget ptxas issues all eight 128-bit pixel loads before the
// first use (63 registers); left to itself it settles on 32 registers and loads two rows at a time, and the kernel
// waits on its loads (measured: 49.8 -> 48.7 us per encode).
#ifndef ICB_PVRTC_MORPH_MIN_CTAS
#define ICB_PVRTC_MORPH_MIN_CTAS 8
#endif
__global__ void __launch_bounds__(128, ICB_PVRTC_MORPH_MIN_CTAS) pvrtc_morph_kernel(const PvrtcParams p) {
  pv_launch_dependents();
  pv_wait_for_previous();  // whatever wrote the image; the previous encode's Pack (reads the colours written below)
  const uint32_t lw = p.width >> 3, lh = p.height >> 2;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= lw * p.morph_rows) return;
  const uint32_t bx = t & (lw - 1u), by = (p.morph_row0 + (t >> p.lw_shift)) & (lh - 1u);
  const uint32_t *origin = pv_src_row(p, by * 4u) + bx * 8;  // a block's four rows are contiguous in the buffer
  uint32_t px[32];
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    const uint4 *row = reinterpret_cast<const uint4 *>(origin + static_cast<size_t>(y) * p.width);
    const uint4 u = row[0], v = row[1];
    px[8 * y + 0] = u.x; px[8 * y + 1] = u.y; px[8 * y + 2] = u.z; px[8 * y + 3] = u.w;
    px[8 * y + 4] = v.x; px[8 * y + 5] = v.y; px[8 * y + 6] = v.z; px[8 * y + 7] = v.w;
  }
  // (32-bit offset: a block's four rows span less than 4 * width pixels, and one IMAD.WIDE then forms the address)
  auto fetch = [&](uint32_t j) { return origin[(j >> 3) * p.width + (j & 7u)]; };
  uint32_t ca, cb;
  pv_block_extremes(px, *p.first_pixel, p.key_scale, fetch, &ca, &cb);
  p.low[by * lw + bx] = make_uint2(ca, cb);
}

// Thread t handles pixels 8*bx .. 8*bx+7 of rows y0 .. y0+kModRows-1, y0 = 2 + kModRows * (mod_unit0 + t / lw) mod h.
// Resident CTAs per SM to compile for, as a build macro for A/B runs (tools/gpu_next_round.sh).  0 = no minimum: ptxas
// takes 80 registers (6 CTAs = 24 warps per SM); 7 gives 72 registers without spills, 8 gives 64 with 32 bytes of
// spill.  Unmeasured so far.
#ifndef ICB_PVRTC_MOD_MIN_CTAS
#define ICB_PVRTC_MOD_MIN_CTAS 0
#endif
#if ICB_PVRTC_MOD_MIN_CTAS > 0
#define ICB_PVRTC_MOD_BOUNDS __launch_bounds__(kModThreads, ICB_PVRTC_MOD_MIN_CTAS)
#else
#define ICB_PVRTC_MOD_BOUNDS __launch_bounds__(kModThreads)
#endif
// One Modulate unit: pixels 8*bx .. 8*bx+7 of rows y0 .. y0+kModRows-1 (y0 = 4g+2, so all of them interpolate between
// low-resolution rows g and g+1).  `wanted(r)` says whether row y0+r is needed at all; `store(r, y, bits)` receives the
// eight 2-bit modulation values of row y = y0+r (mod height).
template <typename Wanted, typename Store>
__device__ __forceinline__ void pv_modulate_unit(const PvrtcParams &p, uint32_t bx, uint32_t y0, Wanted wanted, Store store) {
  const uint32_t lw = p.width >> 3, lh = p.height >> 2;
  // Low-resolution rows/columns these pixel rows interpolate between, wrapped (pvrtc_compressor.cc:216-223).
  const uint32_t top = ((y0 - 2u) & (p.height - 1u)) >> 2, bottom = (top + 1u) & (lh - 1u);
  const uint32_t col[3] = {(bx + lw - 1u) & (lw - 1u), bx, (bx + 1u) & (lw - 1u)};
  // A and B colours of the top / bottom low-resolution row, three columns, as 16-bit lane pairs: index [c][k] =
  // column c, lane pair k (0 = A's (r,b), 1 = A's (g,a), 2 = B's (r,b), 3 = B's (g,a)).
  uint32_t top_c[3][4], bot_c[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const uint2 ct = p.low[top * lw + col[i]], cb = p.low[bottom * lw + col[i]];
    const PvLanes at = pv_split(ct.x), bt = pv_split(ct.y), ab = pv_split(cb.x), bb = pv_split(cb.y);
    top_c[i][0] = at.rb; top_c[i][1] = at.ga; top_c[i][2] = bt.rb; top_c[i][3] = bt.ga;
    bot_c[i][0] = ab.rb; bot_c[i][1] = ab.ga; bot_c[i][2] = bb.rb; bot_c[i][3] = bb.ga;
  }
  // Horizontal steps between neighbouring columns, as differences of the packed words: a word is the exact integer
  // lane0 + 65536 * lane1, so the difference of two words is the exact integer of the lane differences, negative lanes
  // and all, and everything below is linear in it modulo 2^32.
  uint32_t top_d[2][4], bot_d[2][4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      top_d[h][k] = top_c[h + 1][k] - top_c[h][k];
      bot_d[h][k] = bot_c[h + 1][k] - bot_c[h][k];
    }
#pragma unroll
  for (uint32_t r = 0; r < kModRows; ++r) {
    if (!wanted(r)) continue;
    const uint32_t y = (y0 + r) & (p.height - 1u);  // (only the unit that straddles the bottom edge wraps)
    // (y + 2) & 3; with four rows per unit (y0 = 4g + 2) that is r itself, a compile-time constant of the unrolled loop:
    // the vertical weights become immediates and the bottom row drops out of the unit's first pixel row
    const uint32_t fy = kModRows == 4 ? r : ((y + 2u) & 3u);
    // The bilinear blend ((4-fy)(8-fx) c00 + (4-fy) fx c01 + fy (8-fx) c10 + fy fx c11) / 32 of a channel, scaled by 8 so
    // that the divisor is 256 (every lane stays below 2^16: 255 * 4 * 64), is LINEAR in fx within a half segment:
    //   word(fx) = 64 * V_left + fx * 8 * (V_right - V_left),   V = (4-fy) * top + fy * bottom  (vertical blend)
    // so one multiply-add per lane pair and pixel walks along the row (two before: both columns weighted afresh for
    // every pixel), after a per-row set-up of a start word and a step word per half and lane pair.  Pixels 0..3 have
    // fx = 4..7 between columns (bx-1, bx), pixels 4..7 have fx = 0..3 between (bx, bx+1).
    uint32_t start[2][4], step[2][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int h = 0; h < 2; ++h) step[h][k] = top_d[h][k] * (8u * (4u - fy)) + bot_d[h][k] * (8u * fy);
      start[0][k] = top_c[0][k] * (64u * (4u - fy)) + bot_c[0][k] * (64u * fy) + 4u * step[0][k];  // fx = 4 at pixel 0
      start[1][k] = top_c[1][k] * (64u * (4u - fy)) + bot_c[1][k] * (64u * fy);                    // fx = 0 at pixel 4
    }
    const uint4 *row = reinterpret_cast<const uint4 *>(pv_src_row(p, y) + bx * 8);
    const uint4 u = row[0], v = row[1];
    const uint32_t px[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
    uint32_t bits = 0;
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const int h = x < 4 ? 0 : 1;
      const uint32_t n = static_cast<uint32_t>(x & 3);
      // the quotient by 256 is each lane's high byte: one byte permute divides and re-interleaves (r,g,b,a)
      const uint32_t ca = __byte_perm(start[h][0] + n * step[h][0], start[h][1] + n * step[h][1], 0x7351);
      const uint32_t cb = __byte_perm(start[h][2] + n * step[h][2], start[h][3] + n * step[h][3], 0x7351);
      bits |= pv_pick_modulation(px[x], ca, cb) << (2 * x);
    }
    store(r, y, bits);
  }
}

// The modulation value of ONE pixel (x, y): what pv_modulate_unit computes for pixel 0 of segment x/8 (x % 8 == 0).
__device__ __forceinline__ uint32_t pv_modulate_first_pixel_of_segment(const PvrtcParams &p, uint32_t bx, uint32_t y) {
  const uint32_t lw = p.width >> 3, lh = p.height >> 2;
  const uint32_t top = ((y - 2u) & (p.height - 1u)) >> 2, bottom = (top + 1u) & (lh - 1u);
  const uint32_t cl = (bx + lw - 1u) & (lw - 1u);
  const uint32_t fy = (y + 2u) & 3u;
  const uint2 tl = p.low[top * lw + cl], tr = p.low[top * lw + bx], bl = p.low[bottom * lw + cl], br = p.low[bottom * lw + bx];
  auto vblend = [&](uint32_t t, uint32_t b) {
    const PvLanes lt = pv_split(t), lb = pv_split(b);
    return PvLanes{lt.rb * (4u - fy) + lb.rb * fy, lt.ga * (4u - fy) + lb.ga * fy};
  };
  // pixel 0 of a segment: fx = 4, between columns (bx-1, bx)
  const uint32_t ca = pv_blend256(vblend(tl.x, bl.x), 32u, vblend(tr.x, br.x), 32u);
  const uint32_t cb = pv_blend256(vblend(tl.y, bl.y), 32u, vblend(tr.y, br.y), 32u);
  return pv_pick_modulation(pv_src_row(p, y)[bx * 8], ca, cb);
}

__global__ void ICB_PVRTC_MOD_BOUNDS pvrtc_modulate_kernel(const PvrtcParams p) {
  pv_launch_dependents();
  pv_wait_for_previous();  // Morph's A/B colours
  const uint32_t lw = p.width >> 3;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= lw * p.mod_units) return;
  const uint32_t bx = t & (lw - 1u), y0 = (2u + kModRows * (p.mod_unit0 + (t >> p.lw_shift))) & (p.height - 1u);
  pv_modulate_unit(
      p, bx, y0, [](uint32_t) { return true; },
      [&](uint32_t, uint32_t y, uint32_t bits) { p.mod[y * lw + bx] = static_cast<uint16_t>(bits); });
}

__global__ void __launch_bounds__(128) pvrtc_pack_kernel(const PvrtcParams p) {
  pv_launch_dependents();
  pv_wait_for_previous();  // Modulate's 2-bit values (and, transitively, Morph's colours)
  const uint32_t lw = p.width >> 3, lh = p.height >> 2;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= lw * p.pack_rows) return;
  const uint32_t bx = t & (lw - 1u), by = (p.pack_row0 + (t >> p.lw_shift)) & (lh - 1u);
  const uint32_t right_bx = (bx + 1u) & (lw - 1u);
  uint32_t row[5], right[4];
#pragma unroll
  for (int y = 0; y < 5; ++y) {
    const uint32_t sy = (by * 4u + y) & (p.height - 1u);  // y == 4: the wrapped row below
    row[y] = p.mod[sy * lw + bx];
    if (y < 4) right[y] = p.mod[sy * lw + right_bx] & 3u;  // wrapped pixel to the right of the row
  }
  bool one_bpp;
  const uint32_t mod_bits = pv_pack_modulation(row, right, &one_bpp);
  const uint2 ab = p.low[by * lw + bx];
  const uint32_t colours = pv_pack_colours(ab.x, ab.y, one_bpp);
  p.dst[pv_z_index(bx, by)] = make_uint2(mod_bits, colours);
}

// ---------------------------------------------------------------------------------------------------------
// Modulate + Pack in one kernel (whole images of at least 256 x 32 pixels)
// ---------------------------------------------------------------------------------------------------------
//
// A CTA owns a tile of kFusedBx x kFusedBy blocks (256 x 32 pixels).  Pack needs the modulation values of the tile's
// pixel rows plus the row below and the pixel column to the right (pvrtc_compressor.cc:417-424), so phase 1 computes
// modulation rows 4*by0 .. 4*by0 + 32 for the tile's 32 segments -- warp w takes Modulate unit g = by0 - 1 + w, the
// first and the last warp only the rows that fall into the range -- plus the first pixel of the segment to the right
// for the tile's 32 rows (four lanes of each warp, one row each), all into shared memory; after one __syncthreads
// phase 2 packs the tile's 256 blocks from there.  The reference's byte-per-pixel modulation image never exists, not
// even as 2-bit words in L2, and the third kernel with its launch and tail is gone -- at the price of ~10 % more
// instructions (the extra row and column, two partly idle warps per CTA) in a pipeline that is instruction-issue
// bound: measured 50.4 us against 48.5 us for the three kernels at 4096^2, so the launcher uses it only on request
// (ICB_PVRTC_FUSED=1).
constexpr uint32_t kFusedBx = 32, kFusedBy = 8;
constexpr uint32_t kFusedWarps = kFusedBy + 1, kFusedThreads = 32 * kFusedWarps;
struct PvTileMods {
  uint16_t rows[4 * kFusedBy + 1][kFusedBx];  // eight 2-bit values per word: row (y - 4*by0), segment (bx - bx0)
  uint8_t right[4 * kFusedBy];                // value of the pixel right of the tile, rows 4*by0 .. 4*by0 + 31
};

__device__ __forceinline__ void pv_fused_phase1(const PvrtcParams &p, PvTileMods &tile, uint32_t tile_bx0, uint32_t tile_by0,
                                                uint32_t thread) {
  const uint32_t lw = p.width >> 3, lh = p.height >> 2;
  const uint32_t warp = thread >> 5, lane = thread & 31u;
  const uint32_t g = (tile_by0 + lh - 1u + warp) & (lh - 1u);  // Modulate unit: rows 4g+2 .. 4g+5
  const uint32_t y0 = (4u * g + 2u) & (p.height - 1u);
  // tile-relative index of the unit's first row: -2, 2, 6, ...; rows outside [0, 32] are not needed
  const int rel0 = static_cast<int>(4u * warp) - 2;
  auto wanted = [&](uint32_t r) { return rel0 + static_cast<int>(r) >= 0 && rel0 + static_cast<int>(r) <= static_cast<int>(4u * kFusedBy); };
  pv_modulate_unit(p, tile_bx0 + lane, y0, wanted,
                   [&](uint32_t r, uint32_t, uint32_t bits) { tile.rows[rel0 + static_cast<int>(r)][lane] = static_cast<uint16_t>(bits); });
  // the column right of the tile (wrapped): lane r of this warp takes row r of the unit
  const int rel = rel0 + static_cast<int>(lane);
  if (lane < 4u && rel >= 0 && rel < static_cast<int>(4u * kFusedBy)) {
    const uint32_t y = (y0 + lane) & (p.height - 1u);
    tile.right[rel] = static_cast<uint8_t>(pv_modulate_first_pixel_of_segment(p, (tile_bx0 + kFusedBx) & (lw - 1u), y));
  }
}

__device__ __forceinline__ void pv_fused_phase2(const PvrtcParams &p, const PvTileMods &tile, uint32_t tile_bx0,
                                                uint32_t tile_by0, uint32_t thread) {
  if (thread >= kFusedBx * kFusedBy) return;
  const uint32_t lw = p.width >> 3;
  const uint32_t lbx = thread & (kFusedBx - 1u), lby = thread / kFusedBx;
  uint32_t row[5], right[4];
#pragma unroll
  for (int y = 0; y < 5; ++y) {
    row[y] = tile.rows[4u * lby + y][lbx];
    if (y < 4) right[y] = lbx + 1u < kFusedBx ? (tile.rows[4u * lby + y][lbx + 1u] & 3u) : tile.right[4u * lby + y];
  }
  bool one_bpp;
  const uint32_t mod_bits = pv_pack_modulation(row, right, &one_bpp);
  const uint32_t bx = tile_bx0 + lbx, by = tile_by0 + lby;
  const uint2 ab = p.low[by * lw + bx];
  p.dst[pv_z_index(bx, by)] = make_uint2(mod_bits, pv_pack_colours(ab.x, ab.y, one_bpp));
}

#ifndef ICB_HOST_EMULATION  // (tests/hostemu runs the two phases itself, thread by thread, around its own "barrier")
__global__ void __launch_bounds__(kFusedThreads) pvrtc_modpack_kernel(const PvrtcParams p) {
  __shared__ PvTileMods tile;
  pv_launch_dependents();
  pv_wait_for_previous();  // Morph's A/B colours
  const uint32_t tile_bx0 = blockIdx.x * kFusedBx, tile_by0 = blockIdx.y * kFusedBy;
  pv_fused_phase1(p, tile, tile_bx0, tile_by0, threadIdx.x);
  __syncthreads();
  pv_fused_phase2(p, tile, tile_bx0, tile_by0, threadIdx.x);
}
#endif

// True when the fused kernel covers this launch: whole image, at least one full tile each way.
inline bool pvrtc_use_fused(uint32_t h, uint32_t w, bool whole) { return whole && (w / 8) >= kFusedBx && (h / 4) >= kFusedBy; }

}  // namespace icb
