// pvrtc_kernels.cuh -- image-level PVRTC1 2bpp pipeline (reference: CompressPVRTC_RGBA_2BPP,
// /root/reference/image_compression/internal/pvrtc_compressor.cc:586-597 = Morph :506-521, Modulate :527-540,
// Encode :551-580).
//
// Two kernels instead of the reference's three passes:
//   pvrtc_morph_kernel       one thread per 8x4 block -> bit-reduced A and B colours (two w/8 x h/4 images)
//   pvrtc_modulate_kernel    one thread per block: bilinear upscale of A and B over the block plus the wrapped
//                            pixel column to its right and row below (all the mode decision needs), modulation
//                            choice per pixel, mode + bit packing, store at the block's Z-order slot.  The
//                            per-pixel modulation image of the reference is never written to memory.
#pragma once
#include <cstdint>

#include "pvrtc_encode.cuh"

namespace icb {

struct PvrtcParams {
  const uint32_t *src;  // RGBA8 pixels, row-major, no padding
  uint32_t *low_a;      // (w/8) x (h/4) A colours
  uint32_t *low_b;      // (w/8) x (h/4) B colours
  uint2 *dst;           // w*h/32 blocks in Z-order
  uint32_t width, height;
};

__global__ void __launch_bounds__(128) pvrtc_morph_kernel(const PvrtcParams p) {
  const uint32_t lw = p.width >> 3, lh = p.height >> 2;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= lw * lh) return;
  const uint32_t bx = t % lw, by = t / lw;
  uint32_t px[32];
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    const uint4 *row = reinterpret_cast<const uint4 *>(p.src + static_cast<size_t>(by * 4 + y) * p.width + bx * 8);
    const uint4 u = __ldg(row), v = __ldg(row + 1);
    px[8 * y + 0] = u.x; px[8 * y + 1] = u.y; px[8 * y + 2] = u.z; px[8 * y + 3] = u.w;
    px[8 * y + 4] = v.x; px[8 * y + 5] = v.y; px[8 * y + 6] = v.z; px[8 * y + 7] = v.w;
  }
  uint32_t ca, cb;
  pv_block_extremes(px, __ldg(p.src), &ca, &cb);
  p.low_a[t] = ca;
  p.low_b[t] = cb;
}

__global__ void __launch_bounds__(128) pvrtc_modulate_kernel(const PvrtcParams p) {
  const uint32_t lw = p.width >> 3, lh = p.height >> 2;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= lw * lh) return;
  const uint32_t bx = t % lw, by = t / lw;

  // 3x3 neighbourhood of low-resolution colours, wrapped (pvrtc_compressor.cc:216-223).
  PvLanes na[3][3], nb[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint32_t sx = (bx + lw + i - 1) & (lw - 1), sy = (by + lh + j - 1) & (lh - 1);
      na[j][i] = pv_split(__ldg(p.low_a + sy * lw + sx));
      nb[j][i] = pv_split(__ldg(p.low_b + sy * lw + sx));
    }

  uint32_t m[5][9];
#pragma unroll
  for (int y = 0; y < 5; ++y) {
    // Source row (wrapped below the image) and the vertical blend shared by the whole row.
    const uint32_t sy = (by * 4 + y) & (p.height - 1);
    const uint32_t *row = p.src + static_cast<size_t>(sy) * p.width;
    const int top = (y & 3) < 2 ? (y >> 2) : (y >> 2) + 1;  // rows y=0,1 use (by-1,by); 2,3 (by,by+1); 4 like 0 of next
    const uint32_t fy = (y + 2) & 3;
    // For y == 4 the pixel belongs to block by+1, whose neighbourhood is shifted one low-res row down; its
    // rows (by, by+1) are still inside our 3x3 window.
    PvLanes va[3], vb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      va[i].rb = na[top][i].rb * (4u - fy) + na[top + 1][i].rb * fy;
      va[i].ga = na[top][i].ga * (4u - fy) + na[top + 1][i].ga * fy;
      vb[i].rb = nb[top][i].rb * (4u - fy) + nb[top + 1][i].rb * fy;
      vb[i].ga = nb[top][i].ga * (4u - fy) + nb[top + 1][i].ga * fy;
    }
#pragma unroll
    for (int x = 0; x < 9; ++x) {
      if (y == 4 && x == 8) continue;  // corner is never read
      const int left = (x & 7) < 4 ? (x >> 3) : (x >> 3) + 1;
      const uint32_t fx = (x + 4) & 7;
      const uint32_t sx = (bx * 8 + x) & (p.width - 1);
      const uint32_t pixel = __ldg(row + sx);
      const PvLanes ca = pv_mix(va[left], 8u - fx, va[left + 1], fx, 5u);
      const PvLanes cb = pv_mix(vb[left], 8u - fx, vb[left + 1], fx, 5u);
      m[y][x] = pv_pick_modulation(pixel, ca, cb);
    }
  }
  m[4][8] = 0;
  bool one_bpp;
  const uint32_t mod_bits = pv_pack_modulation(m, &one_bpp);
  const uint32_t colours = pv_pack_colours(pv_join(na[1][1]), pv_join(nb[1][1]), one_bpp);
  p.dst[pv_z_index(bx, by)] = make_uint2(mod_bits, colours);
}

}  // namespace icb
