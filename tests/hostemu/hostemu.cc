// hostemu.cc -- TEST INFRASTRUCTURE ONLY: the device code of image_compression_b200/csrc (encoders, decoders, PVRTC
// kernels, compressed-domain operations -- the very headers nvcc compiles into libicb200.so) compiled for the host
// with the CUDA intrinsics emulated (cuda_emulation.h), so that CPU-only test runs can step through the GPU
// arithmetic and compare it with the oracle.  It checks the ARITHMETIC of the device code; the TMA ring drivers,
// launch configuration and memory system are only exercised on a GPU (tests/*_gpu.py).
//
// Built by tests/test_hostemu.py (g++, one translation unit) into tests/hostemu/libhostemu.so; never linked with or
// loaded by the product library, bench.py or __graft_entry__.py.
//
// A kernel launch <<<grid, block>>> becomes nested loops over blockIdx / threadIdx; the kernels used here have no
// intra-block synchronisation, so running their threads one after another is an exact emulation.
#include "cuda_emulation.h"

#include <vector>

#include "../../image_compression_b200/csrc/block4x4_generic.cuh"
#include "../../image_compression_b200/csrc/blockops_kernels.cuh"
#include "../../image_compression_b200/csrc/decode4x4_kernels.cuh"
#include "../../image_compression_b200/csrc/pvrtc_kernels.cuh"

namespace {

template <typename F>
void launch(uint32_t grid_x, uint32_t grid_y, uint32_t block_x, F kernel) {
  gridDim = EmuDim3{grid_x, grid_y, 1};
  blockDim = EmuDim3{block_x, 1, 1};
  for (uint32_t by = 0; by < grid_y; ++by)
    for (uint32_t bx = 0; bx < grid_x; ++bx)
      for (uint32_t tx = 0; tx < block_x; ++tx) {
        blockIdx = EmuDim3{bx, by, 1};
        threadIdx = EmuDim3{tx, 0, 0};
        kernel();
      }
}

uint32_t blocks_of(uint32_t n) { return (n + 3) / 4; }

template <int kCodec, int kNcomp>
void run_generic(const icb::Encode4x4Params &p) {
  launch(3, 1, 128, [&] { icb::encode4x4_generic_kernel<kCodec, kNcomp>(p); });  // grid-stride: any grid covers the image
}

constexpr int kOk = 0, kInvalid = -1, kUnsupported = -4;  // include/icb200.h status values

}  // namespace

extern "C" {

// What the other lanes of the (emulated) warp answer to a vote: 0 = like this lane, 1 = "no", 2 = "no" to the
// thread's first vote only (cuda_emulation.h).
// 1: whole images keep the three-kernel PVRTC pipeline (Modulate and Pack as separate kernels, the library's default)
// where the fused Modulate/Pack kernel could run, so that both are checked on the same inputs.
bool g_emu_pvrtc_unfused = false;
void emu_set_pvrtc_unfused(int on) { g_emu_pvrtc_unfused = on != 0; }
void emu_set_vote(int vote) {
  // 1000 + m: every vote agrees, and the rest of the warp contributes m to a warp-wide minimum (ETC1's margin + 128)
  g_emu_min_cap = vote >= 1000 ? static_cast<uint32_t>(vote - 1000) : 0xffffffffu;
  if (vote >= 1000) vote = 0;
  g_emu_vote = vote == 2 ? kEmuVoteFirstNo : (vote ? kEmuVoteNo : kEmuVoteAgree);
  g_emu_votes_cast = 0;
}

// Spot checks of the emulated instructions against values worked out by hand from the PTX ISA definitions.
// Returns 0, or the number of the first check that failed.
int emu_self_check(void) {
  int n = 0;
#define CHECK(cond) \
  do {              \
    ++n;            \
    if (!(cond)) return n; \
  } while (0)
  CHECK(__byte_perm(0x33221100u, 0x77665544u, 0x5410) == 0x55441100u);
  CHECK(__byte_perm(0x33221100u, 0x77665544u, 0x7351) == 0x77335511u);
  CHECK(__byte_perm(0xddccbbaau, 0u, 0x0123) == 0xaabbccddu);
  CHECK(__dp4a(0x04030201u, 0x01010101u, 5u) == 15u);
  CHECK(__dp4a(0xff00ff00u, 0xff00ff00u, 0u) == 2u * 255u * 255u);
  CHECK(__funnelshift_r(0x00000004u, 0x00000003u, 2) == 0xc0000001u);
  CHECK(__vminu2(0x0005fff0u, 0xfff00005u) == 0x00050005u);
  CHECK(__vimax3_u16x2(0x00010009u, 0x00080002u, 0x00030003u) == 0x00080009u);
  CHECK(__viaddmin_s32_relu(250, 9, 255) == 255 && __viaddmin_s32_relu(3, -9, 255) == 0 && __viaddmin_s32_relu(3, 9, 255) == 12);
  CHECK(__viaddmin_s16x2_relu(0x00fa0003u, 0xfff7fff7u, 0x00ff00ffu) == 0x00f10000u);  // (250-9, 3-9) -> (241, 0)
  CHECK(__viaddmin_s16x2_relu(0x00fa0003u, 0x00090009u, 0x00ff00ffu) == 0x00ff000cu);
  CHECK(__vabsdiffu4(0x10ff0005u, 0x20000a05u) == 0x10ff0a00u);
  CHECK(__vsadu4(0x10ff0005u, 0x20000a05u) == 0x10u + 0xffu + 0x0au);
  CHECK(__umulhi(765u, 683u << 21) == 255u);
  CHECK(__usad(3u, 10u, 1u) == 8u);
  // packed half: 0x6500 = 1280, 0x3c00 = 1, 0xbc00 = -1, 0x6400 = 1024, 0x7bff = 65504
  using namespace icb_emu;
  CHECK(half_to_double(0x6500) == 1280.0 && half_to_double(0x3c00) == 1.0 && half_to_double(0xbc00) == -1.0);
  CHECK(half_to_double(0x0001) == std::ldexp(1.0, -24) && half_to_double(0x7bff) == 65504.0);
  for (uint32_t h = 0; h < 0x7c00u; ++h) CHECK(double_to_half(half_to_double(static_cast<uint16_t>(h))) == h);  // every finite half
  CHECK(double_to_half(2049.0) == double_to_half(2048.0));  // tie to even (spacing 2 above 2048)
  CHECK(double_to_half(2051.0) == double_to_half(2052.0));
  CHECK(double_to_half(65519.9) == 0x7bffu && double_to_half(65520.0) == 0x7c00u && double_to_half(-1e9) == 0xfc00u);
  CHECK(add_f16x2(0x65ff6500u, 0xe5fee500u, true) == 0x3c000000u);   // (1535-1534, 1280-1280) clamped to [0,1]
  CHECK(add_f16x2(0x65006500u, 0xe5fee5feu, true) == 0x00000000u);   // negative -> 0
  CHECK(add_f16x2(0x64006400u, 0x3c000000u, false) == 0x64016400u);  // 1024 + 1 (exact: spacing 1 at 1024)
  CHECK(fma_f16x2(0x3c003c00u, 0xdbf8dbf8u, 0x65ff65ffu, false) == 0x65006500u);  // 1 * -255 + 1535 = 1280
  CHECK(fma_f16x2(0x65806580u, 0x3c003c00u, 0xe500e4ffu, true) == 0x3c003c00u);   // 1408 - 1279 / - 1280 -> 1
  CHECK(fma_f16x2(0x40004000u, 0x40004000u, 0x3c00bc00u, false) == 0x45004200u);  // 2*2-1 = 3, 2*2+1 = 5
#undef CHECK
  return 0;
}

// The generic image driver over the whole coded grid (icb_api.cu:encode4x4 hands it every block the TMA tiles do not
// cover; here it gets all of them).  codec: 0 DXT1, 1 DXT5, 2 ETC1.
int emu_encode4x4(int codec, int ncomp, const uint8_t *src, uint32_t h, uint32_t w, uint32_t pitch, uint32_t coded_h,
                  uint32_t coded_w, int swap_rb, int strategy, uint8_t *dst) {
  icb::Encode4x4Params p;
  p.src = src;
  p.dst = dst;
  p.height = h;
  p.width = w;
  p.pitch = pitch;
  p.grid_cols = blocks_of(coded_w);
  p.row0 = 0;
  p.row1 = blocks_of(coded_h);
  p.col0 = 0;
  p.col1 = p.grid_cols;
  p.swap_rb = swap_rb;
  p.etc_strategy = strategy;
  if (codec == 0 && ncomp == 3) run_generic<icb::kCodecDxt1, 3>(p);
  else if (codec == 0 && ncomp == 4) run_generic<icb::kCodecDxt1, 4>(p);
  else if (codec == 1 && ncomp == 4) run_generic<icb::kCodecDxt5, 4>(p);
  else if (codec == 2 && ncomp == 3) run_generic<icb::kCodecEtc1, 3>(p);
  else return kInvalid;
  return kOk;
}

// The RGB888 -> DXT1 form the TMA consumers use (block4x4_kernels.cuh): keys straight from the three 32-bit words of
// each block row.  h and w must be multiples of 4 (the TMA drivers only see whole blocks).
int emu_dxt1_rgb888_rows(const uint8_t *src, uint32_t h, uint32_t w, uint32_t pitch, int swap_rb, uint8_t *dst) {
  if (h % 4 || w % 4) return kInvalid;
  for (uint32_t br = 0; br < h / 4; ++br)
    for (uint32_t bc = 0; bc < w / 4; ++bc) {
      const uint8_t *win = src + static_cast<size_t>(4 * br) * pitch + 12 * bc;
      uint32_t rows[4][3];
      for (int y = 0; y < 4; ++y) std::memcpy(rows[y], win + static_cast<size_t>(y) * pitch, 12);
      auto fetch = [&](uint32_t i) {
        const uint8_t *q = win + static_cast<size_t>(i >> 2) * pitch + (i & 3u) * 3u;
        return static_cast<uint32_t>(q[0]) | (static_cast<uint32_t>(q[1]) << 8) | (static_cast<uint32_t>(q[2]) << 16);
      };
      const uint2 c = icb::dxt1_encode_rgb888_rows<true>(rows, swap_rb != 0, false, fetch);
      std::memcpy(dst + (static_cast<size_t>(br) * (w / 4) + bc) * 8, &c, 8);
    }
  return kOk;
}

// PVRTC: the three kernels with the launch shapes of icb_api.cu:pvrtc_launch.  nstripes == 1: whole image.  Otherwise
// the image is cut into nstripes runs of block rows and each is encoded the way icb_pvrtc2_encode_stripe does it: from
// a private copy of its own pixel rows plus one halo block row above and below (wrapped), blocks stored at their
// Z-order slot of the shared output.
int emu_pvrtc2(const uint8_t *src, uint32_t h, uint32_t w, uint32_t nstripes, uint8_t *dst) {
  if (h == 0 || w != h || (w & (w - 1)) || w < 8) return kUnsupported;
  const uint32_t lw = w / 8, lh = h / 4;
  std::vector<uint8_t> scratch(static_cast<size_t>(lw) * lh * 8 + static_cast<size_t>(lw) * h * 2);
  auto run = [&](const icb::PvrtcParams &p) {
    launch((lw * p.morph_rows + 127) / 128, 1, 128, [&] { icb::pvrtc_morph_kernel(p); });
    if (icb::pvrtc_use_fused(h, w, p.morph_rows == lh && p.src_row0 == 0 && p.pack_rows == lh) && !g_emu_pvrtc_unfused) {
      // the fused Modulate + Pack kernel, one "CTA" at a time: phase 1 for every thread, the barrier, phase 2
      for (uint32_t cy = 0; cy < lh / icb::kFusedBy; ++cy)
        for (uint32_t cx = 0; cx < lw / icb::kFusedBx; ++cx) {
          icb::PvTileMods tile;
          std::memset(&tile, 0xEE, sizeof(tile));
          for (uint32_t t = 0; t < icb::kFusedThreads; ++t) icb::pv_fused_phase1(p, tile, cx * icb::kFusedBx, cy * icb::kFusedBy, t);
          for (uint32_t t = 0; t < icb::kFusedThreads; ++t) icb::pv_fused_phase2(p, tile, cx * icb::kFusedBx, cy * icb::kFusedBy, t);
        }
      return;
    }
    launch((lw * p.mod_units + icb::kModThreads - 1) / icb::kModThreads, 1, icb::kModThreads, [&] { icb::pvrtc_modulate_kernel(p); });
    launch((lw * p.pack_rows + 127) / 128, 1, 128, [&] { icb::pvrtc_pack_kernel(p); });
  };
  if (nstripes <= 1) {
    run(icb::pvrtc_make_params(src, src, scratch.data(), dst, h, w, 0, 0, lh, true));
    return kOk;
  }
  if (nstripes > lh) return kInvalid;
  for (uint32_t s = 0; s < nstripes; ++s) {
    const uint32_t r0 = static_cast<uint32_t>(static_cast<uint64_t>(lh) * s / nstripes);
    const uint32_t r1 = static_cast<uint32_t>(static_cast<uint64_t>(lh) * (s + 1) / nstripes);
    if (r1 - r0 + 2 > lh) return kInvalid;  // icb_pvrtc2_encode_stripe's rule
    const uint32_t src_row0 = 4 * ((r0 + lh - 1) & (lh - 1)), nrows = 4 * (r1 - r0 + 2);
    std::vector<uint8_t> rows(static_cast<size_t>(nrows) * w * 4);
    for (uint32_t y = 0; y < nrows; ++y)
      std::memcpy(rows.data() + static_cast<size_t>(y) * w * 4, src + static_cast<size_t>((src_row0 + y) & (h - 1)) * w * 4,
                  static_cast<size_t>(w) * 4);
    uint32_t first_pixel;
    std::memcpy(&first_pixel, src, 4);
    std::fill(scratch.begin(), scratch.end(), 0xEE);  // a stripe must not depend on what another one left behind
    run(icb::pvrtc_make_params(rows.data(), &first_pixel, scratch.data(), dst, h, w, src_row0, r0, r1, false));
  }
  return kOk;
}

// codec: 0 DXT1 (RGB888 out), 1 DXT5 (RGBA out), 2 ETC1 (RGB888 out); launch shape of icb_api.cu:icb_decode4x4.
int emu_decode4x4(int codec, const uint8_t *blocks, uint32_t h, uint32_t w, uint32_t block_cols, int swap_rb, uint8_t *dst,
                  uint32_t pitch) {
  icb::Decode4x4Params p;
  p.blocks = blocks;
  p.dst = dst;
  p.height = h;
  p.width = w;
  p.pitch = pitch;
  p.block_cols = block_cols;
  p.block_rows = blocks_of(h);
  p.swap_rb = swap_rb;
  const uint32_t gx = (block_cols + 127) / 128, gy = p.block_rows < 3 ? p.block_rows : 3;
  if (codec == 0) launch(gx, gy, 128, [&] { icb::decode4x4_kernel<icb::kCodecDxt1>(p); });
  else if (codec == 1) launch(gx, gy, 128, [&] { icb::decode4x4_kernel<icb::kCodecDxt5>(p); });
  else if (codec == 2) launch(gx, gy, 128, [&] { icb::decode4x4_kernel<icb::kCodecEtc1>(p); });
  else return kInvalid;
  return kOk;
}

// Downsample with icb_downsample4x4's grid rules and refusals.  out must hold out_rows * out_cols blocks.
int emu_downsample4x4(int codec, int strategy, const uint8_t *blocks, uint32_t h, uint32_t w, uint8_t *out) {
  icb::Downsample4x4Params p;
  p.in = blocks;
  p.out = out;
  p.in_rows = blocks_of(h);
  p.in_cols = blocks_of(w);
  if ((p.in_rows > 1 && p.in_rows % 2) || (p.in_cols > 1 && p.in_cols % 2)) return kUnsupported;
  if (p.in_rows == 1 && p.in_cols == 1 && (h == 3 || w == 3)) return kUnsupported;
  p.out_rows = p.in_rows > 1 ? p.in_rows / 2 : 1;
  p.out_cols = p.in_cols > 1 ? p.in_cols / 2 : 1;
  p.height = h;
  p.width = w;
  p.etc_strategy = strategy;
  if (codec == 0) launch(2, 1, icb::kBlockOpThreads, [&] { icb::downsample4x4_kernel<icb::kCodecDxt1>(p); });
  else if (codec == 1) launch(2, 1, icb::kBlockOpThreads, [&] { icb::downsample4x4_kernel<icb::kCodecDxt5>(p); });
  else if (codec == 2) launch(2, 1, icb::kBlockOpThreads, [&] { icb::downsample4x4_kernel<icb::kCodecEtc1>(p); });
  else return kInvalid;
  return kOk;
}

// Pad with icb_pad4x4's rules (ch, cw: compressed size; ph, pw: padded size).
int emu_pad4x4(int codec, int strategy, const uint8_t *blocks, uint32_t ch, uint32_t cw, uint32_t ph, uint32_t pw, uint8_t *out) {
  const size_t block_bytes = codec == 1 ? 16 : 8;
  icb::Pad4x4Params p;
  p.in = blocks;
  p.out = out;
  p.in_rows = blocks_of(ch);
  p.in_cols = blocks_of(cw);
  if (ch >= ph && cw >= pw) {
    std::memcpy(out, blocks, static_cast<size_t>(p.in_rows) * p.in_cols * block_bytes);
    return kOk;
  }
  p.out_rows = blocks_of(ph);
  p.out_cols = blocks_of(pw);
  if (p.out_rows < p.in_rows || p.out_cols < p.in_cols) return kUnsupported;
  p.etc_strategy = strategy;
  if (codec == 0) launch(2, 1, icb::kBlockOpThreads, [&] { icb::pad4x4_kernel<icb::kCodecDxt1>(p); });
  else if (codec == 1) launch(2, 1, icb::kBlockOpThreads, [&] { icb::pad4x4_kernel<icb::kCodecDxt5>(p); });
  else if (codec == 2) launch(2, 1, icb::kBlockOpThreads, [&] { icb::pad4x4_kernel<icb::kCodecEtc1>(p); });
  else return kInvalid;
  return kOk;
}

int emu_fill_solid4x4(int codec, uint32_t colour, uint64_t num_blocks, uint8_t *out) {
  if (codec == 0) launch(2, 1, 256, [&] { icb::fill_solid4x4_kernel<icb::kCodecDxt1>(out, num_blocks, colour); });
  else if (codec == 1) launch(2, 1, 256, [&] { icb::fill_solid4x4_kernel<icb::kCodecDxt5>(out, num_blocks, colour); });
  else if (codec == 2) launch(2, 1, 256, [&] { icb::fill_solid4x4_kernel<icb::kCodecEtc1>(out, num_blocks, colour); });
  else return kInvalid;
  return kOk;
}

int emu_transcode_dxt1_to_etc1(uint8_t *blocks, uint64_t num_blocks) {
  launch(2, 1, icb::kBlockOpThreads, [&] { icb::transcode_dxt1_to_etc1_kernel(blocks, num_blocks); });
  return kOk;
}

}  // extern "C"
