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
// Row v: {5-bit 1/3 e0,e1, 5-bit 1/2 e0,e1, 6-bit 1/3 e0,e1, 6-bit 1/2 e0,e1}.
struct DxtConstEndpoints {
  uint8_t v[256][8];
};
constexpr DxtConstEndpoints kDxtConstEndpoints = {{
#include "dxt_const_table.inc"
}};

// Blinn rounded quantiser: round(v * max / 255)  (internal/color_util.h:156-164).
constexpr __host__ __device__ uint32_t quant_round(uint32_t v, uint32_t maxv) {
  const uint32_t i = v * maxv + 128u;
  return (i + (i >> 8)) >> 8;
}

__device__ __forceinline__ uint32_t to_565(uint32_t r, uint32_t g, uint32_t b) {
  return (quant_round(r, 31u) << 11) | (quant_round(g, 63u) << 5) | quant_round(b, 31u);
}

constexpr __host__ __device__ uint32_t expand5(uint32_t v) { return (v << 3) | (v >> 2); }
constexpr __host__ __device__ uint32_t expand6(uint32_t v) { return (v << 2) | (v >> 4); }

// floor(x / 3) for 0 <= x <= 765 (one IMAD + one shift).
__device__ __forceinline__ uint32_t div3_small(uint32_t x) { return (x * 683u) >> 11; }

// What the constant-colour search (GetBestDxtcConstColors, internal/dxtc_const_color_table.cc:322-392) needs to know
// about one 8-bit channel value v, derived from the endpoint table at compile time.  The search compares
// (4|dr| + 8|dg| + |db|)^2 of three candidate colours -- v quantised, the 1/2 blend and the 1/3 blend of the table's
// endpoint pairs -- and every |d| depends on its own channel only, so it is tabulated: one 16-byte row per value,
//   bytes 0-3   5-bit endpoints  {1/3 e0, 1/3 e1, 1/2 e0, 1/2 e1}          (v as red or blue)
//   bytes 4-7   6-bit endpoints  {1/3 e0, 1/3 e1, 1/2 e0, 1/2 e1}          (v as green)
//   bytes 8-11  5-bit {|v - expand(q)|, |v - 1/2 blend|, |v - 1/3 blend|, q = quantised v}
//   bytes 12-15 6-bit, the same.
// No |d| exceeds 4 (checked below), so the three weighted sums of a colour fit side by side in the bytes of one word.
struct DxtConstRows {
  uint8_t v[256][16];
};
constexpr uint32_t dxt_abs_diff(uint32_t a, uint32_t b) { return a > b ? a - b : b - a; }
constexpr DxtConstRows make_dxt_const_rows() {
  DxtConstRows t = {};
  for (uint32_t v = 0; v < 256; ++v) {
    const uint8_t(&e)[8] = kDxtConstEndpoints.v[v];
    for (int k = 0; k < 8; ++k) t.v[v][k] = e[k];
    const uint32_t q5 = quant_round(v, 31u), q6 = quant_round(v, 63u);
    t.v[v][8] = static_cast<uint8_t>(dxt_abs_diff(v, expand5(q5)));
    t.v[v][9] = static_cast<uint8_t>(dxt_abs_diff(v, (expand5(e[2]) + expand5(e[3])) >> 1));
    t.v[v][10] = static_cast<uint8_t>(dxt_abs_diff(v, (2u * expand5(e[0]) + expand5(e[1])) / 3u));
    t.v[v][11] = static_cast<uint8_t>(q5);
    t.v[v][12] = static_cast<uint8_t>(dxt_abs_diff(v, expand6(q6)));
    t.v[v][13] = static_cast<uint8_t>(dxt_abs_diff(v, (expand6(e[6]) + expand6(e[7])) >> 1));
    t.v[v][14] = static_cast<uint8_t>(dxt_abs_diff(v, (2u * expand6(e[4]) + expand6(e[5])) / 3u));
    t.v[v][15] = static_cast<uint8_t>(q6);
  }
  return t;
}
constexpr uint32_t dxt_const_rows_max_diff(const DxtConstRows &t) {
  uint32_t m = 0;
  for (int v = 0; v < 256; ++v)
    for (int k : {8, 9, 10, 12, 13, 14}) m = t.v[v][k] > m ? t.v[v][k] : m;
  return m;
}
constexpr DxtConstRows kDxtConstRows = make_dxt_const_rows();
static_assert(13u * dxt_const_rows_max_diff(kDxtConstRows) < 256u, "weighted error sums must fit one byte each");
__device__ __align__(16) const DxtConstRows g_dxt_const_rows = kDxtConstRows;

// Constant-colour block (rare path, divergent by design).  `t` is the target colour as bytes (c0,c1,c2) in the order
// the reference looks it up (memory order, see the caller).  Returns c0 | c1 << 16 | index << 32: the two 565
// endpoints and the 2-bit index to replicate.  (Packed return value: pointer out-parameters of a non-inlined function
// would force the caller's registers through local memory on every block, not just on constant ones.)
// Squared errors compare like their roots, so the search compares the weighted sums themselves; three 16-byte loads
// and about forty instructions replace the expansion, blending and squaring of the three candidates.
__device__ __noinline__ uint64_t dxt_const_colour(uint32_t t, bool always4) {
  const uint4 *rows = reinterpret_cast<const uint4 *>(g_dxt_const_rows.v);
  const uint4 r = rows[t & 255u], g = rows[(t >> 8) & 255u], b = rows[(t >> 16) & 255u];
  // byte 0: quantised colour, byte 1: 1/2 blend (three-colour block, DXT1 only), byte 2: 1/3 blend; byte 3 is junk
  const uint32_t sums = (r.z << 2) + (g.w << 3) + b.z;
  const uint32_t err_q = sums & 255u, err_half = (sums >> 8) & 255u, err_third = (sums >> 16) & 255u;
  const bool half = !always4 && err_half < err_q;                  // strict, in the reference's order of trials
  const bool third = err_third < (half ? err_half : err_q);
  // endpoint bytes of the chosen blend: (e0, e1) are bytes (0,1) of the row word for thirds, (2,3) for halves
  const uint32_t shift = third ? 0u : 16u;
  const uint32_t er = r.x >> shift, eg = g.y >> shift, eb = b.x >> shift;
  const uint32_t p0 = ((er & 255u) << 11) | ((eg & 255u) << 5) | (eb & 255u);
  const uint32_t p1 = (((er >> 8) & 255u) << 11) | (((eg >> 8) & 255u) << 5) | ((eb >> 8) & 255u);
  uint32_t c0 = ((r.z >> 24) << 11) | ((g.w >> 24) << 5) | (b.z >> 24), c1 = c0, which = 0;
  if (third) {
    which = p0 > p1 ? 2u : 3u;
    c0 = p0 > p1 ? p0 : p1;
    c1 = p0 > p1 ? p1 : p0;
  } else if (half) {
    which = 2;
    c0 = p0 < p1 ? p0 : p1;
    c1 = p0 < p1 ? p1 : p0;
  }
  return c0 | (c1 << 16) | (static_cast<uint64_t>(which) << 32);
}

// Weights for IDP.4A: 8*(4,8,1) on the logical (r,g,b); the alpha byte always gets weight 0.  (Scale 8, not 16: the
// keys 8*lum + index stay below 2^15 -- lum <= 3315 -- so they are also valid SIGNED 16-bit lanes, which the
// index search's VIADDMNMX.S16x2 needs; half-integer crossing points only need a scale of 2.)
__device__ __forceinline__ uint32_t dxt_lum_weights(bool swap_rb) { return swap_rb ? 0x00204008u : 0x00084020u; }

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
//   * walking upwards, candidate b replaces the current one a iff 2l > L_a + L_b, or 2l == L_a + L_b and b has the
//     smaller index (that is what "first strict minimum" does with a tie); candidates with equal luminance are
//     represented by their smallest index;
//   * a pixel's index is the start index plus the index changes of the crossings it has passed.
// Keys.  kf[i] = 8*lum(pixel i) + (i & 7): produced by the IDP.4A that computes the luminance, with the lane index in
// its accumulator -- a 16-bit key below 2^15, i.e. also a valid SIGNED lane.  (The rare general path reads a key as the
// float 2^23 + key: one byte permute with the exponent word.)
// Pixel pair (k, k+8) shares one register, pixel k in the low lane: the low lane of an accumulator then collects the
// index bits of pixels 0..7 and the high lane those of pixels 8..15, i.e. the finished 32-bit index word.
constexpr uint32_t kDxtLumBias = 0x4b000000u;
constexpr uint32_t kDxtKeyStep = 8u;  // key units per unit of luminance

// Round 2 measured the alternatives on B200 (profiles/r02b_driver_ab.txt): the same search on fp32 (five FADD/FFMA per
// pixel, nothing on the integer pipe -- the round-1 form) for all, half or every other pixel pair is slower for every
// DXT kernel (DXT5 87.4 / 83.9 / 83.9 us against 82.7; RGB888 48.3 / 45.1 / 47.3 against 43.0), and so are forms that move
// the key packing and the reversed index to the FMA pipe behind an opaque multiplier: these kernels are bound by the
// NUMBER of instructions issued, not by one pipe.
// The accumulator of pixel i's key IDP: its lane index (pixel i sits in lane i / 8 of register i % 8).
constexpr __host__ __device__ uint32_t dxt_key_seed(int i) { return static_cast<uint32_t>(i & 7); }

// A caller that stages pixels in a buffer it wants back early passes `release`; it is called once, right after the
// two base colours have been re-read through `fetch` (nothing reads the staged pixels after that).
struct NoRelease {
  __device__ __forceinline__ void operator()() const {}
};

template <bool kFullWarp, typename Fetch, typename Release = NoRelease>
__device__ __forceinline__ uint2 dxt1_encode_from_keys(const uint32_t (&kf)[16], bool swap_rb, bool always4, Fetch fetch,
                                                       Release release = Release()) {
  // First minimum / first maximum in raster order: 16-bit keys 8*lum + k, pixel k in the low lane and pixel k + 8 in
  // the high lane of register k; the maximum uses the index field reversed (^7) so that ties resolve to the lowest
  // index.  VIMNMX3.U16x2 folds two more registers (four pixels) per instruction.
  uint32_t pk[8], xk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    pk[k] = kf[k + 8] * 65536u + kf[k];  // bare keys < 2^15: one IMAD joins the pair
    // Index field reversed for the maximum: the field of register k holds exactly k in both lanes, so k ^ 7 = 7 - k is
    // the addition of 7 - 2k to each lane, and since no lane goes negative (key >= k) the two lane additions are one
    // 32-bit addition.
    xk[k] = pk[k] + static_cast<uint32_t>(7 - 2 * k) * 0x10001u;
  }
  uint32_t mn = __vimin3_u16x2(pk[0], pk[1], pk[2]);
  mn = __vimin3_u16x2(mn, pk[3], pk[4]);
  mn = __vimin3_u16x2(mn, pk[5], pk[6]);
  mn = __vminu2(mn, pk[7]);
  uint32_t mx = __vimax3_u16x2(xk[0], xk[1], xk[2]);
  mx = __vimax3_u16x2(mx, xk[3], xk[4]);
  mx = __vimax3_u16x2(mx, xk[5], xk[6]);
  mx = __vmaxu2(mx, xk[7]);
  // Between the lanes: every pixel of the low lane precedes every pixel of the high lane, so the low lane wins ties
  // of luminance whatever the index fields say.
  const uint32_t mn_lo = mn & 0xffffu, mn_hi = mn >> 16, mx_lo = mx & 0xffffu, mx_hi = mx >> 16;
  const bool min_in_lo = mn_lo <= (mn_hi | 7u), max_in_lo = (mx_lo | 7u) >= mx_hi;
  const uint32_t kmin = min_in_lo ? mn_lo : mn_hi, kmax = max_in_lo ? mx_lo : mx_hi;
  const uint32_t imin = (kmin & 7u) + (min_in_lo ? 0u : 8u), imax = ((kmax & 7u) ^ 7u) + (max_in_lo ? 0u : 8u);
  uint32_t p0 = fetch(imin), p1 = fetch(imax);  // base colours, memory byte order
  release();
  const uint32_t lum0 = kmin & 0xfff8u, lum1 = kmax & 0xfff8u;      // 8 * luminance of p0 / p1: lum0 <= lum1
  // The base colours' channels on lanes: memory bytes 0 and 2 (red and blue, or blue and red when swap_rb) side by side
  // in the 16-bit lanes of one register, byte 1 (green) alone in another.  Quantisation and interpolation then cost one
  // multiply-add for two channels, with immediate multipliers -- the per-channel IDP.4A form needed six weight
  // constants in uniform registers, which ptxas re-materialises for every block.
  const uint32_t rb0 = p0 & 0x00ff00ffu, rb1 = p1 & 0x00ff00ffu;
  const uint32_t g0 = __byte_perm(p0, 0u, 0x4441), g1 = __byte_perm(p1, 0u, 0x4441);
  // 565 quantisation: round(v*31/255) == (v*249 + 1024) >> 11, round(v*63/255) == (v*253 + 512) >> 10 for every 8-bit v
  // (tests/test_host_math.py); v*249 + 1024 < 2^16, so the two lanes do not meet.
  const uint32_t x0 = rb0 * 249u + 0x04000400u, x1 = rb1 * 249u + 0x04000400u;
  const uint32_t y0 = g0 * 253u + 512u, y1 = g1 * 253u + 512u;
  const uint32_t q0 = swap_rb ? (((x0 >> 16) & 0xf800u) | ((y0 >> 5) & 0x07e0u) | ((x0 >> 11) & 0x1fu))
                              : ((x0 & 0xf800u) | ((y0 >> 5) & 0x07e0u) | (x0 >> 27));
  const uint32_t q1 = swap_rb ? (((x1 >> 16) & 0xf800u) | ((y1 >> 5) & 0x07e0u) | ((x1 >> 11) & 0x1fu))
                              : ((x1 & 0xf800u) | ((y1 >> 5) & 0x07e0u) | (x1 >> 27));
  // Everything up to the warp vote below is computed for constant blocks too (and ignored): the vote has to sit
  // where the warp has not yet diverged on "is this block constant".
  const bool constant = q0 == q1;
  // The reference swaps the two base colours (565 and 8-bit alike) when c0 < c1, so that c0 > c1 (four-colour mode).
  // Nothing is swapped here: with p0 the darker base colour always, the candidates in luminance order are
  // (p0, (2 p0 + p1)/3, (p0 + 2 p1)/3, p1) either way -- the interpolants of the swapped pair are the same two colours
  // in the other order -- and the swap only renames the indices, 0 <-> 1 and 2 <-> 3: the low bit of every index
  // flips (one XOR at the end) and ties go to the OTHER candidate of the middle pair (one selected constant).
  const bool swapped = q0 < q1;
  uint32_t c0 = max(q0, q1), c1 = min(q0, q1);
  // Interpolants come from the UNQUANTISED base colours, channel by channel with truncation.  floor(x / 3) =
  // umulhi(x, 683 << 21) for x <= 765; for the channel in the HIGH lane of a sum word s = lo + 65536 * hi the shift is
  // part of the multiplier: umulhi(s, 683 << 5) = floor(hi * 683 / 2048 + lo * 683 / 2^27) = floor(hi / 3), because the
  // two error terms add up to less than 0.13 (tests/test_host_math.py runs every lo, hi).
  constexpr uint32_t kThird = 683u << 21, kThirdOfHighLane = 683u << 5;
  const uint32_t s_rb = 2u * rb0 + rb1, s_g = 2u * g0 + g1;    // the interpolant next to p0
  const uint32_t t_rb = rb0 + 2u * rb1, t_g = g0 + 2u * g1;    // the interpolant next to p1
  const uint32_t w_lo = swap_rb ? 8u : 32u, w_hi = swap_rb ? 32u : 8u;  // 8 * (4, 8, 1) on (red, green, blue)
  const uint32_t lum2 = w_lo * __umulhi(s_rb & 0xffffu, kThird) + 64u * __umulhi(s_g, kThird) + w_hi * __umulhi(s_rb, kThirdOfHighLane);
  const uint32_t lum3 = w_lo * __umulhi(t_rb & 0xffffu, kThird) + 64u * __umulhi(t_g, kThird) + w_hi * __umulhi(t_rb, kThirdOfHighLane);
  // Usual case, decided once per warp so the branch never diverges: the interpolants lie strictly between the
  // base colours, i.e. the candidates are ordered along the luminance line with no two equal.  Then the crossing
  // order, the tie rules and the index changes are fixed and only the three midpoints have to be computed.
  // (Constant blocks vote yes: they take neither path, and a no would send the warp's other blocks down the slower
  // general path -- flat image regions would pay for it.)
  const bool strictly = lum0 < lum2 && lum2 < lum3 && lum3 < lum1;
  const uint32_t vote_mask = kFullWarp ? 0xffffffffu : __activemask();
  const bool all_regular = __all_sync(vote_mask, constant || strictly);
  // Second chance for the line search, again decided once per warp: blocks whose interpolants are only WEAKLY between
  // the base colours (equal luminances: narrow-range blocks in flat, dark or slowly varying image regions -- most of a
  // real texture).  Candidates that tie with a lower index never win ("first strict minimum"), so they drop out of the
  // sequence; what is left is still 0,[2],[3],1 (or 1,[3],[2],0 when swapped) along the line and the same three
  // crossings classify it once the crossings of the missing candidates are collapsed onto their neighbours' (below).
  const bool up = !swapped;  // reference indices along the ascending line: (0,2,3,1) unswapped, (1,3,2,0) swapped
  bool all_monotone = all_regular;
  if (!all_regular) {  // (uniform branch: the usual warp does not pay for the extra comparisons)
    const bool weakly = lum0 <= lum2 && lum2 <= lum3 && lum3 <= lum1;  // (lum0 < lum1 whenever the block is not constant)
    all_monotone = __all_sync(vote_mask, constant || weakly);
  }
  uint32_t bits;
  if (constant) {
    // The reference swaps red and blue a second time here (dxtc_compressor.cc:360), i.e. it looks up the
    // memory-order colour.
    const uint64_t packed = dxt_const_colour(p0, always4);
    c0 = static_cast<uint32_t>(packed) & 0xffffu;
    c1 = (static_cast<uint32_t>(packed) >> 16) & 0xffffu;
    bits = static_cast<uint32_t>(packed >> 32) * 0x55555555u;
  } else if (all_monotone) {
    // Ascending index sequence along the luminance line: rising 0,2,3,1, falling 1,3,2,0 = the rising one with the low
    // bit of every index flipped (one XOR of the finished word); a tie goes to the smaller index.  Crossing points
    // h1 <= h2 <= h3 (multiples of 8, "crossed iff 8*l >= h"), index changes +2, +1, -2:
    //   index = 2*[v >= h1] + [v >= h2] - 2*[v >= h3]
    const uint32_t a0 = lum0, a1 = lum2, a2 = lum3, a3 = lum1;  // ascending
    uint32_t h1, h2, h3;
    if (all_regular) {
      h1 = ((a0 + a1 + 16u) >> 1) & ~7u;                      // 0->2 / 1->3: larger index, tie stays
      h2 = ((a1 + a2 + (up ? 16u : 8u)) >> 1) & ~7u;          // 2->3 stays on tie, 3->2 moves
      h3 = ((a2 + a3 + 8u) >> 1) & ~7u;                       // 3->1 / 2->0: smaller index, tie moves
    } else {
      // Which of the two inner candidates survive.  Ascending order carries indices (0,2,3,1) when rising and (1,3,2,0)
      // when falling; a candidate is dead when an equal luminance exists under a smaller index.
      const bool inner_tie = a1 == a2;
      const bool dead1 = a1 == a0 || a1 == a3 || (!up && inner_tie);   // index 2 (rising) / 3 (falling)
      const bool dead2 = a2 == a0 || a2 == a3 || (up && inner_tie);    // index 3 (rising) / 2 (falling)
      const uint32_t first_above = !dead1 ? a1 : (!dead2 ? a2 : a3);   // first live candidate above a0
      const uint32_t last_below = !dead2 ? a2 : (!dead1 ? a1 : a0);    // last live candidate below a3
      h1 = ((a0 + first_above + 16u) >> 1) & ~7u;             // into a larger index: a tie stays
      h3 = ((last_below + a3 + 8u) >> 1) & ~7u;               // into a smaller index: a tie moves
      h2 = ((a1 + a2 + (up ? 16u : 8u)) >> 1) & ~7u;
      if (dead1 && dead2) {                                   // only the base colours are left: one crossing, 0->1 / 1->0
        h1 = up ? h1 : h3;
        h3 = h1;
      }
      // The low bit is set for indices 3 and 1: it flips at the inner crossing; without the second inner candidate
      // (index 3 rising, 2 falling) it flips with the band's far edge, without the first one with its near edge.
      if (dead1 || dead2) h2 = dead2 ? h3 : h1;
    }
    // t = relu(min(key + (1 - h), 1)) is 1 once the pixel has passed crossing h -- the key's index field (< 8) cannot
    // carry it over a multiple of 8 -- for both pixels of the pair in one VIADDMNMX.S16x2.RELU, and t * (step << 2k)
    // drops the index change at the pair's bit position of both lanes in one IMAD with an immediate multiplier.  All
    // arithmetic is modulo 2^32 and the final fields are indices 0..3, so the order of the additions does not matter.
    // 1 - h in both lanes as ONE multiply-add: (65537 - h) * 65537 = 0x00020001 - h * 65537 (mod 2^32), and 65537 - h fits
    // a lane because every crossing is at least 8 (h1, h3 round up from a sum >= 8; h2 is replaced when its candidates are dead).
    const uint32_t n1 = h1 * 0xfffeffffu + 0x00020001u, n2 = h2 * 0xfffeffffu + 0x00020001u, n3 = h3 * 0xfffeffffu + 0x00020001u;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc += __viaddmin_s16x2_relu(pk[k], n1, 0x00010001u) * (2u << (2 * k));
      acc += __viaddmin_s16x2_relu(pk[k], n2, 0x00010001u) * (1u << (2 * k));
      acc -= __viaddmin_s16x2_relu(pk[k], n3, 0x00010001u) * (2u << (2 * k));
    }
    bits = acc ^ (swapped ? 0x55555555u : 0u);
  } else {
    float acc0, cross[3], step[3];
    {
      // General case (crossed candidates): sort the candidates as keys 8*L_c + c, c = the reference's index; per pixel,
      // each of the three crossings is one saturating float add (1.0 if crossed, else 0.0) and one float multiply-add
      // that accumulates the index change -- exact, since every value is an integer below 2^24.
      uint32_t s0 = swapped ? lum1 : lum0, s1 = (swapped ? lum0 : lum1) + 1u, s2 = (swapped ? lum3 : lum2) + 2u,
               s3 = (swapped ? lum2 : lum3) + 3u;
      sort2(s0, s1); sort2(s2, s3); sort2(s0, s2); sort2(s1, s3); sort2(s1, s2);
      const uint32_t sorted[4] = {s0, s1, s2, s3};
      uint32_t rep = s0;                                  // lowest-index candidate of the current luminance
      acc0 = __uint_as_float(kDxtLumBias + (s0 & 3u));    // 2^23 + index of the lowest candidate
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint32_t b = sorted[j + 1];
        const bool same_lum = (b - rep) < 4u;             // keys differ only in the index bits
        const uint32_t cb = b & 3u, cr = rep & 3u;
        // pixel key v = 8*l (+ lane index < 8) crosses iff v >= h, h = 8 * ceil((L_a + L_b + (cb < cr ? 0 : 1)) / 2)
        const uint32_t h = ((rep + b + (cb < cr ? 8u : 16u)) >> 1) & ~7u;
        cross[j] = __uint_as_float(same_lum ? 0x4b7fffffu : kDxtLumBias + h - 1u);
        step[j] = static_cast<float>((cb - cr) & 3u);  // irrelevant when same_lum: that crossing never fires
        rep = same_lum ? rep : b;
      }
    }
    bits = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      // 2^23 + key of pixel i, rebuilt from the packed 16-bit keys (one byte permute) so that the sixteen key words need
      // not stay in registers across the vote just for this path
      const float v = __uint_as_float(__byte_perm(pk[i & 7], kDxtLumBias, i < 8 ? 0x7610 : 0x7632));
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
  const uint32_t w8 = dxt_lum_weights(swap_rb);
  uint32_t kf[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i] = __dp4a(px[i], w8, dxt_key_seed(i));
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
    kf[4 * y + 0] = __dp4a(rows[y][0], wa, dxt_key_seed(4 * y + 0));
    kf[4 * y + 1] = __dp4a(rows[y][1], wb_hi, __dp4a(rows[y][0], wb_lo, dxt_key_seed(4 * y + 1)));
    kf[4 * y + 2] = __dp4a(rows[y][2], wc_hi, __dp4a(rows[y][1], wc_lo, dxt_key_seed(4 * y + 2)));
    kf[4 * y + 3] = __dp4a(rows[y][2], wd, dxt_key_seed(4 * y + 3));
  }
  return dxt1_encode_from_keys<kFullWarp>(kf, swap_rb, always4, fetch, release);
}

// ---------------------------------------------------------------------------------------------------------
// DXT5 alpha block
// ---------------------------------------------------------------------------------------------------------

// Crossing-point table, 512 entries x 16 words; layout and derivation in tools/gen_dxt5_alpha_table.py, which also
// verifies the table-driven search against the reference's direct search for every (a0, a1, alpha).
__device__ __align__(16) const uint32_t g_dxt5_alpha_table[512 * 16] = {
#include "dxt5_alpha_table.inc"
};
constexpr int kDxt5AlphaTableBytes = 512 * 64;

// Encodes the DXT5 alpha half from the top byte of each pixel (ComputeBaseAlphas, dxtc_compressor.cc:374-424;
// ComputeAlphaBits :427-479; bit layout Dxt5AlphaBits :103-158).  Returns the 8 output bytes as two words.
// `table` points at the crossing-point table (global memory, L1-resident).
//
// Everything runs two pixels per instruction on 16-bit integer lanes (VIADD.16x2, VIMNMX3.U16x2,
// VIADDMNMX.S16x2.RELU in SASS): pixel pair (i, i+8) lives in one register as (a_{i+8} << 16) | a_i.
//   statistics  a - 1 (mod 2^16) sends 0 to 0xffff: the unsigned minimum of those keys is the smallest alpha that is
//               not 0, and the keys' high bytes add up to 255 * (number of zeros) in one IDP.4A per register;
//               a + 0xff01 (mod 2^16) sends 255 to 0: the maximum is the largest alpha that is not 255 and the high
//               bytes add up to 255 * (number of alphas that are not 255).
//   indices     nearest-candidate search as seven crossings walked in ascending order (see the table generator):
//               t = relu(min(a - a0 + c_s, 1)) is 1 once the pixel has passed crossing s (one VIADDMNMX, integer
//               pipe), acc += t * step_s (one IMAD, FMA pipe).  Steps are the true signed index differences, so every
//               partial sum is an index 0..7: no masking, and the two pixel pairs of an accumulator keep their 3-bit
//               fields apart without borrows.
// Round 1 ran this on packed fp16 (HFMA2.SAT + HFMA2 per crossing, both on the half-rate FMA pipe, 322 instructions
// per block); this form needs about 240 and splits them between the two pipes.
// Part 1: the two endpoint alphas, packed a0 | a1 << 8 (ComputeBaseAlphas).
// The sixteen alphas as eight lane pairs: x[i] = (alpha of pixel i + 8) << 16 | alpha of pixel i.
__device__ __forceinline__ void dxt5_alpha_lanes(const uint32_t (&px)[16], uint32_t (&x)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = __byte_perm(px[i], px[i + 8], 0x7733) & 0x00ff00ffu;
}

// The two lane-wise addends of the statistics keys live in constant memory, NOT in the instruction stream: with them as
// immediates, CUDA 12.9's ptxas miscompiles one of the eight additions in some instantiations (the 48-register ring
// build, the generic DXT5 kernel: "VIADD.16x2 R17,R4,0x0 / VIADD.16x2 R15,R4,0xff01ff01 / PRMT R19,R17,0x7610,R15" --
// the low lane of one key loses its addend, and the block's alpha endpoint is off whenever that pixel holds the
// extreme; found by the bench's parity flag and the GPU suite, profiles/r02b_driver_ab.txt).  Constant memory is
// writable from the host, so ptxas cannot split these values into lanes.
__constant__ uint32_t c_dxt5_minus_one_lanes = 0xffffffffu, c_dxt5_plus_ff01_lanes = 0xff01ff01u;

__device__ __forceinline__ uint32_t dxt5_alpha_endpoints(const uint32_t (&x)[8]) {
  // ---- statistics
  uint32_t fk[8], gk[8], zeros255 = 0, not255 = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    fk[i] = __vadd2(x[i], c_dxt5_minus_one_lanes);   // a - 1; a == 0 -> 0xffff
    gk[i] = __vadd2(x[i], c_dxt5_plus_ff01_lanes);   // a + 0xff01; a == 255 -> 0
    zeros255 = __dp4a(fk[i], 0x01000100u, zeros255);  // += high bytes: 255 per zero alpha
    not255 = __dp4a(gk[i], 0x01000100u, not255);      // += 255 per alpha that is not 255
  }
  uint32_t fmin = __vimin3_u16x2(fk[0], fk[1], fk[2]);
  fmin = __vimin3_u16x2(fmin, fk[3], fk[4]);
  fmin = __vimin3_u16x2(fmin, fk[5], fk[6]);
  fmin = __vminu2(fmin, fk[7]);
  uint32_t gmax = __vimax3_u16x2(gk[0], gk[1], gk[2]);
  gmax = __vimax3_u16x2(gmax, gk[3], gk[4]);
  gmax = __vimax3_u16x2(gmax, gk[5], gk[6]);
  gmax = __vmaxu2(gmax, gk[7]);
  // smallest alpha that is neither 0 nor 255 (255 when there is none), largest such alpha (0 when there is none)
  int lo = static_cast<int>(min(min(fmin & 0xffffu, fmin >> 16), 254u)) + 1;
  int hi = max(static_cast<int>(max(gmax & 0xffffu, gmax >> 16)) - 0xff01, 0);
  if (lo > hi) {  // every alpha is 0 or 255
    lo = 0;
    hi = 255;
  }
  uint32_t a0, a1;
  if (zeros255 > 255u || not255 < 14u * 255u + 1u) {  // more than one fully transparent or fully opaque pixel
    a0 = lo;
    a1 = hi;
  } else {
    if (zeros255 != 0u) lo = 0;
    if (not255 != 16u * 255u) hi = 255;
    a0 = hi;
    a1 = lo;
  }
  return a0 | (a1 << 8);
}

// Part 2: the sixteen 3-bit indices for endpoints a0 | a1 << 8 (ComputeAlphaBits).  Returns the 8 output bytes.
template <bool kFullWarp = false>
__device__ __forceinline__ uint2 dxt5_alpha_indices(const uint32_t (&x)[8], uint32_t endpoints, const uint4 *table) {
  const uint32_t a0 = endpoints & 255u, a1 = endpoints >> 8;
  // Usual case, decided once per warp (uniform branch): 8-alpha mode and no two candidates equal.  Ascending from a1 the
  // line then carries the indices 1,7,6,5,4,3,2,0 -- start 1, index changes +6, -1 x5, -2 (tools/gen_dxt5_alpha_table.py
  // derives them; tests/test_host_math.py checks that every table entry with D >= 7 says the same) -- so only the seven
  // thresholds come from the table (two of the entry's four 16-byte words), the five -1 crossings are summed with two
  // three-input adds before they meet the accumulator, and the field position of a pixel pair is part of the immediate
  // multiplier: 13 instructions per pixel pair instead of 15.5, no 6-alpha patch-up, no step words.
  if (__all_sync(kFullWarp ? 0xffffffffu : __activemask(), a0 >= a1 + 7u)) {
    const uint4 *e = table + 4u * (256u + a0 - a1);
    const uint4 e0 = e[0], e1 = e[1];
    // -a0 in both lanes as one multiply-add: (65536 - a0) * 65537 = 0x00010000 - a0 * 65537 (mod 2^32); a0 >= 7 here.
    // It goes into the seven thresholds (relative to a0 in the table) rather than into the eight alpha lane pairs.
    const uint32_t minus_a0 = a0 * 0xfffeffffu + 0x00010000u;
    const uint32_t c0 = __vadd2(e0.x, minus_a0), c1 = __vadd2(e0.y, minus_a0), c2 = __vadd2(e0.z, minus_a0),
                   c3 = __vadd2(e0.w, minus_a0), c4 = __vadd2(e1.x, minus_a0), c5 = __vadd2(e1.y, minus_a0),
                   c6 = __vadd2(e1.z, minus_a0);
    uint32_t acc_a = 0x02490249u, acc_b = 0x02490249u;  // index 1 in four 3-bit fields of both lanes
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const uint32_t d = x[p];
      const uint32_t t0 = __viaddmin_s16x2_relu(d, c0, 0x00010001u), t1 = __viaddmin_s16x2_relu(d, c1, 0x00010001u),
                     t2 = __viaddmin_s16x2_relu(d, c2, 0x00010001u), t3 = __viaddmin_s16x2_relu(d, c3, 0x00010001u),
                     t4 = __viaddmin_s16x2_relu(d, c4, 0x00010001u), t5 = __viaddmin_s16x2_relu(d, c5, 0x00010001u),
                     t6 = __viaddmin_s16x2_relu(d, c6, 0x00010001u);
      const uint32_t ones = t1 + t2 + t3 + t4 + t5;  // crossings that lower the index by one
      const uint32_t f = 1u << (3 * (p & 3));        // this pair's field in both lanes
      uint32_t &acc = p < 4 ? acc_a : acc_b;
      // (modulo 2^32; the finished fields are indices 0..7, so intermediate borrows between fields cancel)
      acc += t0 * (6u * f);
      acc -= ones * f;
      acc -= t6 * (2u * f);
    }
    const uint32_t word0 = a0 | (a1 << 8) | ((acc_a & 0xfffu) << 16) | (acc_b << 28);
    const uint32_t word1 = ((acc_b & 0xfffu) >> 4) | ((acc_a >> 16) << 8) | ((acc_b >> 16) << 20);
    return make_uint2(word0, word1);
  }

  // ---- crossings for this (mode, |a0 - a1|)
  const bool six = a0 <= a1;  // 6-alpha mode: candidates 0 and 255 are explicit
  const uint32_t dist = __usad(a0, a1, 0u);
  const uint4 *e = table + 4u * ((six ? 0u : 256u) + dist);
  const uint4 e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3];
  uint32_t c[7] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z};
  uint32_t step[7] = {e2.x, e2.y, e2.z, e2.w, e3.x, e3.y, e3.z};
  uint32_t start = e1.w;
  if (six) {
    // crossings against the explicit candidates: 0 (index 6, or the line's own index 0 when a0 is itself 0) below the
    // line and 255 (index 7, or the line's last index when a1 is itself 255) above it
    const uint32_t a0_zero = a0 == 0u ? 1u : 0u, last = e3.w;
    start = a0_zero ? 0u : 6u;
    c[0] = ((1u + a0 - ((a0 + a0_zero + 1u) >> 1)) & 0xffffu) * 0x10001u;
    step[0] = 0u - start;
    c[6] = ((1u + a0 - ((a1 + 257u) >> 1)) & 0xffffu) * 0x10001u;
    step[6] = a1 == 255u ? 0u : 7u - last;
  }
  const uint32_t minus_a0 = ((0u - a0) & 0xffffu) * 0x10001u;
  start *= 0x10001u;

  // ---- indices: one accumulator per two pixel pairs; word w's 3-bit indices sit at bit 3*(w & 1) of each lane
  uint32_t acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t a = start;
#pragma unroll
    for (int half = 1; half >= 0; --half) {
      const uint32_t d = __vadd2(x[2 * k + half], minus_a0);  // alpha - a0, two's complement per lane
#pragma unroll
      for (int s = 0; s < 7; ++s) a += __viaddmin_s16x2_relu(d, c[s], 0x00010001u) * step[s];
      if (half == 1) a = a * 8u + start;
    }
    acc[k] = a;
  }
  // lanes of acc[k]: low = pixels 2k, 2k+1 (6 bits), high = pixels 2k+8, 2k+9
  const uint32_t acc_a = acc[1] * 64u + acc[0];  // low lane: pixels 0..3 (12 bits), high lane: pixels 8..11
  const uint32_t acc_b = acc[3] * 64u + acc[2];  // low lane: pixels 4..7, high lane: pixels 12..15
  // pixel n's code goes to bit 16 + 3n of the 64-bit block half
  const uint32_t word0 = a0 | (a1 << 8) | ((acc_a & 0xfffu) << 16) | (acc_b << 28);
  const uint32_t word1 = ((acc_b & 0xfffu) >> 4) | ((acc_a >> 16) << 8) | ((acc_b >> 16) << 20);
  return make_uint2(word0, word1);
}

template <bool kFullWarp = false>
__device__ __forceinline__ uint2 dxt5_encode_alpha_lanes(const uint32_t (&x)[8], bool one_pixel, const uint4 *table) {
  // A window entirely outside the image (generic driver only): both endpoints = that alpha, all indices 0.  It still goes
  // through the search below, whose warp vote every lane must reach.
  const uint2 r = dxt5_alpha_indices<kFullWarp>(x, dxt5_alpha_endpoints(x), table);
  if (one_pixel) {
    const uint32_t a = x[0] & 255u;
    return make_uint2(a | (a << 8), 0u);
  }
  return r;
}
template <bool kFullWarp = false>
__device__ __forceinline__ uint2 dxt5_encode_alpha(const uint32_t (&px)[16], bool one_pixel, const uint4 *table) {
  uint32_t x[8];
  dxt5_alpha_lanes(px, x);
  return dxt5_encode_alpha_lanes<kFullWarp>(x, one_pixel, table);
}

}  // namespace icb
