// cuda_emulation.h -- TEST INFRASTRUCTURE ONLY.  Host definitions of the CUDA qualifiers, vector types, built-in
// variables and integer / packed-half intrinsics that image_compression_b200/csrc/*.cuh use, so that the DEVICE
// encoders (the very source the GPU runs) can be compiled with g++ and stepped through on the CPU, one emulated thread
// at a time, by tests/test_hostemu.py.  Nothing here is compiled into, linked with or loaded by the product library:
// the product has no CPU path (image_compression_b200/csrc/icb_api.cu fails when no CUDA device is present).
//
// Each intrinsic follows the CUDA Math API / PTX ISA definition of the instruction it stands for; the ones with
// corner cases (byte permute selectors, packed-half rounding and saturation, funnel shift operand order) are
// exercised on their own in hostemu.cc:emu_self_check().
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define ICB_HOST_EMULATION 1

// ---- qualifiers
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static const  // (constant memory: plain read-only data here)
#define __shared__ static  // threads are emulated one after another; per-thread rows of a shared array stay private

// ---- vector types
struct __attribute__((aligned(8))) uint2 {
  uint32_t x, y;
};
struct __attribute__((aligned(16))) uint4 {
  uint32_t x, y, z, w;
};
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

// ---- built-in variables (set by the harness before every emulated thread)
struct EmuDim3 {
  uint32_t x = 1, y = 1, z = 1;
};
inline EmuDim3 blockIdx, threadIdx, blockDim, gridDim;

// ---- warp votes.  One thread is emulated at a time, so a vote cannot see the other lanes; the harness chooses what
// the rest of the warp "said": kEmuVoteAgree -- every other lane voted like this one (the vote returns the lane's own
// predicate), kEmuVoteNo -- some other lane voted no (every vote returns false), kEmuVoteFirstNo -- some other lane
// voted no in the encoder's FIRST vote and agreed in its second.  Encoders whose result must not depend on the votes
// (dxt1_encode_from_keys: strict line search / weakly monotone line search / general search, picked by two
// warp-uniform votes, the second cast only when the first fails) are run under all three.  In the third mode every
// encode casts exactly two votes, so "first" is simply every even-numbered vote since emu_set_vote().
enum EmuVote { kEmuVoteAgree = 0, kEmuVoteNo = 1, kEmuVoteFirstNo = 2 };
inline EmuVote g_emu_vote = kEmuVoteAgree;
inline int g_emu_votes_cast = 0;
static inline uint32_t __activemask() { return 0xffffffffu; }
static inline bool __all_sync(uint32_t, bool pred) {
  const int nth = g_emu_votes_cast++;
  if (g_emu_vote == kEmuVoteAgree) return pred;
  if (g_emu_vote == kEmuVoteNo) return false;
  return (nth & 1) ? pred : false;
}

// Warp-wide minimum (REDUX): like the votes, the harness decides what the other lanes contribute -- nothing that
// lowers this lane's value (kEmuVoteAgree), or a zero (the other modes), which sends ETC1's codeword search (whose
// operand is a margin + 128) down its general form for every codeword.
// g_emu_min_cap: a value some other lane holds (emu_set_vote(1000 + m)), for minima that must give the same result
// whatever the rest of the warp contributes -- ETC1 is run with every margin that changes the number of line-form codewords.
inline uint32_t g_emu_min_cap = 0xffffffffu;
static inline uint32_t __reduce_min_sync(uint32_t, uint32_t v) { return g_emu_vote == kEmuVoteAgree ? (v < g_emu_min_cap ? v : g_emu_min_cap) : 0u; }

// ---- scalar helpers
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline uint64_t max(uint64_t a, uint64_t b) { return a > b ? a : b; }
static inline float __uint_as_float(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline uint32_t __float_as_uint(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
static inline float __saturatef(float x) { return x != x ? 0.0f : x < 0.0f ? 0.0f : x > 1.0f ? 1.0f : x; }
template <typename T>
static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
static inline uint32_t __usad(uint32_t a, uint32_t b, uint32_t c) { return (a > b ? a - b : b - a) + c; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
// PTX shf.r.wrap: the low 32 bits of (hi:lo) >> (shift & 31).
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
  return static_cast<uint32_t>(((static_cast<uint64_t>(hi) << 32) | lo) >> (shift & 31u));
}

// PTX prmt (default mode) as exposed by __byte_perm: result byte i = byte (selector nibble i) of the eight bytes
// {x, y}.  The Math API defines only the low three bits of a nibble; the device headers never set the fourth
// (PTX would replicate the byte's sign bit), and the emulation refuses a selector that does.
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t pool = (static_cast<uint64_t>(y) << 32) | x;
  uint32_t out = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t sel = (s >> (4 * i)) & 0xfu;
    if (sel & 8u) std::abort();
    out |= static_cast<uint32_t>((pool >> (8 * sel)) & 0xffu) << (8 * i);
  }
  return out;
}

// dp4a: four byte products plus accumulator; the unsigned overload treats all bytes as unsigned, the signed one as
// signed (the device headers use only the unsigned form -- a mixed call would be ambiguous here as it is in CUDA).
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {
  for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
  return c;
}
static inline int __dp4a(int a, int b, int c) {
  for (int i = 0; i < 4; ++i) c += static_cast<int8_t>(a >> (8 * i)) * static_cast<int8_t>(b >> (8 * i));
  return c;
}

// ---- SIMD-in-a-word integer instructions
static inline uint32_t emu_lanes16(uint32_t a, uint32_t b, uint32_t (*f)(uint32_t, uint32_t)) {
  return (f(a & 0xffffu, b & 0xffffu) & 0xffffu) | (f(a >> 16, b >> 16) << 16);
}
static inline uint32_t emu_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t emu_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t emu_add(uint32_t a, uint32_t b) { return a + b; }
static inline uint32_t __vadd2(uint32_t a, uint32_t b) { return emu_lanes16(a, b, emu_add); }  // each lane wraps to 16 bits
static inline uint32_t __vminu2(uint32_t a, uint32_t b) { return emu_lanes16(a, b, emu_umin); }
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b) { return emu_lanes16(a, b, emu_umax); }
static inline uint32_t __vimin3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vminu2(__vminu2(a, b), c); }
static inline uint32_t __vimax3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxu2(__vmaxu2(a, b), c); }
static inline uint32_t __vimin3_u32(uint32_t a, uint32_t b, uint32_t c) { return emu_umin(emu_umin(a, b), c); }
static inline uint32_t __vimax3_u32(uint32_t a, uint32_t b, uint32_t c) { return emu_umax(emu_umax(a, b), c); }
// max(min(a + b, c), 0)
static inline int __viaddmin_s32_relu(int a, int b, int c) {
  const int s = static_cast<int>(static_cast<uint32_t>(a) + static_cast<uint32_t>(b));
  const int m = s < c ? s : c;
  return m > 0 ? m : 0;
}
// the same on two signed 16-bit lanes (the sum wraps to 16 bits)
static inline uint32_t __viaddmin_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t out = 0;
  for (int k = 0; k < 2; ++k) {
    const int16_t s = static_cast<int16_t>(static_cast<uint16_t>((a >> (16 * k)) + (b >> (16 * k))));
    const int16_t lim = static_cast<int16_t>(c >> (16 * k));
    const int16_t m = s < lim ? s : lim;
    out |= static_cast<uint32_t>(static_cast<uint16_t>(m > 0 ? m : 0)) << (16 * k);
  }
  return out;
}
// max(max(a + b, c), 0) on two signed 16-bit lanes (the sum wraps to 16 bits)
static inline uint32_t __viaddmax_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t out = 0;
  for (int k = 0; k < 2; ++k) {
    const int16_t s = static_cast<int16_t>(static_cast<uint16_t>((a >> (16 * k)) + (b >> (16 * k))));
    const int16_t lim = static_cast<int16_t>(c >> (16 * k));
    const int16_t m = s > lim ? s : lim;
    out |= static_cast<uint32_t>(static_cast<uint16_t>(m > 0 ? m : 0)) << (16 * k);
  }
  return out;
}
// dp2a.lo.u32.u32: the two 16-bit halves of a times the two LOW bytes of b, plus c
static inline uint32_t __dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { return c + (a & 0xffffu) * (b & 0xffu) + (a >> 16) * ((b >> 8) & 0xffu); }
static inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b) {
  uint32_t out = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = (a >> (8 * i)) & 0xffu, y = (b >> (8 * i)) & 0xffu;
    out |= (x > y ? x - y : y - x) << (8 * i);
  }
  return out;
}
static inline uint32_t __vsadu4(uint32_t a, uint32_t b) {
  const uint32_t d = __vabsdiffu4(a, b);
  return (d & 0xffu) + ((d >> 8) & 0xffu) + ((d >> 16) & 0xffu) + (d >> 24);
}

// ---- packed half precision (fma.rn[.sat].f16x2, add.rn[.sat].f16x2): every lane is computed exactly in double
// (an fp16 product has 22 significant bits; the sums these headers form stay far inside 53) and rounded ONCE to
// half, round-to-nearest-even, subnormals kept -- what the hardware instruction does.
namespace icb_emu {

inline double half_to_double(uint16_t h) {
  const int sign = h >> 15, exp = (h >> 10) & 31, man = h & 1023;
  double v;
  if (exp == 31)
    v = man ? NAN : INFINITY;
  else if (exp == 0)
    v = std::ldexp(static_cast<double>(man), -24);
  else
    v = std::ldexp(static_cast<double>(man + 1024), exp - 25);
  return sign ? -v : v;
}

inline uint16_t double_to_half(double v) {
  if (v != v) return 0x7fffu;
  const uint16_t sign = std::signbit(v) ? 0x8000u : 0u;
  const double a = std::fabs(v);
  if (a >= 65520.0) return sign | 0x7c00u;  // rounds to infinity (65520 is halfway to 2^16, ties to even = 2^16)
  if (a < std::ldexp(1.0, -14)) return sign | static_cast<uint16_t>(std::nearbyint(std::ldexp(a, 24)));  // subnormal (or 2^-14)
  int e;
  std::frexp(a, &e);  // a = m * 2^e, m in [0.5, 1)
  int exp = e - 1;    // a in [2^exp, 2^(exp+1))
  double q = std::nearbyint(std::ldexp(a, 10 - exp));  // in [1024, 2048], ties to even (default rounding mode)
  if (q == 2048.0) {
    q = 1024.0;
    ++exp;
  }
  return sign | static_cast<uint16_t>(((exp + 15) << 10) + (static_cast<int>(q) - 1024));
}

inline uint16_t half_finish(double v, bool sat) {
  if (sat) v = v != v ? 0.0 : v < 0.0 ? 0.0 : v > 1.0 ? 1.0 : v;
  return double_to_half(v);
}

inline uint32_t fma_f16x2(uint32_t a, uint32_t b, uint32_t c, bool sat) {
  uint32_t out = 0;
  for (int k = 0; k < 2; ++k) {
    const double r = half_to_double(a >> (16 * k)) * half_to_double(b >> (16 * k)) + half_to_double(c >> (16 * k));
    out |= static_cast<uint32_t>(half_finish(r, sat)) << (16 * k);
  }
  return out;
}

inline uint32_t add_f16x2(uint32_t a, uint32_t b, bool sat) {
  uint32_t out = 0;
  for (int k = 0; k < 2; ++k)
    out |= static_cast<uint32_t>(half_finish(half_to_double(a >> (16 * k)) + half_to_double(b >> (16 * k)), sat)) << (16 * k);
  return out;
}

}  // namespace icb_emu
