// pipe_rates.cu -- issue-rate microbenchmark for the integer / float / SIMD instructions the encoders lean on.
// For each op: 8 independent dependency chains per thread, 1024 threads per CTA, one CTA per SM; reports
// warp-instructions per clock per SM (4 = every scheduler issues every cycle).  Pairs of ops are also timed
// interleaved: if two ops share a pipe the pair runs at the sum of their times, otherwise at the max.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 512

enum Op { IMAD, IADD, LOP, SHF, PRMT, VABSDIFF, VABSDIFF4, VIMNMX, VIMNMX3, VIMNMX16X2, VIADDMNMX, IDP, FFMA, FADDSAT, HFMA2, HADD2SAT,
          IMADHI, ISETPSEL, LEA, NOPS };
static const char *kNames[] = {"IMAD", "IADD3", "LOP3", "SHF.funnel", "PRMT", "VABSDIFF.U32", "VABSDIFF4", "VIMNMX", "VIMNMX3",
                               "VIMNMX.U16x2", "VIADDMNMX.RELU", "IDP.4A", "FFMA", "FADD.SAT", "HFMA2", "HADD2.SAT", "IMAD.HI",
                               "ISETP+SEL", "LEA"};

template <int OP>
__device__ __forceinline__ uint32_t apply(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  if (OP == IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == IADD) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else if (OP == LOP) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %1, %2, 3;" : "=r"(d) : "r"(a), "r"(b));
  else if (OP == PRMT) asm volatile("prmt.b32 %0, %1, %2, 0x3715;" : "=r"(d) : "r"(a), "r"(b));
  else if (OP == VABSDIFF) d = __usad(a, b, c);
  else if (OP == VABSDIFF4) d = __vabsdiffu4(a, b);
  else if (OP == VIMNMX) d = min(a, b);
  else if (OP == VIMNMX3) d = __vimin3_u32(a, b, c);
  else if (OP == VIMNMX16X2) d = __vminu2(a, b);
  else if (OP == VIADDMNMX) d = (uint32_t)__viaddmin_s32_relu((int)a, (int)b, 255);
  else if (OP == IDP) d = __dp4a(a, b, c);
  else if (OP == FFMA) d = __float_as_uint(fmaf(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c)));
  else if (OP == FADDSAT) d = __float_as_uint(__saturatef(__uint_as_float(a) + __uint_as_float(b)));
  else if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == HADD2SAT) asm volatile("add.sat.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else if (OP == IMADHI) d = __umulhi(a, b);
  else if (OP == ISETPSEL) d = a < b ? c : a;
  else if (OP == LEA) d = (a << 3) + b;
  else d = a;
  return d;
}

template <int OP1, int OP2>
__global__ void __launch_bounds__(1024) rate_kernel(uint32_t *out, long long *cycles, uint32_t seed) {
  uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) {
    x[k] = seed * (k + 3) + threadIdx.x;
    y[k] = seed * (k + 11) ^ threadIdx.x;
  }
  const uint32_t b = seed | 1, c = seed + 7;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) {
      x[k] = apply<OP1>(x[k], b, c);
      if (OP2 != NOPS) y[k] = apply<OP2>(y[k], c, b);
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) acc ^= x[k] ^ y[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP1, int OP2>
double run(uint32_t *out, long long *cycles, int sms) {
  rate_kernel<OP1, OP2><<<sms, 1024>>>(out, cycles, 12345u);
  rate_kernel<OP1, OP2><<<sms, 1024>>>(out, cycles, 12345u);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)h[i];
  avg /= sms;
  const double warp_instr = 32.0 * ITERS * CHAINS * (OP2 == NOPS ? 1 : 2);  // 32 warps per CTA
  return warp_instr / avg;
}

#define SINGLE(OP) printf("%-16s %6.2f warp-instr/clk/SM\n", kNames[OP], run<OP, NOPS>(out, cycles, sms));
#define PAIR(A, B) printf("%-16s + %-16s %6.2f warp-instr/clk/SM (both)\n", kNames[A], kNames[B], run<A, B>(out, cycles, sms));

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  uint32_t *out;
  long long *cycles;
  cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
  cudaMalloc(&cycles, sizeof(long long) * sms);
  printf("%s, %d SMs\n", prop.name, sms);
  SINGLE(IMAD) SINGLE(IADD) SINGLE(LOP) SINGLE(SHF) SINGLE(PRMT) SINGLE(VABSDIFF) SINGLE(VABSDIFF4) SINGLE(VIMNMX) SINGLE(VIMNMX3)
  SINGLE(VIMNMX16X2) SINGLE(VIADDMNMX) SINGLE(IDP) SINGLE(FFMA) SINGLE(FADDSAT) SINGLE(HFMA2) SINGLE(HADD2SAT) SINGLE(IMADHI)
  SINGLE(ISETPSEL) SINGLE(LEA)
  PAIR(LOP, IMAD) PAIR(LOP, IDP) PAIR(IMAD, IDP) PAIR(LOP, FFMA) PAIR(IMAD, FFMA) PAIR(IDP, FFMA) PAIR(LOP, HFMA2) PAIR(IMAD, HFMA2)
  PAIR(FFMA, HFMA2) PAIR(LOP, VABSDIFF) PAIR(LOP, VIMNMX3) PAIR(IMAD, VABSDIFF) PAIR(LOP, SHF) PAIR(LOP, PRMT) PAIR(IMAD, IMADHI)
  PAIR(FFMA, FADDSAT) PAIR(HFMA2, HADD2SAT) PAIR(IDP, HFMA2) PAIR(LOP, VABSDIFF4) PAIR(IMAD, VABSDIFF4) PAIR(LOP, VIADDMNMX)
  return 0;
}
