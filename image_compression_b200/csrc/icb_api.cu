// icb_api.cu -- the extern "C" boundary (include/icb200.h): argument checking, path selection (TMA fast path vs
// generic driver), launches, and the host-buffer pipeline.  No CPU fallback exists: without a CUDA device every
// compute entry point fails with ICB_ERR_CUDA.
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/icb200.h"
#include "block4x4_kernels.cuh"
#include "blockops_kernels.cuh"
#include "decode4x4_kernels.cuh"
#include "pvrtc_kernels.cuh"

namespace {

thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_tma_mode{-1};

int fail(int status, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return status;
}

#define ICB_CUDA(call)                                                                           \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) return fail(ICB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));   \
  } while (0)

// ---- stream-ordered scratch memory -------------------------------------------------------------------------
// Scratch the library allocates on a caller's stream (PVRTC without a caller-provided buffer) comes from the library's
// own memory pool, which keeps freed blocks instead of returning them to the driver at the next synchronisation: the
// default pool's release threshold of zero turns every call after a sync into a fresh device allocation.
std::mutex g_scratch_mu;
cudaMemPool_t g_scratch_pools[64] = {};

cudaMemPool_t scratch_pool_if_created(int dev) {
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  return (dev >= 0 && dev < 64) ? g_scratch_pools[dev] : nullptr;
}

int scratch_pool(cudaMemPool_t *out) {
  std::mutex &mu = g_scratch_mu;
  cudaMemPool_t *pools = g_scratch_pools;
  int dev = 0;
  ICB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(ICB_ERR_CUDA, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    ICB_CUDA(cudaMemPoolCreate(&pools[dev], &props));
    uint64_t keep = ~0ull;
    ICB_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
  }
  *out = pools[dev];
  return ICB_OK;
}

// ---- per-device facts ------------------------------------------------------------------------------------

struct DeviceInfo {
  int sm_count = 0;
  bool ok = false;
};

int device_info(DeviceInfo *out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  int dev = 0;
  ICB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(ICB_ERR_CUDA, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (!cache[dev].ok) {
    cudaDeviceProp prop;
    ICB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
      return fail(ICB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
    cache[dev].sm_count = prop.multiProcessorCount;
    cache[dev].ok = true;
  }
  *out = cache[dev];
  return ICB_OK;
}

// cuTensorMapEncodeTiled comes from the driver; resolve it through the runtime so we need not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// ---- 4x4 codecs ------------------------------------------------------------------------------------------

using icb::Encode4x4Params;

template <int kCodec, int kNcomp>
int launch_generic(const Encode4x4Params &p, int sm_count, cudaStream_t stream) {
  const uint64_t total = static_cast<uint64_t>(p.row1 - p.row0) * (p.col1 - p.col0);
  if (total == 0) return ICB_OK;
  const uint64_t want = (total + 127) / 128;
  const uint32_t grid = static_cast<uint32_t>(want < static_cast<uint64_t>(sm_count) * 32 ? want : sm_count * 32);
  icb::encode4x4_generic_kernel<kCodec, kNcomp><<<grid, 128, 0, stream>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  ICB_CUDA(cudaGetLastError());
  return ICB_OK;
}

template <int kCodec, int kNcomp, bool kSwapRb>
int launch_tma_typed(const Encode4x4Params &p, int sm_count, cudaStream_t stream) {
  using Shape = icb::TileShape<kCodec, kNcomp>;
  if (static_cast<uint64_t>(p.grid_cols) * icb::CodecTraits<kCodec>::kBlockBytes > 0xffffffffull)
    return fail(ICB_ERR_INVALID, "image too wide");
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) return fail(ICB_ERR_CUDA, "cuTensorMapEncodeTiled not available from this driver");
  // The image rows as 32-bit words: width words for RGBA8888, 3*width/4 for RGB888 (width % 4 == 0 here).
  CUtensorMap map;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(p.width) * kNcomp / 4, p.height};
  const cuuint64_t strides[1] = {p.pitch};
  const cuuint32_t box[2] = {Shape::kRowWords, Shape::kRows};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t *>(p.src), dims, strides, box,
                            elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ICB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));

  // Driver and ring depth.  DXT5 and DXT1 from four-byte pixels run the producer-less ring kernel with two stages (see
  // block4x4_kernels.cuh; for DXT1 from RGBA8 it took over from the producer-warp kernel once the integer-lane index
  // search had shortened the encoder: 48.8-49.1 us against 50.9-51.2 at 8192^2, profiles/r02b_driver_ab.txt); DXT1 from
  // RGB888 (43.0-44.8 against 46.6) and ETC1 run the producer-warp kernel with 3 or 4 stages, whichever gives the most
  // resident CTAs (ties -> deeper ring).  ICB_DRIVER=ring|producer and ICB_TMA_STAGES=2|3|4 override for experiments.
  struct Config {
    void (*kernel)(const CUtensorMap, const Encode4x4Params, uint32_t, uint32_t);
    size_t smem;
    int ctas_per_sm;
    int threads;
  };
  static thread_local Config chosen[64] = {};
  int dev = 0;
  ICB_CUDA(cudaGetDevice(&dev));
  if (chosen[dev].kernel == nullptr) {
    constexpr size_t kStage = Shape::kBytes + 24;  // tile + its barrier / counter words + its output origin
    constexpr size_t kTable = 0;  // (round 1 kept DXT5's crossing table behind the ring; it is read through L1 now)
    const char *driver = getenv("ICB_DRIVER"), *force = getenv("ICB_TMA_STAGES");
    const bool ring = driver ? strcmp(driver, "ring") == 0 : (kCodec == icb::kCodecDxt5 || (kCodec == icb::kCodecDxt1 && kNcomp == 4));
    Config cand[3];
    int n = 0;
    if (ring) {
      cand[n++] = {icb::encode4x4_ring_kernel<kCodec, kNcomp, 2, kSwapRb>, 2 * kStage, 0, Shape::kConsumerThreads};
      cand[n++] = {icb::encode4x4_ring_kernel<kCodec, kNcomp, 3, kSwapRb>, 3 * kStage, 0, Shape::kConsumerThreads};
    } else {
      cand[n++] = {icb::encode4x4_tma_kernel<kCodec, kNcomp, 4, kSwapRb>, 4 * kStage + kTable, 0, Shape::kConsumerThreads + 32};
      cand[n++] = {icb::encode4x4_tma_kernel<kCodec, kNcomp, 3, kSwapRb>, 3 * kStage + kTable, 0, Shape::kConsumerThreads + 32};
    }
    int best = -1;
    for (int c = 0; c < n; ++c) {
      ICB_CUDA(cudaFuncSetAttribute(cand[c].kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cand[c].smem)));
      ICB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cand[c].ctas_per_sm, cand[c].kernel, cand[c].threads, cand[c].smem));
      if (cand[c].ctas_per_sm < 1) cand[c].ctas_per_sm = 1;
      if (force && static_cast<size_t>(atoi(force)) * kStage + (ring ? 0 : kTable) == cand[c].smem) best = c;
    }
    if (best < 0) {
      best = 0;  // first candidate unless a later one fits more CTAs per SM
      for (int c = 1; c < n; ++c)
        if (cand[c].ctas_per_sm > cand[best].ctas_per_sm) best = c;
    }
    chosen[dev] = cand[best];
  }
  const Config cfg = chosen[dev];
  const uint32_t tiles_x = (p.col1 - p.col0 + Shape::kBlocksX - 1) / Shape::kBlocksX;
  const uint32_t tiles_y = (p.row1 - p.row0 + Shape::kBlocksY - 1) / Shape::kBlocksY;
  const uint64_t num_tiles64 = static_cast<uint64_t>(tiles_x) * tiles_y;
  if (num_tiles64 == 0) return ICB_OK;
  if (num_tiles64 > 0x7fffffffull) return fail(ICB_ERR_INVALID, "image too large");
  const uint32_t num_tiles = static_cast<uint32_t>(num_tiles64);
  const uint32_t max_ctas = static_cast<uint32_t>(sm_count * cfg.ctas_per_sm);
  const uint32_t grid = num_tiles < max_ctas ? num_tiles : max_ctas;
  // Launched with programmatic stream serialization: the kernel's launch latency and prologue overlap the tail of
  // whatever kernel precedes it in the stream; it waits (griddepcontrol.wait) for that kernel to complete before it
  // reads or writes global memory, so stream order is preserved for every caller.
  cudaLaunchConfig_t launch = {};
  launch.gridDim = dim3(grid);
  launch.blockDim = dim3(cfg.threads);
  launch.dynamicSmemBytes = cfg.smem;
  launch.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  launch.attrs = attr;
  // Measured (8192^2 / 4096^2, back to back): it gains 4 us on DXT1 and 2.6 us on DXT5, whose tiles are short, and
  // COSTS 17 us on ETC1 -- there the early CTAs of the next launch sit on the SMs spinning at their barriers while the
  // long tail tiles of the current one still need every integer-pipe slot -- so ETC1 launches plainly.
  launch.numAttrs = (kCodec == icb::kCodecEtc1 || getenv("ICB_NO_PDL")) ? 0 : 1;
  ICB_CUDA(cudaLaunchKernelEx(&launch, cfg.kernel, map, p, tiles_x, num_tiles));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ICB_OK;
}

// The red/blue exchange of kBGR / kBGRA sources is a template parameter of the tile kernels (ETC1 ignores it).
template <int kCodec, int kNcomp>
int launch_tma(const Encode4x4Params &p, int sm_count, cudaStream_t stream) {
  if (kCodec != icb::kCodecEtc1 && p.swap_rb) return launch_tma_typed<kCodec, kNcomp, kCodec != icb::kCodecEtc1>(p, sm_count, stream);
  return launch_tma_typed<kCodec, kNcomp, false>(p, sm_count, stream);
}

template <int kCodec, int kNcomp>
int encode4x4_typed(Encode4x4Params p, uint32_t coded_h, uint32_t coded_w, uint32_t r0, uint32_t r1,
                    cudaStream_t stream) {
  DeviceInfo info;
  if (int s = device_info(&info)) return s;
  const uint32_t grid_rows = (coded_h + 3) / 4, grid_cols = (coded_w + 3) / 4;
  if (r1 > grid_rows || r0 > r1) return fail(ICB_ERR_INVALID, "block row range [%u,%u) outside grid of %u rows", r0, r1, grid_rows);
  p.grid_cols = grid_cols;
  // d_dst is the stripe's first block: rebase so kernels can index by absolute block row.
  p.dst -= static_cast<size_t>(r0) * grid_cols * icb::CodecTraits<kCodec>::kBlockBytes;

  const int mode = g_tma_mode.load(std::memory_order_relaxed);
  const bool aligned = (reinterpret_cast<uintptr_t>(p.src) % 16 == 0) && (p.pitch % 16 == 0) &&
                       (kNcomp == 4 || p.width % 4 == 0);
  if (mode == 1 && !aligned) return fail(ICB_ERR_INVALID, "TMA path forced but source base/pitch is not 16-byte aligned");
  const bool use_tma = aligned && mode != 0;

  // Split of the launch's block rows [r0, r1) x all grid columns:
  //   A  blocks whose 4x4 window lies inside the image, if they span at least one tile each way
  //      ............................................. TMA kernel     rows [r0, a_r1) x cols [0, a_c1)
  //   B  right of A, same rows ....................... generic kernel rows [r0, a_r1) x cols [a_c1, grid_cols)
  //   C  everything below A .......................... generic kernel rows [a_r1, r1) x cols [0, grid_cols)
  // (B and C are the ragged image edge, whose windows clamp, and CompressAndPad's pad region.)
  using Shape = icb::TileShape<kCodec, kNcomp>;
  uint32_t a_r1 = r0, a_c1 = 0;
  if (use_tma) {
    const uint32_t full_rows = p.height / 4, full_cols = p.width / 4;  // blocks that need no clamping
    const uint32_t top = r1 < full_rows ? r1 : full_rows;
    // RGB888: a tile must start on a 16-byte boundary of its row (TMA faults otherwise), i.e. on a multiple of
    // four blocks; the last, shifted tile starts at a_c1 - kBlocksX, so a_c1 is rounded down accordingly.
    const uint32_t usable_cols = kNcomp == 3 ? full_cols & ~3u : full_cols;
    if (top >= r0 + Shape::kBlocksY && usable_cols >= Shape::kBlocksX) {
      a_r1 = top;
      a_c1 = usable_cols;
    }
  }
  if (a_r1 > r0) {
    p.row0 = r0; p.row1 = a_r1; p.col0 = 0; p.col1 = a_c1;
    if (int s = launch_tma<kCodec, kNcomp>(p, info.sm_count, stream)) return s;
    if (a_c1 < grid_cols) {
      p.col0 = a_c1; p.col1 = grid_cols;
      if (int s = launch_generic<kCodec, kNcomp>(p, info.sm_count, stream)) return s;
    }
  }
  if (a_r1 < r1) {
    p.row0 = a_r1; p.row1 = r1; p.col0 = 0; p.col1 = grid_cols;
    if (int s = launch_generic<kCodec, kNcomp>(p, info.sm_count, stream)) return s;
  }
  return ICB_OK;
}

// Block streams are read and written with 8-byte (DXT1, ETC1, PVRTC) or 16-byte (DXT5 output) vector accesses.
bool block_aligned(int codec, const void *p, bool as_input = false) {
  const uintptr_t a = (codec == ICB_CODEC_DXT5 && !as_input) ? 16 : 8;
  return reinterpret_cast<uintptr_t>(p) % a == 0;
}

int encode4x4(int codec, int ncomp, const void *d_src, uint32_t h, uint32_t w, size_t pitch, uint32_t coded_h,
              uint32_t coded_w, int swap_rb, int strategy, uint32_t r0, uint32_t r1, void *d_dst, void *stream) {
  if (!d_src || !d_dst) return fail(ICB_ERR_INVALID, "null device pointer");
  if (h == 0 || w == 0) return fail(ICB_ERR_INVALID, "zero image dimension");
  if (coded_h < h || coded_w < w) return fail(ICB_ERR_INVALID, "coded size %ux%u smaller than image %ux%u", coded_h, coded_w, h, w);
  if (pitch < static_cast<size_t>(w) * ncomp || pitch > 0xffffffffull) return fail(ICB_ERR_INVALID, "bad source pitch %zu", pitch);
  if (strategy < 0 || strategy > 3) return fail(ICB_ERR_INVALID, "unknown ETC strategy %d", strategy);
  // blocks are stored with one 8- or 16-byte vector store each: a misaligned destination would raise a sticky
  // misaligned-address fault instead of an error code
  if (!block_aligned(codec, d_dst)) return fail(ICB_ERR_INVALID, "destination %p is not aligned to the %d-byte block", d_dst, codec == ICB_CODEC_DXT5 ? 16 : 8);
  Encode4x4Params p{};
  p.src = static_cast<const uint8_t *>(d_src);
  p.dst = static_cast<uint8_t *>(d_dst);
  p.height = h;
  p.width = w;
  p.pitch = static_cast<uint32_t>(pitch);
  p.swap_rb = swap_rb ? 1 : 0;
  p.etc_strategy = strategy;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (codec == ICB_CODEC_DXT1 && ncomp == 3) return encode4x4_typed<icb::kCodecDxt1, 3>(p, coded_h, coded_w, r0, r1, s);
  if (codec == ICB_CODEC_DXT1 && ncomp == 4) return encode4x4_typed<icb::kCodecDxt1, 4>(p, coded_h, coded_w, r0, r1, s);
  if (codec == ICB_CODEC_DXT5 && ncomp == 4) return encode4x4_typed<icb::kCodecDxt5, 4>(p, coded_h, coded_w, r0, r1, s);
  if (codec == ICB_CODEC_ETC1 && ncomp == 3) return encode4x4_typed<icb::kCodecEtc1, 3>(p, coded_h, coded_w, r0, r1, s);
  return fail(ICB_ERR_INVALID, "codec %d does not take %d-component pixels", codec, ncomp);
}

// ---- synthetic stream ------------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t splitmix_word(uint64_t seed, uint64_t k) {
  uint64_t z = seed + (k + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void fill_synthetic_kernel(uint8_t *dst, size_t bytes, uint64_t seed, uint64_t byte_offset) {
  // One 64-bit stream word per thread-iteration; unaligned head/tail bytes handled bytewise.
  const uint64_t first_word = byte_offset >> 3, last_word = (byte_offset + bytes + 7) >> 3;
  for (uint64_t k = first_word + blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < last_word;
       k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t v = splitmix_word(seed, k);
    const uint64_t pos = k << 3;  // stream position of byte 0 of this word
    if (pos >= byte_offset && pos + 8 <= byte_offset + bytes && ((reinterpret_cast<uintptr_t>(dst) + (pos - byte_offset)) & 7) == 0) {
      *reinterpret_cast<uint64_t *>(dst + (pos - byte_offset)) = v;
    } else {
      for (int b = 0; b < 8; ++b) {
        const uint64_t q = pos + b;
        if (q >= byte_offset && q < byte_offset + bytes) dst[q - byte_offset] = static_cast<uint8_t>(v >> (8 * b));
      }
    }
  }
}

// ---- parallel host copies ----------------------------------------------------------------------------------
//
// A caller of Compress() usually hands over ordinary pageable memory.  cudaMemcpyAsync from pageable memory goes
// through the driver's single-threaded bounce buffer (measured on the B200 box: 10 GB/s, 26 ms for an 8192^2 RGBA8
// image against 5 ms from pinned memory), so the host pipeline stages such buffers itself: worker threads copy each
// chunk into a ring of pinned buffers while the previous chunk's DMA is in flight, and the packed blocks come back
// the same way, with streaming stores (26 -> 6.9 ms).  Pinned or registered caller memory (icb_host_alloc) skips all of this.
// memcpy with streaming (non-temporal) stores for the bulk: the destination of a staging copy is read next by the DMA
// engine (or much later by the caller), never by this core, so write-allocating its lines only costs a third memory
// stream.  Falls back to memcpy off x86-64 / without AVX2 and for short copies.
#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_streaming_avx2(uint8_t *dst, const uint8_t *src, size_t n) {
  const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head) {
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
  }
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
}
#endif
void copy_bulk(uint8_t *dst, const uint8_t *src, size_t n) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && n >= 4096) {
    copy_streaming_avx2(dst, src, n);
    return;
  }
#endif
  memcpy(dst, src, n);
}

class CopyPool {
 public:
  static CopyPool &get() {
    static CopyPool *pool = new CopyPool();  // never destroyed: workers may outlive static destruction order
    return *pool;
  }
  // `rows` rows of row_bytes from src (stride src_pitch) to dst (stride dst_pitch); never reads or writes beyond the
  // last row's row_bytes.  Blocks until done.  One parallel copy at a time per process.
  void copy_rows(uint8_t *dst, size_t dst_pitch, const uint8_t *src, size_t src_pitch, size_t row_bytes, size_t rows) {
    if (rows == 0 || row_bytes == 0) return;
    Job job{dst, dst_pitch, src, src_pitch, row_bytes, rows};
    const size_t total = rows * row_bytes;
    if (threads_.empty() || total < (1u << 20) || getpid() != pid_) {  // small copy, or a forked child (no workers there)
      run_slice(job, 0, 1);
      return;
    }
    std::lock_guard<std::mutex> call(call_mu_);
    {
      std::lock_guard<std::mutex> lock(mu_);
      job_ = job;
      pending_ = static_cast<int>(threads_.size());
      ++generation_;
    }
    cv_start_.notify_all();
    run_slice(job, 0, static_cast<int>(threads_.size()) + 1);  // the caller takes the first slice
    std::unique_lock<std::mutex> lock(mu_);
    cv_done_.wait(lock, [&] { return pending_ == 0; });
  }

 private:
  struct Job {
    uint8_t *dst;
    size_t dst_pitch;
    const uint8_t *src;
    size_t src_pitch, row_bytes, rows;
  };
  CopyPool() {
    // With streaming stores eight threads already carry 40 GB/s (sixteen measured the same): half the cores, at
    // most eight including the caller.  ICB_STAGING_THREADS overrides (0 = leave pageable copies to the driver).
    int n = static_cast<int>(std::thread::hardware_concurrency()) / 2 - 1;
    if (const char *e = getenv("ICB_STAGING_THREADS")) n = atoi(e) - 1;
    if (n > 7) n = 7;
    pid_ = getpid();
    for (int i = 0; i < n; ++i) threads_.emplace_back([this, i] { worker(i + 1); });
    for (auto &t : threads_) t.detach();
  }
  static void run_slice(const Job &j, int part, int parts) {
    const size_t r0 = j.rows * part / parts, r1 = j.rows * (part + 1) / parts;
    if (r1 <= r0) return;
    if (j.dst_pitch == j.src_pitch) {  // one contiguous span, the last row without its padding
      copy_bulk(j.dst + r0 * j.dst_pitch, j.src + r0 * j.src_pitch, (r1 - r0 - 1) * j.src_pitch + j.row_bytes);
    } else {
      for (size_t r = r0; r < r1; ++r) copy_bulk(j.dst + r * j.dst_pitch, j.src + r * j.src_pitch, j.row_bytes);
    }
  }
  void worker(int part) {
    uint64_t seen = 0;
    while (true) {
      Job job;
      int parts;
      {
        std::unique_lock<std::mutex> lock(mu_);
        cv_start_.wait(lock, [&] { return generation_ != seen; });
        seen = generation_;
        job = job_;
        parts = static_cast<int>(threads_.size()) + 1;
      }
      run_slice(job, part, parts);
      {
        std::lock_guard<std::mutex> lock(mu_);
        if (--pending_ == 0) cv_done_.notify_all();
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex call_mu_, mu_;
  std::condition_variable cv_start_, cv_done_;
  Job job_{};
  uint64_t generation_ = 0;
  int pending_ = 0;
  pid_t pid_ = 0;
};

// True when cudaMemcpyAsync from/to p would go through the driver's pageable path.
bool is_pageable(const void *p) {
  if (getenv("ICB_STAGING_THREADS") && atoi(getenv("ICB_STAGING_THREADS")) == 0) return false;  // staging switched off
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered;
}

// ---- host-buffer pipeline --------------------------------------------------------------------------------

struct HostPipe {  // per host thread, per device: streams, events and grow-only device buffers
  int device = -1;
  cudaStream_t copy_in = nullptr, compute = nullptr, copy_out = nullptr;
  static constexpr int kMaxChunks = 16;
  cudaEvent_t in_done[kMaxChunks] = {}, enc_done[kMaxChunks] = {};
  void *d_src = nullptr, *d_dst = nullptr, *d_scratch = nullptr;
  size_t src_cap = 0, dst_cap = 0, scratch_cap = 0;
  // pinned staging rings for pageable caller memory (see CopyPool)
  static constexpr int kStageBufs = 3;
  void *stage_in[kStageBufs] = {}, *stage_out[kStageBufs] = {};
  size_t stage_in_cap = 0, stage_out_cap = 0;
  cudaEvent_t stage_in_free[kStageBufs] = {}, out_done[kMaxChunks] = {};

  // Creates the streams and events on the CURRENT device (once per pipe; a pipe never changes device).
  int prepare() {
    int dev = 0;
    ICB_CUDA(cudaGetDevice(&dev));
    if (device == dev) return ICB_OK;
    release();
    device = dev;  // from here on release() has something to undo, also after a partial failure below
    ICB_CUDA(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking));
    ICB_CUDA(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
    ICB_CUDA(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < kMaxChunks; ++i) {
      ICB_CUDA(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
      ICB_CUDA(cudaEventCreateWithFlags(&enc_done[i], cudaEventDisableTiming));
      ICB_CUDA(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < kStageBufs; ++i) ICB_CUDA(cudaEventCreateWithFlags(&stage_in_free[i], cudaEventDisableTiming));
    return ICB_OK;
  }
  // Waits for everything this pipe has in flight (error paths: the next user may free or overwrite its buffers).
  void quiesce() {
    if (device < 0) return;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != device) cudaSetDevice(device);
    if (copy_in) cudaStreamSynchronize(copy_in);
    if (compute) cudaStreamSynchronize(compute);
    if (copy_out) cudaStreamSynchronize(copy_out);
    cudaGetLastError();
    if (cur != device) cudaSetDevice(cur);
  }
  size_t device_bytes() const { return src_cap + dst_cap + scratch_cap; }
  size_t pinned_bytes() const { return kStageBufs * (stage_in_cap + stage_out_cap); }
  static int grow_pinned(void *(&bufs)[kStageBufs], size_t *cap, size_t need) {
    if (*cap >= need) return ICB_OK;
    for (int i = 0; i < kStageBufs; ++i) {
      if (bufs[i]) cudaFreeHost(bufs[i]);
      bufs[i] = nullptr;
    }
    *cap = 0;
    for (int i = 0; i < kStageBufs; ++i) ICB_CUDA(cudaMallocHost(&bufs[i], need));
    *cap = need;
    return ICB_OK;
  }
  static int grow(void **ptr, size_t *cap, size_t need) {
    if (*cap >= need) return ICB_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    ICB_CUDA(cudaMalloc(ptr, need));
    *cap = need;
    return ICB_OK;
  }
  void release() {
    if (device < 0) return;
    cudaFree(d_src); cudaFree(d_dst); cudaFree(d_scratch);
    d_src = d_dst = d_scratch = nullptr;
    src_cap = dst_cap = scratch_cap = 0;
    for (int i = 0; i < kStageBufs; ++i) {
      if (stage_in[i]) cudaFreeHost(stage_in[i]);
      if (stage_out[i]) cudaFreeHost(stage_out[i]);
      if (stage_in_free[i]) cudaEventDestroy(stage_in_free[i]);
      stage_in[i] = stage_out[i] = nullptr;
      stage_in_free[i] = nullptr;
    }
    stage_in_cap = stage_out_cap = 0;
    for (int i = 0; i < kMaxChunks; ++i) {
      if (in_done[i]) cudaEventDestroy(in_done[i]);
      if (enc_done[i]) cudaEventDestroy(enc_done[i]);
      if (out_done[i]) cudaEventDestroy(out_done[i]);
      in_done[i] = enc_done[i] = out_done[i] = nullptr;
    }
    if (copy_in) cudaStreamDestroy(copy_in);
    if (compute) cudaStreamDestroy(compute);
    if (copy_out) cudaStreamDestroy(copy_out);
    copy_in = compute = copy_out = nullptr;
    device = -1;
  }
  // Deliberately no destructor work: at process exit the CUDA context may already be gone.
};

// Pipes live in a process-wide pool, one free list per device, and are LEASED for the duration of one host-buffer call:
// a thread that exits leaves nothing behind (round 1 kept them thread_local, so a thread-per-request caller leaked
// three streams, 51 events and up to ~300 MB of device + pinned memory per dead thread), and concurrent callers each
// get their own pipe.  Buffers are grow-only while a pipe is pooled -- a pipe that has encoded an 8192^2 RGBA8 image
// keeps 268 MB + the output on the device -- so the pool holds at most `max concurrent calls` pipes per device;
// icb_trim() releases the idle ones.
class PipePool {
 public:
  static PipePool &get() {
    static PipePool *pool = new PipePool();  // never destroyed (see HostPipe: no CUDA calls at static destruction)
    return *pool;
  }
  // A prepared pipe for the CURRENT device.
  int acquire(HostPipe **out) {
    int dev = 0;
    ICB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(ICB_ERR_CUDA, "device ordinal %d out of range", dev);
    HostPipe *pipe = nullptr;
    {
      std::lock_guard<std::mutex> lock(mu_);
      if (!idle_[dev].empty()) {
        pipe = idle_[dev].back();
        idle_[dev].pop_back();
      }
    }
    if (!pipe) {
      pipe = new HostPipe();
      if (int s = pipe->prepare()) {
        pipe->release();
        delete pipe;
        return s;
      }
    }
    *out = pipe;
    return ICB_OK;
  }
  void give_back(HostPipe *pipe, bool clean) {
    if (!pipe) return;
    if (!clean) pipe->quiesce();
    std::lock_guard<std::mutex> lock(mu_);
    idle_[pipe->device].push_back(pipe);
  }
  // Frees every idle pipe (streams, events, device and pinned buffers).  Returns the bytes released.
  size_t trim() {
    std::vector<HostPipe *> victims;
    {
      std::lock_guard<std::mutex> lock(mu_);
      for (auto &list : idle_) {
        victims.insert(victims.end(), list.begin(), list.end());
        list.clear();
      }
    }
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    size_t bytes = 0;
    for (HostPipe *pipe : victims) {
      bytes += pipe->device_bytes() + pipe->pinned_bytes();
      cudaSetDevice(pipe->device);
      pipe->quiesce();
      pipe->release();
      delete pipe;
    }
    if (have_cur) cudaSetDevice(cur);
    cudaGetLastError();
    return bytes;
  }
  size_t idle_count() {
    std::lock_guard<std::mutex> lock(mu_);
    size_t n = 0;
    for (auto &list : idle_) n += list.size();
    return n;
  }

 private:
  static constexpr int kMaxDevices = 64;
  std::mutex mu_;
  std::vector<HostPipe *> idle_[kMaxDevices];
};

// The lease of one call: pipes are handed back when it goes out of scope -- after a quiesce unless the call said
// it finished cleanly (every exit path of the host entry points that does not reach `clean = true` is an error path,
// possibly with copies or kernels still in flight).
struct PipeLease {
  HostPipe *pipes[16] = {};
  int count = 0;
  bool clean = false;
  int acquire(HostPipe **out) {  // for the current device
    if (count >= 16) return fail(ICB_ERR_INVALID, "too many devices in one call");
    if (int s = PipePool::get().acquire(out)) return s;
    pipes[count++] = *out;
    return ICB_OK;
  }
  int finish(int status) {
    clean = status == ICB_OK;
    return status;
  }
  ~PipeLease() {
    for (int i = 0; i < count; ++i) PipePool::get().give_back(pipes[i], clean);
  }
};

std::atomic<int> g_host_devices{-1};  // icb_set_host_devices(); -1 = ICB_HOST_DEVICES from the environment, else one

// Devices icb_compress_host spreads one image over: the current device only, unless ICB_HOST_DEVICES=N|all asks for
// more -- then the current device and the next N-1 ordinals (modulo the device count).
int host_devices(int *devs, int *count) {
  int cur = 0, total = 0;
  ICB_CUDA(cudaGetDevice(&cur));
  ICB_CUDA(cudaGetDeviceCount(&total));
  int want = g_host_devices.load(std::memory_order_relaxed);
  if (want == 0) want = total;  // "all"
  if (want < 0) {
    want = 1;
    if (const char *e = getenv("ICB_HOST_DEVICES")) want = strcmp(e, "all") == 0 ? total : atoi(e);
  }
  if (want < 1) want = 1;
  if (want > total) want = total;
  if (want > 16) want = 16;
  for (int i = 0; i < want; ++i) devs[i] = (cur + i) % total;
  *count = want;
  return ICB_OK;
}

// Contiguous host -> device copy on `stream`; pageable sources go through the pinned ring in 16 MiB pieces so that the
// parallel host copy of one piece overlaps the DMA of the previous one.
int upload_contiguous(HostPipe &pipe, void *d_dst, const void *h_src, size_t bytes, cudaStream_t stream) {
  if (!is_pageable(h_src)) {
    ICB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
    return ICB_OK;
  }
  constexpr size_t kPiece = 16u << 20;
  if (int s = HostPipe::grow_pinned(pipe.stage_in, &pipe.stage_in_cap, kPiece)) return s;
  size_t piece = 0;
  for (size_t off = 0; off < bytes; off += kPiece, ++piece) {
    const size_t n = bytes - off < kPiece ? bytes - off : kPiece;
    const int b = static_cast<int>(piece % HostPipe::kStageBufs);
    if (piece >= HostPipe::kStageBufs) ICB_CUDA(cudaEventSynchronize(pipe.stage_in_free[b]));
    CopyPool::get().copy_rows(static_cast<uint8_t *>(pipe.stage_in[b]), n, static_cast<const uint8_t *>(h_src) + off, n, n, 1);
    ICB_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(d_dst) + off, pipe.stage_in[b], n, cudaMemcpyHostToDevice, stream));
    ICB_CUDA(cudaEventRecord(pipe.stage_in_free[b], stream));
  }
  return ICB_OK;
}

// Contiguous device -> host copy on `stream`, then waits for it; pageable destinations are filled from the pinned ring
// by the copy pool, the DMA of piece k+1 overlapping the host copy of piece k.
int download_contiguous_sync(HostPipe &pipe, void *h_dst, const void *d_src, size_t bytes, cudaStream_t stream) {
  if (!is_pageable(h_dst)) {
    ICB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, stream));
    ICB_CUDA(cudaStreamSynchronize(stream));
    return ICB_OK;
  }
  constexpr size_t kPiece = 16u << 20;
  if (int s = HostPipe::grow_pinned(pipe.stage_out, &pipe.stage_out_cap, bytes < kPiece ? bytes : kPiece)) return s;
  const size_t cap = pipe.stage_out_cap, pieces = (bytes + cap - 1) / cap;
  auto issue = [&](size_t k) -> int {
    const size_t off = k * cap, n = bytes - off < cap ? bytes - off : cap;
    const int b = static_cast<int>(k % HostPipe::kStageBufs);
    ICB_CUDA(cudaMemcpyAsync(pipe.stage_out[b], static_cast<const uint8_t *>(d_src) + off, n, cudaMemcpyDeviceToHost, stream));
    ICB_CUDA(cudaEventRecord(pipe.out_done[b], stream));
    return ICB_OK;
  };
  if (pieces)
    if (int s = issue(0)) return s;
  for (size_t k = 0; k < pieces; ++k) {
    if (k + 1 < pieces)
      if (int s = issue(k + 1)) return s;  // buffer (k+1) % 3 was emptied at iteration k - 2
    const size_t off = k * cap, n = bytes - off < cap ? bytes - off : cap;
    const int b = static_cast<int>(k % HostPipe::kStageBufs);
    ICB_CUDA(cudaEventSynchronize(pipe.out_done[b]));
    CopyPool::get().copy_rows(static_cast<uint8_t *>(h_dst) + off, n, static_cast<const uint8_t *>(pipe.stage_out[b]), n, n, 1);
  }
  ICB_CUDA(cudaStreamSynchronize(stream));
  return ICB_OK;
}

}  // namespace

// =============================================================================================================
// extern "C"
// =============================================================================================================

extern "C" {

int icb_abi_version(void) { return ICB_ABI_VERSION; }

const char *icb_last_error(void) { return t_last_error.c_str(); }

int icb_device_count(void) {
  int n = 0;
  ICB_CUDA(cudaGetDeviceCount(&n));
  return n;
}

uint64_t icb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int icb_set_tma_mode(int mode) { return g_tma_mode.exchange(mode < 0 ? -1 : (mode ? 1 : 0)); }

size_t icb_compressed_size(int codec, uint32_t coded_height, uint32_t coded_width) {
  if (coded_height == 0 || coded_width == 0) return 0;
  const size_t blocks = static_cast<size_t>((coded_height + 3) / 4) * ((coded_width + 3) / 4);
  switch (codec) {
    case ICB_CODEC_DXT1: return blocks * 8;
    case ICB_CODEC_DXT5: return blocks * 16;
    case ICB_CODEC_ETC1: return blocks * 8;
    case ICB_CODEC_PVRTC2: return static_cast<size_t>(coded_width) * coded_height / 4;
    default: return 0;
  }
}

int icb_dxt1_encode_rgb8(const void *d_src, uint32_t h, uint32_t w, size_t pitch, uint32_t ch, uint32_t cw, int swap_rb,
                         void *d_dst, void *stream) {
  return encode4x4(ICB_CODEC_DXT1, 3, d_src, h, w, pitch, ch, cw, swap_rb, 0, 0, (ch + 3) / 4, d_dst, stream);
}
int icb_dxt1_encode_rgba8(const void *d_src, uint32_t h, uint32_t w, size_t pitch, uint32_t ch, uint32_t cw,
                          int swap_rb, void *d_dst, void *stream) {
  return encode4x4(ICB_CODEC_DXT1, 4, d_src, h, w, pitch, ch, cw, swap_rb, 0, 0, (ch + 3) / 4, d_dst, stream);
}
int icb_dxt5_encode_rgba8(const void *d_src, uint32_t h, uint32_t w, size_t pitch, uint32_t ch, uint32_t cw,
                          int swap_rb, void *d_dst, void *stream) {
  return encode4x4(ICB_CODEC_DXT5, 4, d_src, h, w, pitch, ch, cw, swap_rb, 0, 0, (ch + 3) / 4, d_dst, stream);
}
int icb_etc1_encode_rgb8(const void *d_src, uint32_t h, uint32_t w, size_t pitch, uint32_t ch, uint32_t cw,
                         int strategy, void *d_dst, void *stream) {
  return encode4x4(ICB_CODEC_ETC1, 3, d_src, h, w, pitch, ch, cw, 0, strategy, 0, (ch + 3) / 4, d_dst, stream);
}

int icb_encode4x4_stripe(int codec, int ncomp, const void *d_src, uint32_t h, uint32_t w, size_t pitch, uint32_t ch,
                         uint32_t cw, int swap_rb, int strategy, uint32_t r0, uint32_t r1, void *d_dst, void *stream) {
  return encode4x4(codec, ncomp, d_src, h, w, pitch, ch, cw, swap_rb, strategy, r0, r1, d_dst, stream);
}

// low-resolution A/B colour pairs (8 B per block) + 2-bit modulation per pixel (2 B per 8 pixels)
size_t icb_pvrtc2_scratch_size(uint32_t h, uint32_t w) {
  return static_cast<size_t>(w / 8) * (h / 4) * 4 * 2 + static_cast<size_t>(w / 8) * h * 2;
}

namespace {
// Shared by the whole-image and the stripe entry points: block rows [r0, r1) of an h x w image whose resident pixel
// rows start at image row src_row0 (see PvrtcParams).
int pvrtc_launch(const void *d_src, const void *d_first_pixel, uint32_t h, uint32_t w, uint32_t src_row0, uint32_t r0,
                 uint32_t r1, bool whole, void *d_dst, void *d_scratch, cudaStream_t st) {
  DeviceInfo info;
  if (int s = device_info(&info)) return s;
  void *scratch = d_scratch;
  if (reinterpret_cast<uintptr_t>(scratch) % 8 != 0) return fail(ICB_ERR_INVALID, "PVRTC scratch must be 8-byte aligned");
  if (!scratch) {
    cudaMemPool_t pool;
    if (int s = scratch_pool(&pool)) return s;
    ICB_CUDA(cudaMallocFromPoolAsync(&scratch, icb_pvrtc2_scratch_size(h, w), pool, st));
  }
  const icb::PvrtcParams p = icb::pvrtc_make_params(d_src, d_first_pixel, scratch, d_dst, h, w, src_row0, r0, r1, whole);
  const uint32_t lw = w / 8;
  // Programmatic dependent launch for all three (see pvrtc_kernels.cuh); ICB_NO_PDL=1 launches them plainly.
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = getenv("ICB_NO_PDL") ? 0 : 1;
  cfg.gridDim = dim3((lw * p.morph_rows + 127) / 128);
  cfg.blockDim = dim3(128);
  cudaError_t e = cudaLaunchKernelEx(&cfg, icb::pvrtc_morph_kernel, p);
  // Modulate + Pack fused through shared memory exists and is bit-exact (tests), but measured SLOWER than the two
  // separate kernels on B200 (4096^2: 50.4 vs 48.5 us; both are instruction-issue bound and the tile's extra row,
  // extra column and partial warps add 10 % instructions, more than the third launch costs under programmatic
  // dependent launch), so it is opt-in: ICB_PVRTC_FUSED=1.
  const bool fused = icb::pvrtc_use_fused(h, w, whole) && getenv("ICB_PVRTC_FUSED") != nullptr;
  if (e == cudaSuccess && fused) {  // one CTA per 32 x 8-block tile
    cfg.gridDim = dim3(lw / icb::kFusedBx, (h / 4) / icb::kFusedBy);
    cfg.blockDim = dim3(icb::kFusedThreads);
    e = cudaLaunchKernelEx(&cfg, icb::pvrtc_modpack_kernel, p);
  }
  if (e == cudaSuccess && !fused) {
    cfg.gridDim = dim3((lw * p.mod_units + icb::kModThreads - 1) / icb::kModThreads);
    cfg.blockDim = dim3(icb::kModThreads);
    e = cudaLaunchKernelEx(&cfg, icb::pvrtc_modulate_kernel, p);
  }
  if (e == cudaSuccess && !fused) {
    cfg.gridDim = dim3((lw * p.pack_rows + 127) / 128);
    cfg.blockDim = dim3(128);
    e = cudaLaunchKernelEx(&cfg, icb::pvrtc_pack_kernel, p);
  }
  if (e == cudaSuccess) {
    g_launches.fetch_add(fused ? 2 : 3, std::memory_order_relaxed);
    e = cudaGetLastError();
  }
  // the library's own scratch goes back to the pool on every path (stream-ordered: after whatever did launch)
  if (!d_scratch) {
    const cudaError_t f = cudaFreeAsync(scratch, st);
    if (e == cudaSuccess) e = f;
  }
  if (e != cudaSuccess) return fail(ICB_ERR_CUDA, "PVRTC launch: %s", cudaGetErrorString(e));
  return ICB_OK;
}

int pvrtc_check_shape(uint32_t h, uint32_t w) {
  if (h == 0 || w == 0) return fail(ICB_ERR_INVALID, "zero image dimension");
  if ((w & (w - 1)) || (h & (h - 1)) || w != h || w % 8 != 0 || h % 4 != 0)
    return fail(ICB_ERR_UNSUPPORTED, "PVRTC needs a square power-of-two image of at least 8x8, got %ux%u", h, w);
  return ICB_OK;
}
}  // namespace

int icb_pvrtc2_encode_rgba8(const void *d_src, uint32_t h, uint32_t w, void *d_dst, void *d_scratch, void *stream) {
  if (!d_src || !d_dst) return fail(ICB_ERR_INVALID, "null device pointer");
  if (int s = pvrtc_check_shape(h, w)) return s;
  if (reinterpret_cast<uintptr_t>(d_src) % 16 != 0) return fail(ICB_ERR_INVALID, "PVRTC source must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(d_dst) % 8 != 0) return fail(ICB_ERR_INVALID, "PVRTC destination must be 8-byte aligned");
  return pvrtc_launch(d_src, d_src, h, w, 0, 0, h / 4, true, d_dst, d_scratch, static_cast<cudaStream_t>(stream));
}

int icb_pvrtc2_encode_stripe(const void *d_rows, const void *d_first_pixel, uint32_t h, uint32_t w, uint32_t block_row_begin,
                             uint32_t block_row_end, void *d_dst, void *d_scratch, void *stream) {
  if (!d_rows || !d_first_pixel || !d_dst) return fail(ICB_ERR_INVALID, "null device pointer");
  if (int s = pvrtc_check_shape(h, w)) return s;
  if (reinterpret_cast<uintptr_t>(d_rows) % 16 != 0) return fail(ICB_ERR_INVALID, "PVRTC source must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(d_dst) % 8 != 0) return fail(ICB_ERR_INVALID, "PVRTC destination must be 8-byte aligned");
  if (reinterpret_cast<uintptr_t>(d_first_pixel) % 4 != 0) return fail(ICB_ERR_INVALID, "first-pixel copy must be 4-byte aligned");
  const uint32_t lh = h / 4;
  if (block_row_begin >= block_row_end || block_row_end > lh)
    return fail(ICB_ERR_INVALID, "block row range [%u,%u) outside grid of %u rows", block_row_begin, block_row_end, lh);
  if (block_row_end - block_row_begin + 2 > lh)
    return fail(ICB_ERR_INVALID, "a stripe plus its two halo block rows must not exceed the image (use icb_pvrtc2_encode_rgba8)");
  const uint32_t src_row0 = 4 * ((block_row_begin + lh - 1) & (lh - 1));
  return pvrtc_launch(d_rows, d_first_pixel, h, w, src_row0, block_row_begin, block_row_end, false, d_dst, d_scratch,
                      static_cast<cudaStream_t>(stream));
}

int icb_decode4x4(int codec, const void *d_blocks, uint32_t h, uint32_t w, uint32_t block_cols, int swap_rb, void *d_dst,
                  size_t dst_pitch, void *stream) {
  if (!d_blocks || !d_dst) return fail(ICB_ERR_INVALID, "null device pointer");
  if (h == 0 || w == 0 || block_cols == 0) return fail(ICB_ERR_INVALID, "zero dimension");
  if (codec < ICB_CODEC_DXT1 || codec > ICB_CODEC_ETC1) return fail(ICB_ERR_INVALID, "codec %d has no decoder", codec);
  if (!block_aligned(codec, d_blocks, true)) return fail(ICB_ERR_INVALID, "block stream %p is not 8-byte aligned", d_blocks);
  const size_t ncomp = codec == ICB_CODEC_DXT5 ? 4 : 3;
  if (dst_pitch < w * ncomp || dst_pitch > 0xffffffffull) return fail(ICB_ERR_INVALID, "bad destination pitch %zu", dst_pitch);
  DeviceInfo info;
  if (int s = device_info(&info)) return s;
  icb::Decode4x4Params p;
  p.blocks = static_cast<const uint8_t *>(d_blocks);
  p.dst = static_cast<uint8_t *>(d_dst);
  p.height = h;
  p.width = w;
  p.pitch = static_cast<uint32_t>(dst_pitch);
  p.block_cols = block_cols;
  p.block_rows = (h + 3) / 4;
  p.swap_rb = swap_rb ? 1 : 0;
  // x covers the block columns 128 at a time; y strides the block rows with enough CTAs to fill the GPU
  const uint32_t gx = (block_cols + 127) / 128;
  uint32_t gy = gx ? (static_cast<uint32_t>(info.sm_count) * 32 + gx - 1) / gx : 0;
  if (gy > p.block_rows) gy = p.block_rows;
  if (gy > 65535u) gy = 65535u;
  if (gx == 0 || gy == 0) return ICB_OK;
  const dim3 grid(gx, gy);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (codec == ICB_CODEC_DXT1) icb::decode4x4_kernel<0><<<grid, 128, 0, st>>>(p);
  if (codec == ICB_CODEC_DXT5) icb::decode4x4_kernel<1><<<grid, 128, 0, st>>>(p);
  if (codec == ICB_CODEC_ETC1) icb::decode4x4_kernel<2><<<grid, 128, 0, st>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  ICB_CUDA(cudaGetLastError());
  return ICB_OK;
}

int icb_decompress_host(int codec, int format, uint32_t h, uint32_t w, uint32_t block_cols, const void *blocks,
                        size_t blocks_size, void *dst, size_t dst_size) {
  if (!blocks || !dst || h == 0 || w == 0 || block_cols == 0) return fail(ICB_ERR_INVALID, "null buffer or zero dimension");
  if (codec < ICB_CODEC_DXT1 || codec > ICB_CODEC_ETC1) return fail(ICB_ERR_INVALID, "codec %d has no decoder", codec);
  if (format < ICB_RGB || format > ICB_BGRA) return fail(ICB_ERR_INVALID, "unknown format %d", format);
  const size_t ncomp = codec == ICB_CODEC_DXT5 ? 4 : 3, block_bytes = codec == ICB_CODEC_DXT5 ? 16 : 8;
  const size_t need_in = static_cast<size_t>((h + 3) / 4) * block_cols * block_bytes, need_out = static_cast<size_t>(h) * w * ncomp;
  if (blocks_size < need_in) return fail(ICB_ERR_SIZE, "block stream is %zu bytes, need %zu", blocks_size, need_in);
  if (dst_size != need_out) return fail(ICB_ERR_SIZE, "destination is %zu bytes, need %zu", dst_size, need_out);
  PipeLease lease;
  HostPipe *pipe_ptr = nullptr;
  if (int s = lease.acquire(&pipe_ptr)) return s;
  HostPipe &pipe = *pipe_ptr;
  if (int s = HostPipe::grow(&pipe.d_dst, &pipe.dst_cap, need_in)) return s;     // blocks live in the "dst" buffer
  if (int s = HostPipe::grow(&pipe.d_src, &pipe.src_cap, need_out)) return s;    // pixels in the "src" buffer
  if (int s = upload_contiguous(pipe, pipe.d_dst, blocks, need_in, pipe.compute)) return s;
  const int swap_rb = (format == ICB_BGR || format == ICB_BGRA);
  if (int s = icb_decode4x4(codec, pipe.d_dst, h, w, block_cols, swap_rb, pipe.d_src, w * ncomp, pipe.compute)) return s;
  return lease.finish(download_contiguous_sync(pipe, dst, pipe.d_src, need_out, pipe.compute));
}

// ---- compressed-domain operations ----------------------------------------------------------------------------

static inline uint32_t blocks_of(uint32_t pixels) { return (pixels + 3) / 4; }

static int blockop_grid(uint64_t total, int threads, uint32_t *grid) {
  DeviceInfo info;
  if (int s = device_info(&info)) return s;
  const uint64_t want = (total + threads - 1) / threads;
  *grid = static_cast<uint32_t>(want < static_cast<uint64_t>(info.sm_count) * 32 ? want : info.sm_count * 32);
  return ICB_OK;
}

static int check_4x4_codec(int codec, int strategy, const void *d_in = nullptr, const void *d_out = nullptr) {
  if (codec < ICB_CODEC_DXT1 || codec > ICB_CODEC_ETC1) return fail(ICB_ERR_INVALID, "codec %d is not a 4x4 block codec", codec);
  if (strategy < 0 || strategy > 3) return fail(ICB_ERR_INVALID, "unknown ETC strategy %d", strategy);
  if (!block_aligned(codec, d_in, true)) return fail(ICB_ERR_INVALID, "block stream %p is not 8-byte aligned", d_in);
  if (!block_aligned(codec, d_out)) return fail(ICB_ERR_INVALID, "destination %p is not aligned to the %d-byte block", d_out, codec == ICB_CODEC_DXT5 ? 16 : 8);
  return ICB_OK;
}

int icb_downsample4x4(int codec, int strategy, const void *d_blocks, uint32_t h, uint32_t w, void *d_dst, void *stream) {
  if (!d_blocks || !d_dst) return fail(ICB_ERR_INVALID, "null device pointer");
  if (h == 0 || w == 0) return fail(ICB_ERR_INVALID, "zero dimension");
  if (int s = check_4x4_codec(codec, strategy, d_blocks, d_dst)) return s;
  icb::Downsample4x4Params p;
  p.in = static_cast<const uint8_t *>(d_blocks);
  p.out = static_cast<uint8_t *>(d_dst);
  p.in_rows = blocks_of(h);
  p.in_cols = blocks_of(w);
  // compressor4x4_helper.h:281-284: even block counts, except a single block
  if ((p.in_rows > 1 && p.in_rows % 2) || (p.in_cols > 1 && p.in_cols % 2))
    return fail(ICB_ERR_UNSUPPORTED, "Downsample needs an even number of blocks per dimension (or one), got %ux%u", p.in_rows, p.in_cols);
  if (p.in_rows == 1 && p.in_cols == 1 && (h == 3 || w == 3))  // :335
    return fail(ICB_ERR_UNSUPPORTED, "Downsample of a single block refuses a 3-pixel dimension");
  p.out_rows = p.in_rows > 1 ? p.in_rows / 2 : 1;
  p.out_cols = p.in_cols > 1 ? p.in_cols / 2 : 1;
  p.height = h;
  p.width = w;
  p.etc_strategy = strategy;
  uint32_t grid;
  if (int s = blockop_grid(static_cast<uint64_t>(p.out_rows) * p.out_cols, icb::kBlockOpThreads, &grid)) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (codec == ICB_CODEC_DXT1) icb::downsample4x4_kernel<icb::kCodecDxt1><<<grid, icb::kBlockOpThreads, 0, st>>>(p);
  if (codec == ICB_CODEC_DXT5) icb::downsample4x4_kernel<icb::kCodecDxt5><<<grid, icb::kBlockOpThreads, 0, st>>>(p);
  if (codec == ICB_CODEC_ETC1) icb::downsample4x4_kernel<icb::kCodecEtc1><<<grid, icb::kBlockOpThreads, 0, st>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  ICB_CUDA(cudaGetLastError());
  return ICB_OK;
}

int icb_pad4x4(int codec, int strategy, const void *d_blocks, uint32_t ch, uint32_t cw, uint32_t ph, uint32_t pw,
               void *d_dst, void *stream) {
  if (!d_blocks || !d_dst) return fail(ICB_ERR_INVALID, "null device pointer");
  if (ch == 0 || cw == 0) return fail(ICB_ERR_INVALID, "zero dimension");
  if (int s = check_4x4_codec(codec, strategy, d_blocks, d_dst)) return s;
  const size_t block_bytes = codec == ICB_CODEC_DXT5 ? 16 : 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  icb::Pad4x4Params p;
  p.in = static_cast<const uint8_t *>(d_blocks);
  p.out = static_cast<uint8_t *>(d_dst);
  p.in_rows = blocks_of(ch);
  p.in_cols = blocks_of(cw);
  if (ch >= ph && cw >= pw) {  // nothing to pad: the reference duplicates the image (compressor4x4_helper.h:404-408)
    ICB_CUDA(cudaMemcpyAsync(d_dst, d_blocks, static_cast<size_t>(p.in_rows) * p.in_cols * block_bytes, cudaMemcpyDeviceToDevice, st));
    return ICB_OK;
  }
  p.out_rows = blocks_of(ph);
  p.out_cols = blocks_of(pw);
  if (p.out_rows < p.in_rows || p.out_cols < p.in_cols)
    return fail(ICB_ERR_UNSUPPORTED, "Pad to %ux%u would shrink one dimension of a %ux%u image while growing the other", ph, pw, ch, cw);
  p.etc_strategy = strategy;
  uint32_t grid;
  if (int s = blockop_grid(static_cast<uint64_t>(p.out_rows) * p.out_cols, icb::kBlockOpThreads, &grid)) return s;
  if (codec == ICB_CODEC_DXT1) icb::pad4x4_kernel<icb::kCodecDxt1><<<grid, icb::kBlockOpThreads, 0, st>>>(p);
  if (codec == ICB_CODEC_DXT5) icb::pad4x4_kernel<icb::kCodecDxt5><<<grid, icb::kBlockOpThreads, 0, st>>>(p);
  if (codec == ICB_CODEC_ETC1) icb::pad4x4_kernel<icb::kCodecEtc1><<<grid, icb::kBlockOpThreads, 0, st>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  ICB_CUDA(cudaGetLastError());
  return ICB_OK;
}

int icb_copy_subimage4x4(int codec, const void *d_blocks, uint32_t ch, uint32_t cw, uint32_t row, uint32_t col, uint32_t h,
                         uint32_t w, void *d_dst, void *stream) {
  if (!d_blocks || !d_dst) return fail(ICB_ERR_INVALID, "null device pointer");
  if (int s = check_4x4_codec(codec, 0)) return s;
  // compressor4x4_helper.h:556-565
  if (row % 4 || col % 4 || h % 4 || w % 4 || row > ch || col > cw || static_cast<uint64_t>(row) + h > ch ||
      static_cast<uint64_t>(col) + w > cw)
    return fail(ICB_ERR_INVALID, "subimage %ux%u at (%u,%u) is not block-aligned inside %ux%u", h, w, row, col, ch, cw);
  if (h == 0 || w == 0) return ICB_OK;
  const size_t bb = codec == ICB_CODEC_DXT5 ? 16 : 8;
  const uint8_t *src = static_cast<const uint8_t *>(d_blocks) + (static_cast<size_t>(row / 4) * blocks_of(cw) + col / 4) * bb;
  ICB_CUDA(cudaMemcpy2DAsync(d_dst, (w / 4) * bb, src, blocks_of(cw) * bb, (w / 4) * bb, h / 4, cudaMemcpyDeviceToDevice,
                             static_cast<cudaStream_t>(stream)));
  return ICB_OK;
}

int icb_fill_solid4x4(int codec, const uint8_t *colour, uint32_t h, uint32_t w, void *d_dst, void *stream) {
  if (!d_dst || !colour) return fail(ICB_ERR_INVALID, "null pointer");
  if (h == 0 || w == 0) return fail(ICB_ERR_INVALID, "zero dimension");
  if (int s = check_4x4_codec(codec, 0, nullptr, d_dst)) return s;
  const uint32_t packed = colour[0] | (colour[1] << 8) | (colour[2] << 16) |
                          (codec == ICB_CODEC_DXT5 ? static_cast<uint32_t>(colour[3]) << 24 : 0u);
  const uint64_t n = static_cast<uint64_t>(blocks_of(h)) * blocks_of(w);
  uint32_t grid;
  if (int s = blockop_grid(n, 256, &grid)) return s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t *out = static_cast<uint8_t *>(d_dst);
  if (codec == ICB_CODEC_DXT1) icb::fill_solid4x4_kernel<icb::kCodecDxt1><<<grid, 256, 0, st>>>(out, n, packed);
  if (codec == ICB_CODEC_DXT5) icb::fill_solid4x4_kernel<icb::kCodecDxt5><<<grid, 256, 0, st>>>(out, n, packed);
  if (codec == ICB_CODEC_ETC1) icb::fill_solid4x4_kernel<icb::kCodecEtc1><<<grid, 256, 0, st>>>(out, n, packed);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  ICB_CUDA(cudaGetLastError());
  return ICB_OK;
}

int icb_transcode_dxt1_to_etc1(void *d_blocks, size_t num_blocks, void *stream) {
  if (num_blocks == 0) return ICB_OK;
  if (!d_blocks) return fail(ICB_ERR_INVALID, "null device pointer");
  if (reinterpret_cast<uintptr_t>(d_blocks) % 8 != 0) return fail(ICB_ERR_INVALID, "block stream %p is not 8-byte aligned", d_blocks);
  uint32_t grid;
  if (int s = blockop_grid(num_blocks, icb::kBlockOpThreads, &grid)) return s;
  icb::transcode_dxt1_to_etc1_kernel<<<grid, icb::kBlockOpThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint8_t *>(d_blocks), num_blocks);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  ICB_CUDA(cudaGetLastError());
  return ICB_OK;
}

int icb_blockop_host(int op, int codec, int strategy, const uint32_t *args, const void *src, size_t src_size, void *dst,
                     size_t dst_size) {
  if (!dst || (op != ICB_OP_SOLID && !src)) return fail(ICB_ERR_INVALID, "null buffer");
  if (op != ICB_OP_TRANSCODE && !args) return fail(ICB_ERR_INVALID, "null argument list");
  if (op == ICB_OP_TRANSCODE) codec = ICB_CODEC_DXT1;
  if (int s = check_4x4_codec(codec, strategy)) return s;
  const size_t bb = codec == ICB_CODEC_DXT5 ? 16 : 8;
  size_t need_in = 0, need_out = 0;
  switch (op) {
    case ICB_OP_DOWNSAMPLE:
      if (args[0] == 0 || args[1] == 0) return fail(ICB_ERR_INVALID, "zero dimension");
      need_in = static_cast<size_t>(blocks_of(args[0])) * blocks_of(args[1]) * bb;
      need_out = static_cast<size_t>(blocks_of((args[0] + 1) / 2)) * blocks_of((args[1] + 1) / 2) * bb;
      break;
    case ICB_OP_PAD: {
      if (args[0] == 0 || args[1] == 0) return fail(ICB_ERR_INVALID, "zero dimension");
      need_in = static_cast<size_t>(blocks_of(args[0])) * blocks_of(args[1]) * bb;
      const bool copy = args[0] >= args[2] && args[1] >= args[3];
      need_out = copy ? need_in : static_cast<size_t>(blocks_of(args[2])) * blocks_of(args[3]) * bb;
      break;
    }
    case ICB_OP_COPY_SUBIMAGE:
      need_in = static_cast<size_t>(blocks_of(args[0])) * blocks_of(args[1]) * bb;
      need_out = static_cast<size_t>(blocks_of(args[4])) * blocks_of(args[5]) * bb;
      break;
    case ICB_OP_SOLID:
      if (args[0] == 0 || args[1] == 0) return fail(ICB_ERR_INVALID, "zero dimension");
      need_out = static_cast<size_t>(blocks_of(args[0])) * blocks_of(args[1]) * bb;
      break;
    case ICB_OP_TRANSCODE:
      need_in = need_out = src_size / 8 * 8;
      break;
    default:
      return fail(ICB_ERR_INVALID, "unknown block operation %d", op);
  }
  if (op != ICB_OP_SOLID && src_size < need_in) return fail(ICB_ERR_SIZE, "source is %zu bytes, need %zu", src_size, need_in);
  if (dst_size != need_out) return fail(ICB_ERR_SIZE, "destination is %zu bytes, need %zu", dst_size, need_out);
  PipeLease lease;
  HostPipe *pipe_ptr = nullptr;
  if (int s = lease.acquire(&pipe_ptr)) return s;
  HostPipe &pipe = *pipe_ptr;
  if (int s = HostPipe::grow(&pipe.d_src, &pipe.src_cap, need_in > 16 ? need_in : 16)) return s;
  if (int s = HostPipe::grow(&pipe.d_dst, &pipe.dst_cap, need_out > 16 ? need_out : 16)) return s;
  cudaStream_t st = pipe.compute;
  if (need_in)
    if (int u = upload_contiguous(pipe, pipe.d_src, src, need_in, st)) return u;
  int s = ICB_OK;
  void *result = pipe.d_dst;
  switch (op) {
    case ICB_OP_DOWNSAMPLE: s = icb_downsample4x4(codec, strategy, pipe.d_src, args[0], args[1], pipe.d_dst, st); break;
    case ICB_OP_PAD: s = icb_pad4x4(codec, strategy, pipe.d_src, args[0], args[1], args[2], args[3], pipe.d_dst, st); break;
    case ICB_OP_COPY_SUBIMAGE:
      s = icb_copy_subimage4x4(codec, pipe.d_src, args[0], args[1], args[2], args[3], args[4], args[5], pipe.d_dst, st);
      break;
    case ICB_OP_SOLID: {
      const uint8_t colour[4] = {static_cast<uint8_t>(args[2]), static_cast<uint8_t>(args[2] >> 8),
                                 static_cast<uint8_t>(args[2] >> 16), static_cast<uint8_t>(args[2] >> 24)};
      s = icb_fill_solid4x4(codec, colour, args[0], args[1], pipe.d_dst, st);
      break;
    }
    case ICB_OP_TRANSCODE:
      s = icb_transcode_dxt1_to_etc1(pipe.d_src, need_in / 8, st);
      result = pipe.d_src;
      break;
  }
  if (s != ICB_OK) return s;  // the lease quiesces the pipe
  if (need_out) return lease.finish(download_contiguous_sync(pipe, dst, result, need_out, st));
  ICB_CUDA(cudaStreamSynchronize(st));
  return lease.finish(ICB_OK);
}

int icb_fill_synthetic(void *d_dst, size_t bytes, uint64_t seed, uint64_t byte_offset, void *stream) {
  if (bytes == 0) return ICB_OK;
  if (!d_dst) return fail(ICB_ERR_INVALID, "null device pointer");
  const uint64_t words = (bytes + 15) / 8;
  uint64_t grid = (words + 255) / 256;
  if (grid > 148ull * 16) grid = 148ull * 16;
  fill_synthetic_kernel<<<static_cast<uint32_t>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint8_t *>(d_dst), bytes, seed, byte_offset);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  ICB_CUDA(cudaGetLastError());
  return ICB_OK;
}

void *icb_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    fail(ICB_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}

void icb_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int icb_host_register(void *p, size_t bytes) {
  if (!p || bytes == 0) return fail(ICB_ERR_INVALID, "null pointer or zero size");
  ICB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return ICB_OK;
}

int icb_host_unregister(void *p) {
  if (!p) return ICB_OK;
  ICB_CUDA(cudaHostUnregister(p));
  return ICB_OK;
}

// ---- peer-mapped output (multi-GPU stripe path) -----------------------------------------------------------

int icb_device_alloc(size_t bytes, void **d_ptr) {
  if (!d_ptr || bytes == 0) return fail(ICB_ERR_INVALID, "null result pointer or zero size");
  *d_ptr = nullptr;
  ICB_CUDA(cudaMalloc(d_ptr, bytes));
  return ICB_OK;
}

int icb_device_free(void *d_ptr) {
  if (d_ptr) ICB_CUDA(cudaFree(d_ptr));
  return ICB_OK;
}

static_assert(sizeof(cudaIpcMemHandle_t) == ICB_IPC_HANDLE_BYTES, "icb200.h promises a 64-byte handle");

int icb_ipc_export(const void *d_ptr, void *handle) {
  if (!d_ptr || !handle) return fail(ICB_ERR_INVALID, "null pointer");
  cudaIpcMemHandle_t h;
  ICB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
  memcpy(handle, &h, sizeof(h));
  return ICB_OK;
}

int icb_ipc_open(const void *handle, void **d_ptr) {
  if (!handle || !d_ptr) return fail(ICB_ERR_INVALID, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  *d_ptr = nullptr;
  ICB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ICB_OK;
}

int icb_ipc_close(void *d_ptr) {
  if (d_ptr) ICB_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return ICB_OK;
}

// Body of icb_compress_host / icb_ctx_compress_host.  ctx_devs == NULL: the current device (plus ICB_HOST_DEVICES /
// icb_set_host_devices); otherwise exactly those devices, the first being the one that takes the single-device cases.
static int compress_host_impl(const int *ctx_devs, int ctx_ndev, int codec, int format, uint32_t h, uint32_t w,
                              uint32_t padded_h, uint32_t padded_w, uint32_t padding, int strategy, const void *src,
                              void *dst, size_t dst_size) {
  // Same rejections, in the same order of concern, as the reference entry points
  // (dxtc_compressor.cc:739, etc_compressor.cc:751-754, pvrtc_compressor.cc:640-650).
  if (!src || !dst || h == 0 || w == 0) return fail(ICB_ERR_INVALID, "null buffer or zero dimension");
  if (format < ICB_RGB || format > ICB_BGRA) return fail(ICB_ERR_INVALID, "unknown format %d", format);
  const int ncomp = (format == ICB_RGB || format == ICB_BGR) ? 3 : 4;
  const int swap_rb = (format == ICB_BGR || format == ICB_BGRA);
  if (codec == ICB_CODEC_DXT5 && ncomp != 4) return fail(ICB_ERR_INVALID, "DXT5 needs a 4-component format");
  if (codec == ICB_CODEC_ETC1 && format != ICB_RGB) return fail(ICB_ERR_INVALID, "ETC1 supports kRGB only");
  if (codec < ICB_CODEC_DXT1 || codec > ICB_CODEC_PVRTC2) return fail(ICB_ERR_INVALID, "unknown codec %d", codec);

  struct RestoreDevice {  // the caller's current device is unchanged on return, whatever path is taken
    int dev = -1;
    ~RestoreDevice() {
      if (dev >= 0) cudaSetDevice(dev);
    }
  } restore;
  if (ctx_devs) {
    ICB_CUDA(cudaGetDevice(&restore.dev));
    ICB_CUDA(cudaSetDevice(ctx_devs[0]));
  }
  PipeLease lease;
  HostPipe *pipe_ptr = nullptr;
  if (int s = lease.acquire(&pipe_ptr)) return s;
  HostPipe &pipe = *pipe_ptr;

  if (codec == ICB_CODEC_PVRTC2) {
    if ((w & (w - 1)) || (h & (h - 1)) || w != h) return fail(ICB_ERR_UNSUPPORTED, "PVRTC needs a square power-of-two image");
    if (padding != 0) return fail(ICB_ERR_UNSUPPORTED, "PVRTC does not take padded rows");
    if (w % 8 != 0 || h % 4 != 0) return fail(ICB_ERR_UNSUPPORTED, "PVRTC image smaller than one block");
    const size_t need = static_cast<size_t>(w) * h / 4, src_bytes = static_cast<size_t>(w) * h * 4;
    if (dst_size != need) return fail(ICB_ERR_SIZE, "destination is %zu bytes, need %zu", dst_size, need);
    if (int s = HostPipe::grow(&pipe.d_src, &pipe.src_cap, src_bytes)) return s;
    if (int s = HostPipe::grow(&pipe.d_dst, &pipe.dst_cap, need)) return s;
    if (int s = HostPipe::grow(&pipe.d_scratch, &pipe.scratch_cap, icb_pvrtc2_scratch_size(h, w))) return s;
    if (int s = upload_contiguous(pipe, pipe.d_src, src, src_bytes, pipe.compute)) return s;
    if (int s = icb_pvrtc2_encode_rgba8(pipe.d_src, h, w, pipe.d_dst, pipe.d_scratch, pipe.compute)) return s;
    return lease.finish(download_contiguous_sync(pipe, dst, pipe.d_dst, need, pipe.compute));
  }

  const uint32_t coded_h = padded_h > h ? padded_h : h, coded_w = padded_w > w ? padded_w : w;
  const size_t need = icb_compressed_size(codec, coded_h, coded_w);
  if (dst_size != need) return fail(ICB_ERR_SIZE, "destination is %zu bytes, need %zu", dst_size, need);
  const size_t host_pitch = static_cast<size_t>(w) * ncomp + padding;
  // Device copy keeps the caller's pitch when it is already 16-byte aligned (one contiguous copy per chunk, TMA
  // eligible); otherwise rows are re-pitched to a multiple of 16 by cudaMemcpy2D.
  const size_t row_bytes = static_cast<size_t>(w) * ncomp;
  const bool keep_pitch = host_pitch % 16 == 0;
  const size_t dev_pitch = keep_pitch ? host_pitch : (row_bytes + 15) / 16 * 16;
  const uint32_t grid_rows = (coded_h + 3) / 4, grid_cols = (coded_w + 3) / 4;
  const size_t block_bytes = codec == ICB_CODEC_DXT5 ? 16 : 8;
  // Chunks of whole block rows, about 16 MiB of source each, at most kMaxChunks.
  uint32_t chunks = static_cast<uint32_t>((dev_pitch * h + (16u << 20) - 1) / (16u << 20));
  if (chunks < 1) chunks = 1;
  if (chunks > HostPipe::kMaxChunks) chunks = HostPipe::kMaxChunks;
  if (chunks > grid_rows) chunks = grid_rows;
  // whole tile rows per chunk (8 block rows: the tallest tile), so that only the last chunk can leave rows to the
  // generic kernel
  const uint32_t rows_per_chunk = ((grid_rows + chunks - 1) / chunks + 7) / 8 * 8;
  const uint32_t num_chunks = (grid_rows + rows_per_chunk - 1) / rows_per_chunk;

  // Devices: chunk c goes to device c % ndev (ICB_HOST_DEVICES, default one), each over its own PCIe link; every
  // device writes its chunks' blocks straight to their place in the caller's buffer, so nothing is gathered.  Chunks
  // below the image (CompressAndPad with coded_h > h) replicate row h-1, which only the device that uploaded it has:
  // that case stays on one device.
  int devs[16], ndev = 1, home = 0;
  ICB_CUDA(cudaGetDevice(&home));
  if (ctx_devs) {
    ndev = ctx_ndev < 16 ? ctx_ndev : 16;
    for (int k = 0; k < ndev; ++k) devs[k] = ctx_devs[k];
  } else if (int s = host_devices(devs, &ndev)) {
    return s;
  }
  if (static_cast<uint32_t>(ndev) > num_chunks) ndev = static_cast<int>(num_chunks);
  if (coded_h > (h + 3) / 4 * 4) ndev = 1;
  HostPipe *pipes[16];
  pipes[0] = &pipe;
  for (int k = 1; k < ndev; ++k) {
    ICB_CUDA(cudaSetDevice(devs[k]));
    const int s = lease.acquire(&pipes[k]);
    if (s != ICB_OK) {
      cudaSetDevice(home);
      return s;
    }
  }
  // From here on every exit path has to put the caller's device back.
  auto run = [&]() -> int {
    const uint8_t *hsrc = static_cast<const uint8_t *>(src);
    uint8_t *hdst = static_cast<uint8_t *>(dst);
    // Pageable caller memory is staged through pinned rings by the copy pool (both directions independently).
    const bool stage_src = is_pageable(src), stage_dst = is_pageable(dst);
    const size_t chunk_out_bytes = static_cast<size_t>(rows_per_chunk) * grid_cols * block_bytes;
    for (int k = 0; k < ndev; ++k) {
      ICB_CUDA(cudaSetDevice(devs[k]));
      HostPipe &pk = *pipes[k];
      if (int s = HostPipe::grow(&pk.d_src, &pk.src_cap, dev_pitch * h)) return s;
      if (int s = HostPipe::grow(&pk.d_dst, &pk.dst_cap, need)) return s;
      if (stage_src)
        if (int s = HostPipe::grow_pinned(pk.stage_in, &pk.stage_in_cap, static_cast<size_t>(rows_per_chunk) * 4 * dev_pitch)) return s;
      if (stage_dst)
        if (int s = HostPipe::grow_pinned(pk.stage_out, &pk.stage_out_cap, chunk_out_bytes)) return s;
    }
    // Copies chunk c's blocks from its staging buffer to the caller's memory once its D2H has landed.  Chunk c is
    // the (c / ndev)-th chunk of device c % ndev; rings and events are indexed by that local number.
    auto drain = [&](uint32_t c) -> int {
      HostPipe &pk = *pipes[c % ndev];
      const uint32_t local = c / ndev;
      const uint32_t c_r0 = c * rows_per_chunk, c_r1 = c_r0 + rows_per_chunk < grid_rows ? c_r0 + rows_per_chunk : grid_rows;
      const size_t bytes = static_cast<size_t>(c_r1 - c_r0) * grid_cols * block_bytes;
      ICB_CUDA(cudaEventSynchronize(pk.out_done[local]));
      CopyPool::get().copy_rows(hdst + static_cast<size_t>(c_r0) * grid_cols * block_bytes, bytes,
                                static_cast<const uint8_t *>(pk.stage_out[local % HostPipe::kStageBufs]), bytes, bytes, 1);
      return ICB_OK;
    };
    const uint32_t lag = 2u * static_cast<uint32_t>(ndev);  // a chunk is drained two of its own device's chunks later
    uint32_t chunk = 0;
    for (uint32_t r0 = 0; r0 < grid_rows; r0 += rows_per_chunk, ++chunk) {
      const uint32_t r1 = r0 + rows_per_chunk < grid_rows ? r0 + rows_per_chunk : grid_rows;
      HostPipe &pk = *pipes[chunk % ndev];
      const uint32_t local = chunk / ndev;
      if (ndev > 1) ICB_CUDA(cudaSetDevice(devs[chunk % ndev]));
      // Source rows this chunk adds: pixel rows [4*r0, min(4*r1, h)).  Later chunks only ever clamp to row h-1,
      // which the chunk containing it has already uploaded (chunks run in order on the compute stream).
      const uint32_t y0 = 4 * r0 < h ? 4 * r0 : h, y1 = 4 * r1 < h ? 4 * r1 : h;
      if (y1 > y0 && stage_src) {
        const int b = static_cast<int>(local % HostPipe::kStageBufs);
        if (local >= HostPipe::kStageBufs) ICB_CUDA(cudaEventSynchronize(pk.stage_in_free[b]));  // its last DMA has read it
        uint8_t *stage = static_cast<uint8_t *>(pk.stage_in[b]);
        CopyPool::get().copy_rows(stage, dev_pitch, hsrc + y0 * host_pitch, host_pitch, row_bytes, y1 - y0);
        ICB_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(pk.d_src) + y0 * dev_pitch, stage,
                                 static_cast<size_t>(y1 - y0 - 1) * dev_pitch + row_bytes, cudaMemcpyHostToDevice, pk.copy_in));
        ICB_CUDA(cudaEventRecord(pk.stage_in_free[b], pk.copy_in));
      } else if (y1 > y0) {
        if (keep_pitch) {
          const size_t bytes = (y1 == h) ? (static_cast<size_t>(y1 - y0 - 1) * host_pitch + row_bytes)
                                         : static_cast<size_t>(y1 - y0) * host_pitch;
          ICB_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(pk.d_src) + y0 * dev_pitch, hsrc + y0 * host_pitch, bytes,
                                   cudaMemcpyHostToDevice, pk.copy_in));
        } else {
          ICB_CUDA(cudaMemcpy2DAsync(static_cast<uint8_t *>(pk.d_src) + y0 * dev_pitch, dev_pitch, hsrc + y0 * host_pitch,
                                     host_pitch, row_bytes, y1 - y0, cudaMemcpyHostToDevice, pk.copy_in));
        }
      }
      ICB_CUDA(cudaEventRecord(pk.in_done[local], pk.copy_in));
      ICB_CUDA(cudaStreamWaitEvent(pk.compute, pk.in_done[local], 0));
      uint8_t *d_out = static_cast<uint8_t *>(pk.d_dst) + static_cast<size_t>(r0) * grid_cols * block_bytes;
      if (int s = encode4x4(codec, ncomp, pk.d_src, h, w, dev_pitch, coded_h, coded_w, swap_rb, strategy, r0, r1, d_out,
                            pk.compute))
        return s;
      ICB_CUDA(cudaEventRecord(pk.enc_done[local], pk.compute));
      ICB_CUDA(cudaStreamWaitEvent(pk.copy_out, pk.enc_done[local], 0));
      const size_t out_off = static_cast<size_t>(r0) * grid_cols * block_bytes;
      if (stage_dst) {
        // this device's buffer local % 3 was drained two of its chunks ago (below), so it is free again
        ICB_CUDA(cudaMemcpyAsync(pk.stage_out[local % HostPipe::kStageBufs], d_out,
                                 static_cast<size_t>(r1 - r0) * grid_cols * block_bytes, cudaMemcpyDeviceToHost, pk.copy_out));
        ICB_CUDA(cudaEventRecord(pk.out_done[local], pk.copy_out));
        if (chunk >= lag)
          if (int s = drain(chunk - lag)) return s;
      } else {
        ICB_CUDA(cudaMemcpyAsync(hdst + out_off, d_out, static_cast<size_t>(r1 - r0) * grid_cols * block_bytes,
                                 cudaMemcpyDeviceToHost, pk.copy_out));
      }
    }
    if (stage_dst)
      for (uint32_t c = chunk >= lag ? chunk - lag : 0; c < chunk; ++c)
        if (int s = drain(c)) return s;
    for (int k = 0; k < ndev; ++k) {
      if (ndev > 1) ICB_CUDA(cudaSetDevice(devs[k]));
      ICB_CUDA(cudaStreamSynchronize(pipes[k]->copy_out));
      ICB_CUDA(cudaStreamSynchronize(pipes[k]->compute));
    }
    return ICB_OK;
  };
  const int status = run();
  if (ndev > 1) cudaSetDevice(home);
  return lease.finish(status);
}

int icb_compress_host(int codec, int format, uint32_t h, uint32_t w, uint32_t padded_h, uint32_t padded_w,
                      uint32_t padding, int strategy, const void *src, void *dst, size_t dst_size) {
  return compress_host_impl(nullptr, 0, codec, format, h, w, padded_h, padded_w, padding, strategy, src, dst, dst_size);
}

int icb_set_host_devices(int n) {
  int total = 0;
  if (n > 0 && (cudaGetDeviceCount(&total) != cudaSuccess || n > total)) {
    cudaGetLastError();
    return fail(ICB_ERR_INVALID, "icb_set_host_devices(%d): only %d device(s) visible", n, total);
  }
  const int prev = g_host_devices.exchange(n < 0 ? -1 : n);
  return prev < 0 ? 0x7fffffff : prev;
}

size_t icb_trim(void) {
  size_t bytes = PipePool::get().trim();
  // the stream-ordered scratch pools of the devices this process has used
  int cur = 0, total = 0;
  if (cudaGetDevice(&cur) == cudaSuccess && cudaGetDeviceCount(&total) == cudaSuccess) {
    for (int d = 0; d < total && d < 64; ++d) {
      cudaMemPool_t pool = scratch_pool_if_created(d);
      if (pool) cudaMemPoolTrimTo(pool, 0);
    }
  }
  cudaGetLastError();
  return bytes;
}

size_t icb_idle_pipes(void) { return PipePool::get().idle_count(); }

// ---- one image over several GPUs from ONE process -----------------------------

struct icb_ctx {
  int n = 0;
  int devs[16] = {};
  cudaStream_t streams[16] = {};  // [0] unused: the root works on the caller's stream
  cudaEvent_t fork = nullptr, done[16] = {};
  bool peer_stores = false;       // every non-root device can store into the root's memory
};

int icb_ctx_create(int n_dev, const int *dev_ids, icb_ctx **out) {
  if (!out) return fail(ICB_ERR_INVALID, "null result pointer");
  *out = nullptr;
  int total = 0;
  ICB_CUDA(cudaGetDeviceCount(&total));
  if (n_dev == 0) n_dev = total;  // "all"
  if (n_dev < 1 || n_dev > 16 || n_dev > total) return fail(ICB_ERR_INVALID, "asked for %d devices, %d visible (at most 16 per context)", n_dev, total);
  int home = 0;
  ICB_CUDA(cudaGetDevice(&home));
  icb_ctx *ctx = new icb_ctx();
  ctx->n = n_dev;
  for (int k = 0; k < n_dev; ++k) {
    ctx->devs[k] = dev_ids ? dev_ids[k] : k;
    bool dup = ctx->devs[k] < 0 || ctx->devs[k] >= total;
    for (int j = 0; j < k; ++j) dup = dup || ctx->devs[j] == ctx->devs[k];
    if (dup) {
      delete ctx;
      return fail(ICB_ERR_INVALID, "device list entry %d (ordinal %d) is out of range or repeated", k, dev_ids ? dev_ids[k] : k);
    }
  }
  auto bail = [&](int status) {
    icb_ctx_destroy(ctx);
    cudaSetDevice(home);
    return status;
  };
  ctx->peer_stores = true;
  for (int k = 0; k < n_dev; ++k) {
    if (cudaSetDevice(ctx->devs[k]) != cudaSuccess) return bail(fail(ICB_ERR_CUDA, "cudaSetDevice(%d) failed", ctx->devs[k]));
    DeviceInfo info;
    if (int s = device_info(&info)) return bail(s);
    if (k == 0) {
      if (cudaEventCreateWithFlags(&ctx->fork, cudaEventDisableTiming) != cudaSuccess) return bail(fail(ICB_ERR_CUDA, "event creation failed"));
      continue;
    }
    if (cudaStreamCreateWithFlags(&ctx->streams[k], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->done[k], cudaEventDisableTiming) != cudaSuccess)
      return bail(fail(ICB_ERR_CUDA, "stream / event creation on device %d failed", ctx->devs[k]));
    int can = 0;
    cudaDeviceCanAccessPeer(&can, ctx->devs[k], ctx->devs[0]);
    if (can) {
      const cudaError_t e = cudaDeviceEnablePeerAccess(ctx->devs[0], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
      cudaGetLastError();
    }
    if (!can) ctx->peer_stores = false;
  }
  cudaSetDevice(home);
  *out = ctx;
  return ICB_OK;
}

int icb_ctx_destroy(icb_ctx *ctx) {
  if (!ctx) return ICB_OK;
  int home = 0;
  const bool have_home = cudaGetDevice(&home) == cudaSuccess;
  for (int k = 0; k < ctx->n; ++k) {
    if (cudaSetDevice(ctx->devs[k]) != cudaSuccess) continue;
    if (ctx->streams[k]) {
      cudaStreamSynchronize(ctx->streams[k]);
      cudaStreamDestroy(ctx->streams[k]);
    }
    if (ctx->done[k]) cudaEventDestroy(ctx->done[k]);
    if (k == 0 && ctx->fork) cudaEventDestroy(ctx->fork);
  }
  if (have_home) cudaSetDevice(home);
  cudaGetLastError();
  delete ctx;
  return ICB_OK;
}

int icb_ctx_device_count(const icb_ctx *ctx) { return ctx ? ctx->n : 0; }
int icb_ctx_device(const icb_ctx *ctx, int k) { return (ctx && k >= 0 && k < ctx->n) ? ctx->devs[k] : -1; }
int icb_ctx_peer_stores(const icb_ctx *ctx) { return ctx && ctx->peer_stores ? 1 : 0; }

int icb_stripe_partition(int n, uint32_t grid_rows, int root_share_permille, uint32_t *splits) {
  if (n < 1 || !splits) return fail(ICB_ERR_INVALID, "bad partition request");
  if (root_share_permille > 1000) root_share_permille = 1000;
  // Stripes are whole tile rows (4 block rows) wherever the grid allows it, so that only the last stripe can leave
  // rows to the generic kernel.
  const uint32_t unit = grid_rows >= 4u * static_cast<uint32_t>(n) ? 4u : 1u;
  const uint32_t units = grid_rows / unit;  // the remainder goes to the last stripe
  uint32_t root_units = units / n + (units % n ? 1u : 0u);  // even split: the first ranks take the extra unit
  if (n > 1 && root_share_permille >= 0) {
    const uint64_t want = (static_cast<uint64_t>(units) * root_share_permille + 500) / 1000;
    if (want > root_units) root_units = static_cast<uint32_t>(want);
    if (root_units > units) root_units = units;
  }
  splits[0] = 0;
  if (n == 1) {
    splits[1] = grid_rows;
    return ICB_OK;
  }
  const bool even = root_share_permille < 0 || root_units == units / n + (units % n ? 1u : 0u);
  if (even) {
    const uint32_t base = units / n, extra = units % n;
    for (int r = 0; r < n; ++r) splits[r + 1] = splits[r] + (base + (static_cast<uint32_t>(r) < extra ? 1u : 0u)) * unit;
  } else {
    splits[1] = root_units * unit;
    const uint32_t rest = units - root_units, base = rest / (n - 1), extra = rest % (n - 1);
    for (int r = 1; r < n; ++r) splits[r + 1] = splits[r] + (base + (static_cast<uint32_t>(r - 1) < extra ? 1u : 0u)) * unit;
  }
  splits[n] = grid_rows;
  return ICB_OK;
}

int icb_root_share_permille(int codec, int format_components, int n) {
  if (n <= 1) return 1000;
  // Root share f that equalises "root encodes f of the image" with "the other 1-f arrives over the root's NVLink
  // ingress": f = t_link / (t_link + t_kernel), both for the whole image.  Measured on B200:
  // peer stores from ONE sending GPU arrive at ~510 GB/s (its encode kernel is slowed by the remote stores), from
  // three or more senders the root's ingress saturates at ~720 GB/s; whole-image kernel times per output byte below.
  double t_kernel_per_out_byte;
  const double link_gbs = n == 2 ? 510.0 : (n == 3 ? 680.0 : 720.0), t_link_per_out_byte = 1.0 / (link_gbs * 1e9);
  switch (codec) {
    case ICB_CODEC_DXT1: t_kernel_per_out_byte = (format_components == 4 ? 49.0e-6 : 43.5e-6) / 33554432.0; break;
    case ICB_CODEC_DXT5: t_kernel_per_out_byte = 78e-6 / 67108864.0; break;
    case ICB_CODEC_ETC1: t_kernel_per_out_byte = 156e-6 / 8388608.0; break;
    default: return -1;
  }
  const double f = t_link_per_out_byte / (t_link_per_out_byte + t_kernel_per_out_byte);
  const int permille = static_cast<int>(f * 1000.0 + 0.5);
  const int even = (1000 + n - 1) / n;
  return permille > even ? permille : -1;  // -1: an even split is already compute-bound
}

int icb_encode_sharded(icb_ctx *ctx, int codec, int format, uint32_t h, uint32_t w, size_t src_pitch, int strategy,
                       const uint32_t *splits, const void *const *d_src_stripes, void *d_dst, void *stream) {
  if (!ctx || !splits || !d_src_stripes || !d_dst) return fail(ICB_ERR_INVALID, "null argument");
  if (codec < ICB_CODEC_DXT1 || codec > ICB_CODEC_ETC1) return fail(ICB_ERR_UNSUPPORTED, "sharded encode covers the 4x4 codecs (PVRTC: icb_pvrtc2_encode_stripe)");
  if (format < ICB_RGB || format > ICB_BGRA) return fail(ICB_ERR_INVALID, "unknown format %d", format);
  if (h == 0 || w == 0) return fail(ICB_ERR_INVALID, "zero image dimension");
  const int ncomp = (format == ICB_RGB || format == ICB_BGR) ? 3 : 4;
  const int swap_rb = (format == ICB_BGR || format == ICB_BGRA);
  if (codec == ICB_CODEC_DXT5 && ncomp != 4) return fail(ICB_ERR_INVALID, "DXT5 needs a 4-component format");
  if (codec == ICB_CODEC_ETC1 && format != ICB_RGB) return fail(ICB_ERR_INVALID, "ETC1 supports kRGB only");
  if (ctx->n > 1 && !ctx->peer_stores) return fail(ICB_ERR_UNSUPPORTED, "no peer access from every device of the context to device %d", ctx->devs[0]);
  const uint32_t grid_rows = (h + 3) / 4, grid_cols = (w + 3) / 4;
  if (splits[0] != 0 || splits[ctx->n] != grid_rows) return fail(ICB_ERR_INVALID, "splits must run from 0 to %u", grid_rows);
  for (int r = 0; r < ctx->n; ++r)
    if (splits[r] > splits[r + 1]) return fail(ICB_ERR_INVALID, "splits must be non-decreasing");
  const size_t bb = codec == ICB_CODEC_DXT5 ? 16 : 8;
  int home = 0;
  ICB_CUDA(cudaGetDevice(&home));
  cudaStream_t root_stream = static_cast<cudaStream_t>(stream);
  auto run = [&]() -> int {
    ICB_CUDA(cudaSetDevice(ctx->devs[0]));
    if (ctx->n > 1) ICB_CUDA(cudaEventRecord(ctx->fork, root_stream));
    // peers first: their kernels are the ones whose stores cross NVLink, so they should start earliest
    for (int k = 1; k <= ctx->n; ++k) {
      const int r = k % ctx->n;  // 1, 2, ..., n-1, 0
      const uint32_t r0 = splits[r], r1 = splits[r + 1];
      if (r1 == r0) continue;
      if (!d_src_stripes[r]) return fail(ICB_ERR_INVALID, "null stripe pointer for device %d", r);
      ICB_CUDA(cudaSetDevice(ctx->devs[r]));
      cudaStream_t st = r == 0 ? root_stream : ctx->streams[r];
      if (r != 0) ICB_CUDA(cudaStreamWaitEvent(st, ctx->fork, 0));
      // virtual address of pixel (0,0): rows above the stripe are never touched
      const uint8_t *base = static_cast<const uint8_t *>(d_src_stripes[r]) - static_cast<size_t>(r0) * 4 * src_pitch;
      uint8_t *out = static_cast<uint8_t *>(d_dst) + static_cast<size_t>(r0) * grid_cols * bb;
      if (int s = encode4x4(codec, ncomp, base, h, w, src_pitch, h, w, swap_rb, strategy, r0, r1, out, st)) return s;
      if (r != 0) ICB_CUDA(cudaEventRecord(ctx->done[r], st));
    }
    ICB_CUDA(cudaSetDevice(ctx->devs[0]));
    for (int r = 1; r < ctx->n; ++r)
      if (splits[r + 1] > splits[r]) ICB_CUDA(cudaStreamWaitEvent(root_stream, ctx->done[r], 0));
    return ICB_OK;
  };
  const int status = run();
  cudaSetDevice(home);
  return status;
}

int icb_ctx_compress_host(icb_ctx *ctx, int codec, int format, uint32_t h, uint32_t w, uint32_t padded_h,
                          uint32_t padded_w, uint32_t padding, int strategy, const void *src, void *dst, size_t dst_size) {
  if (!ctx) return fail(ICB_ERR_INVALID, "null context");
  return compress_host_impl(ctx->devs, ctx->n, codec, format, h, w, padded_h, padded_w, padding, strategy, src, dst, dst_size);
}

}  // extern "C"
