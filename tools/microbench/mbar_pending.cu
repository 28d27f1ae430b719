// What does mbarrier.pending_count say about the state mbarrier.arrive returns -- the count BEFORE or AFTER the arrival?
// (The ring kernel's "last warp refills the slot" test depends on it.)  One thread, a barrier of 4 arrivals, 9 arrivals;
// prints the pending count of every returned state.  nvcc -arch=sm_100a -o mbar_pending mbar_pending.cu && ./mbar_pending
#include <cstdint>
#include <cstdio>
__global__ void probe(uint32_t *out) {
  __shared__ uint64_t bar;
  const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(&bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(4) : "memory");
  for (int i = 0; i < 9; ++i) {
    uint32_t cnt;
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%1];\nmbarrier.pending_count.b64 %0, st;\n}" : "=r"(cnt) : "r"(b) : "memory");
    out[i] = cnt;
  }
}
int main() {
  uint32_t *d, h[9];
  cudaMalloc(&d, sizeof(h));
  probe<<<1, 1>>>(d);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("pending counts seen by 9 arrivals on a barrier of 4:");
  for (int i = 0; i < 9; ++i) printf(" %u", h[i]);
  printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
