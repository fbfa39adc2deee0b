// Does a kernel that allocates tensor memory co-reside 2 CTAs per SM?  Occupancy API + a timing check.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_occ_probe tmem_occ_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int COLS, bool USE_TMEM>
__global__ void __launch_bounds__(192, 2) k(long long* out, int spin) {
  __shared__ uint32_t slot;
  extern __shared__ uint8_t dyn[];
  if (USE_TMEM && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  __syncthreads();
  if (USE_TMEM && threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(slot), "r"(COLS) : "memory");
  if (threadIdx.x == 0 && dyn[0] == 77) out[0] = 1;
}

template <int COLS, bool USE_TMEM>
void run(const char* name, int smem) {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k<COLS, USE_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<COLS, USE_TMEM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<COLS, USE_TMEM>, 192, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<COLS, USE_TMEM><<<148 * 2, 192, smem>>>(d, 1000);
  cudaEventRecord(e0);
  k<COLS, USE_TMEM><<<148 * 2, 192, smem>>>(d, 2000000);   // ~1 ms per CTA: 2 CTAs/SM -> ~1 ms total, 1 CTA/SM -> ~2 ms
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s smem %6d: occupancy API %d CTAs/SM, 296 CTAs x 2M cycles took %.2f ms (%s)\n", name, smem, nb, ms,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run<256, false>("no tmem", 32768);
  run<256, true>("tmem 256 cols", 32768);
  run<256, true>("tmem 256 cols", 108800);
  run<128, true>("tmem 128 cols", 108800);
  run<512, true>("tmem 512 cols", 32768);
  return 0;
}
