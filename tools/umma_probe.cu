// Microbenchmark: how long do chains of small tcgen05.mma instructions take?  (design input for attention_tc.cu)
// One CTA per SM; one thread issues `reps` rounds of a pattern of MMAs (M=128, N, K=16, bf16), either all into the
// same TMEM accumulator (dependent chain) or round-robin over several accumulators, with A from smem (SS) or TMEM (TS),
// then commits and waits.  Prints cycles per MMA.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o
// tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../avsiam_b200/csrc/common.cuh"
void avs_set_error(const char*, ...) {}
int avs_check_launch(const char*) { return 0; }

__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)lt << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n, int am, int bm) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)am << 15) | ((uint32_t)bm << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// mode: 0 = SS K-major A/B (SW128), 1 = TS (A in TMEM) with MN-major B (SW128), 2 = SS MN-major A and B (SW128)
//       3 = SS K-major SW64 (64-byte rows, 32 bytes per k-step)   4 = TS, B MN-major SW64 (64-byte rows)
//       5 = SS K-major SW32 (32-byte rows, one k-step per slab)   6 = TS, B MN-major SW32 (two 16-element atoms)
//       7 = SS: A = MN-major SW128 (dS tile), B = MN-major SW64   8 = like 7 with B MN-major SW32
template <int N, int NACC, int MODE>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t sa = smem_u32(smem), sb = sa + 32 * 1024;
    const uint64_t da = (MODE == 2 || MODE == 7 || MODE == 8) ? mk_desc(sa, 16384, 1024, 2)
                        : MODE == 3 ? mk_desc(sa, 0, 512, 4) : MODE == 5 ? mk_desc(sa, 0, 256, 6) : mk_desc(sa, 0, 1024, 2);
    const uint64_t db = MODE == 0 ? mk_desc(sb, 0, 1024, 2) : MODE == 3 ? mk_desc(sb, 0, 512, 4)
                        : MODE == 5 ? mk_desc(sb, 0, 256, 6) : (MODE == 4 || MODE == 7) ? mk_desc(sb, 512, 512, 4)
                        : (MODE == 6 || MODE == 8) ? mk_desc(sb, 4096, 256, 6) : mk_desc(sb, 1024, 1024, 2);
    constexpr bool A_MN = (MODE == 2 || MODE == 7 || MODE == 8);
    constexpr bool B_MN = !(MODE == 0 || MODE == 3 || MODE == 5);
    constexpr bool TS = (MODE == 1 || MODE == 4 || MODE == 6);
    constexpr uint32_t ID = idesc(128, N, A_MN ? 1 : 0, B_MN ? 1 : 0);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {  // first round warms up
      t0 = clock64();
      if (elect_one_sync()) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t acc = tmem + 256 + (uint32_t)((k % NACC) * N);
            // B k-step advance (16-byte units): MN-major SW128 16 rows x 128 B = 128; SW64 16 x 64 B = 64; SW32 32
            const uint64_t bk = (MODE == 4 || MODE == 7) ? (uint64_t)(k * 64) : (MODE == 6 || MODE == 8) ? (uint64_t)(k * 32)
                                                                                : (uint64_t)(k * 128);
            if (TS) umma_bf16_ts(acc, tmem + (uint32_t)(k * 8), db + bk, ID, 1u);
            else if (A_MN) umma_bf16_ss(acc, da + (uint64_t)(k * 128), db + bk, ID, 1u);
            else if (MODE == 3) umma_bf16_ss(acc, da + (uint64_t)(2 * (k & 1)), db + (uint64_t)(2 * (k & 1)), ID, 1u);
            else if (MODE == 5) umma_bf16_ss(acc, da + (uint64_t)(256 * (k & 1)), db + (uint64_t)(256 * (k & 1)), ID, 1u);
            else umma_bf16_ss(acc, da + (uint64_t)(2 * (k & 3)), db + (uint64_t)(2 * (k & 3)), ID, 1u);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, (uint32_t)rep);
      t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int NACC, int MODE>
void run(const char* name, long long* d_out) {
  const int reps = 64;
  cudaFuncSetAttribute(probe<N, NACC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<N, NACC, MODE><<<148, 128, 64 * 1024>>>(d_out, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d acc=%d : %7.1f cycles / MMA   (ideal %3d)  %s\n", name, N, NACC, (double)cyc / (reps * 8), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// The exact MMA batch of one 128x128 block-step of attn_bwd_tc_kernel<32> (2 sub-steps), issued back to back.
// spin: number of extra warps polling an mbarrier that never completes until the end (what idle role warps do).
__global__ void __launch_bounds__(576, 1) probe_batch(long long* out, int reps, int spin, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, never;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 576) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    mbar_init(&bar, 1);
    mbar_init(&never, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t sQ = smem_u32(smem), sDO = sQ + 49152, sK = sDO + 49152, sV = sK + 16384, sDS = sV + 16384;
    const uint64_t kK = mk_desc(sK, 0, 512, 4), kV = mk_desc(sV, 0, 512, 4), kQ = mk_desc(sQ, 0, 512, 4),
                   kDO = mk_desc(sDO, 0, 512, 4);
    const uint64_t mQ = mk_desc(sQ, 512, 512, 4), mDO = mk_desc(sDO, 512, 512, 4), mK = mk_desc(sK, 512, 512, 4);
    const uint64_t aDS = mk_desc(sDS, 16384, 1024, 2);
    constexpr uint32_t ID_S = idesc(128, 64, 0, 0), ID_TS = idesc(128, 32, 0, 1), ID_DQ = idesc(128, 32, 1, 1);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one_sync()) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t b = hh, qo = (uint32_t)hh * 256;
            if (variant != 2) {
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_bf16_ss(tmem + b * 64, kK + (uint64_t)(2 * k), kQ + (uint64_t)(qo + 2 * k), ID_S, k);
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_bf16_ss(tmem + 128 + b * 64, kV + (uint64_t)(2 * k), kDO + (uint64_t)(qo + 2 * k), ID_S, k);
            }
            if (variant != 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem + 288, tmem + b * 64 + k * 16, mDO + (uint64_t)(qo + k * 64), ID_TS, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem + 256, tmem + 128 + b * 64 + k * 16, mQ + (uint64_t)(qo + k * 64), ID_TS, 1u);
            }
            if (variant == 1) umma_commit(&never);
          }
          if (variant != 2 && variant != 3) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_bf16_ss(tmem + 320, aDS + (uint64_t)(k * 128), mK + (uint64_t)(k * 64), ID_DQ, 1u);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, (uint32_t)rep);
      t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (lane == 0) mbar_arrive(&never);
  } else if (warp >= 2 && warp < 2 + spin && variant != 1) {
    mbar_wait(&never, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

void run_batch(const char* name, long long* d_out, int spin, int variant) {
  const int reps = 32;
  cudaFuncSetAttribute(probe_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe_batch<<<148, 576, 200 * 1024>>>(d_out, reps, spin, variant);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("%-44s spin warps=%2d : %7.1f cycles / block-step (model 972)  %s\n", name, spin, (double)cyc / reps,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// Same batch, but issued the way the kernel does: per sub-step [S^T,dP^T -> commit s_full] ... [wait p_full] [dV,dK(,dQ)].
// variant 4: only tcgen05.fence::after_thread_sync between the groups; variant 5: real handshake with an instant responder
__global__ void __launch_bounds__(576, 1) probe_hs(long long* out, int reps, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, s_full[2], p_full[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 576) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 1); }
    fence_mbar_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int T = reps * 2;
  if (warp == 1) {
    const uint32_t sQ = smem_u32(smem), sDO = sQ + 49152, sK = sDO + 49152, sV = sK + 16384, sDS = sV + 16384;
    const uint64_t kK = mk_desc(sK, 0, 512, 4), kV = mk_desc(sV, 0, 512, 4), kQ = mk_desc(sQ, 0, 512, 4),
                   kDO = mk_desc(sDO, 0, 512, 4);
    const uint64_t mQ = mk_desc(sQ, 512, 512, 4), mDO = mk_desc(sDO, 512, 512, 4), mK = mk_desc(sK, 512, 512, 4);
    const uint64_t aDS = mk_desc(sDS, 16384, 1024, 2);
    constexpr uint32_t ID_S = idesc(128, 64, 0, 0), ID_TS = idesc(128, 32, 0, 1), ID_DQ = idesc(128, 32, 1, 1);
    const long long t0 = clock64();
    for (int t = 0; t <= T; ++t) {
      const uint32_t b = t & 1;
      if (t < T) {
        const uint32_t qo = b * 256;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16_ss(tmem + b * 64, kK + (uint64_t)(2 * k), kQ + (uint64_t)(qo + 2 * k), ID_S, k);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16_ss(tmem + 128 + b * 64, kV + (uint64_t)(2 * k), kDO + (uint64_t)(qo + 2 * k), ID_S, k);
          umma_commit(&s_full[b]);
        }
        __syncwarp();
      }
      if (t >= 1) {
        const uint32_t pb = b ^ 1, qo = pb * 256;
        const int u = t - 1;
        if (variant == 5) mbar_wait(&p_full[pb], (uint32_t)((u >> 1) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem + 288, tmem + pb * 64 + k * 16, mDO + (uint64_t)(qo + k * 64), ID_TS, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem + 256, tmem + 128 + pb * 64 + k * 16, mQ + (uint64_t)(qo + k * 64), ID_TS, 1u);
          if (pb == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_bf16_ss(tmem + 320, aDS + (uint64_t)(k * 128), mK + (uint64_t)(k * 64), ID_DQ, 1u);
          }
        }
        __syncwarp();
      }
    }
    if (elect_one_sync()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if ((warp == 2 || warp == 3) && variant == 5) {
    const int g = warp - 2;
    for (int n = 0; n < reps; ++n) {
      mbar_wait(&s_full[g], (uint32_t)(n & 1));
      tc_fence_after();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

void run_hs(const char* name, long long* d_out, int variant) {
  const int reps = 32;
  cudaFuncSetAttribute(probe_hs, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe_hs<<<148, 576, 200 * 1024>>>(d_out, reps, variant);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("%-60s : %7.1f cycles / block-step   %s\n", name, (double)cyc / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// N = 256 SS MMAs (the GEMM mainloop) with and without concurrent shared-memory fill traffic from other warps
// (stands in for the TMA writes of the next stages: 12 KB per MMA in the real kernel).
__global__ void __launch_bounds__(320, 1) probe_n256(long long* out, int reps, int writers, const void* gsrc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 320) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    stop = 0;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t sa = smem_u32(smem), sb = sa + 16 * 1024;
    const uint64_t da = mk_desc(sa, 0, 1024, 2), db = mk_desc(sb, 0, 1024, 2);
    constexpr uint32_t ID = idesc(128, 256, 0, 0);
    const long long t0 = clock64();
    if (elect_one_sync()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem + (r & 1) * 256, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), ID, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    stop = 1;
  } else if (warp == 2 && writers < 0) {
    // TMA-engine writes: 1-D bulk copies global -> shared (16 KB each, 4 in flight), like the producer's stage fills
    __shared__ uint64_t wbar[4];
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) mbar_init(&wbar[i], 1);
      fence_mbar_init();
      long long n = 0;
      uint32_t ph[4] = {0, 0, 0, 0};
      const uint8_t* g = reinterpret_cast<const uint8_t*>(gsrc) + (size_t)blockIdx.x * 65536;
      for (int i = 0; i < 4; ++i) {
        mbar_arrive_expect_tx(&wbar[i], 16384);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         smem_u32(smem) + 64 * 1024 + i * 16384), "l"(g + i * 16384), "r"(16384), "r"(smem_u32(&wbar[i])) : "memory");
      }
      while (!stop) {
        for (int i = 0; i < 4; ++i) {
          mbar_wait(&wbar[i], ph[i]);
          ph[i] ^= 1;
          mbar_arrive_expect_tx(&wbar[i], 16384);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                           smem_u32(smem) + 64 * 1024 + i * 16384), "l"(g + i * 16384), "r"(16384), "r"(smem_u32(&wbar[i])) : "memory");
          ++n;
        }
      }
      for (int i = 0; i < 4; ++i) mbar_wait(&wbar[i], ph[i]);
      if (blockIdx.x == 0) out[1] = n * 32;   // in units of 512 B like the st.shared writers
    }
  } else if (warp >= 2 && writers >= 100 && warp < 2 + (writers - 100)) {
    // epilogue-like readers: tcgen05.ld 32x32b.x32 + wait in a loop while the MMA stream runs (writers = 100 + #warps)
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    long long n = 0, t_ld = 0;
    uint32_t acc = 0;
    while (!stop) {
      uint32_t v[32];
      const long long a0 = clock64();
      tmem_ld_32x32b_x32(tl + (uint32_t)((n * 32) & 255), v);
      tmem_ld_wait();
      t_ld += clock64() - a0;
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
      ++n;
    }
    if (acc == 0x12345u) out[15] = acc;
    if (lane == 0 && blockIdx.x == 0 && warp == 2) { out[1] = n; out[2] = t_ld; }
  } else if (warp >= 2 && warp < 2 + writers) {
    // each warp streams 512 B per instruction into a private 16 KB region (different from the operand tiles)
    const uint32_t base = smem_u32(smem) + 64 * 1024 + (warp - 2) * 16 * 1024 + lane * 16;
    long long n = 0;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};\n" ::"r"(base + i * 512), "r"(i) : "memory");
      n += 32;
    }
    if (lane == 0 && blockIdx.x == 0) out[1 + (warp - 2)] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
void run_n256(long long* d_out, int writers) {
  const int reps = 2048;
  cudaMemset(d_out, 0, 128);
  cudaFuncSetAttribute(probe_n256, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  static void* gsrc = nullptr;
  if (!gsrc) { cudaMalloc(&gsrc, 148 * 65536); cudaMemset(gsrc, 0, 148 * 65536); }
  probe_n256<<<148, 320, 200 * 1024>>>(d_out, reps, writers, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, d_out, 128, cudaMemcpyDeviceToHost);
  if (writers >= 100) {
    printf("SS N=256 K-major with %d warps looping tcgen05.ld.x32: %6.1f cycles / MMA (math 128); tcgen05.ld+wait = %6.1f cycles each  %s\n",
           writers - 100, (double)h[0] / (reps * 4), h[1] ? (double)h[2] / (double)h[1] : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    return;
  }
  long long stores = 0;
  for (int i = 0; i < (writers < 0 ? 1 : writers); ++i) stores += h[1 + i];
  printf("SS N=256 K-major, %d writer warps: %6.1f cycles / MMA (math 128); concurrent st.shared traffic %5.1f B/cycle  %s\n",
         writers, (double)h[0] / (reps * 4), (double)stores * 512 / (double)h[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// tcgen05.ld throughput: `nw` warps each read their lane quarter, 32 fp32 columns per instruction, `reps` times.
__global__ void __launch_bounds__(512, 1) probe_ldtm(long long* out, int reps, int nw, int batch) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  uint32_t acc = 0;
  const long long t0 = clock64();
  if (warp < nw) {
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int r = 0; r < reps; ++r) {
      uint32_t v[32], w[32];
      tmem_ld_32x32b_x32(tl + ((r * 64) & 255), v);
      if (batch > 1) tmem_ld_32x32b_x32(tl + ((r * 64 + 32) & 255), w);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
      if (batch > 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= w[i];
      }
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[3] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
void run_ldtm(long long* d_out, int nw, int batch) {
  const int reps = 1024;
  probe_ldtm<<<148, 512, 0>>>(d_out, reps, nw, batch);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)nw * reps * batch * 32 * 32 * 4;
  printf("tcgen05.ld 32x32b.x32: %2d warps, %d loads per wait: %6.1f B/cycle/SM, %6.1f cycles per load  %s\n", nw, batch,
         bytes / cyc, (double)cyc / (reps * batch), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 256);
  run<32, 1, 0>("SS K-major, one accumulator", d_out);
  run<32, 2, 0>("SS K-major, 2 accumulators", d_out);
  run<32, 4, 0>("SS K-major, 4 accumulators", d_out);
  run<64, 1, 0>("SS K-major, one accumulator", d_out);
  run<64, 2, 0>("SS K-major, 2 accumulators", d_out);
  run<128, 1, 0>("SS K-major, one accumulator", d_out);
  run<32, 1, 1>("TS (A in TMEM), B MN-major", d_out);
  run<32, 2, 1>("TS (A in TMEM), B MN-major", d_out);
  run<32, 4, 1>("TS (A in TMEM), B MN-major", d_out);
  run<64, 1, 1>("TS (A in TMEM), B MN-major", d_out);
  run<64, 2, 1>("TS (A in TMEM), B MN-major", d_out);
  run<32, 1, 2>("SS MN-major A and B", d_out);
  run<32, 2, 2>("SS MN-major A and B", d_out);
  run<32, 4, 2>("SS MN-major A and B", d_out);
  run<64, 1, 2>("SS MN-major A and B", d_out);
  run<64, 2, 2>("SS MN-major A and B", d_out);
  run<64, 1, 3>("SS K-major SW64 (S^T today)", d_out);
  run<64, 1, 5>("SS K-major SW32 slabs", d_out);
  run<32, 1, 4>("TS, B MN-major SW64 (dV today)", d_out);
  run<32, 1, 6>("TS, B MN-major SW32", d_out);
  run<32, 1, 7>("SS A=dS SW128 MN, B SW64 MN (dQ)", d_out);
  run<32, 1, 8>("SS A=dS SW128 MN, B SW32 MN", d_out);
  run_batch("attention batch", d_out, 0, 0);
  run_batch("attention batch", d_out, 8, 0);
  run_batch("attention batch", d_out, 16, 0);
  run_batch("attention batch + commit per sub-step", d_out, 0, 1);
  run_batch("only dV/dK (TS) [model 262]", d_out, 0, 2);
  run_batch("only S^T/dP^T (SS N=64) [model 387]", d_out, 0, 3);
  run_hs("kernel-order issue, fences only", d_out, 4);
  run_hs("kernel-order issue, handshake with instant responder", d_out, 5);
  for (int w : {0, 4, -1, 104, 108}) run_n256(d_out, w);
  for (int nw : {1, 4, 8, 16}) for (int b : {1, 2}) run_ldtm(d_out, nw, b);
  return 0;
}
