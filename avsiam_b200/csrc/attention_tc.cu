// tcgen05 / TMEM attention backward for the AVSiam shapes (MAE decoder: S = 708, head_dim 32; encoder / fusion:
// S <= 256, head_dim 64).  One CTA per (sequence, head) computes dQ, dK and dV in ONE pass (the mma.sync path in
// attention.cu needs two kernels that each recompute P), with every product on the 5th-generation tensor cores.
//
// Replaces the backward of F.scaled_dot_product_attention in Attention.forward (cav_mae_base.py:58-77).
//
// Structure (per CTA, 19 warps):
//   warp 0      loader: streams the 128-key blocks K_j, V_j (double-buffered) into swizzled shared memory
//   warps 1, 2  MMA issuers (PV stream: dV, dK, dQ / QK stream: S^T, dP^T), one elected thread each
//   warps 3-18  "softmax" warps in two ping-pong groups of 8 (one per TMEM buffer): thread = one key row (TMEM
//               lane), the two warps of a lane quarter split the sub-step's 64 query columns
// Shared memory: Q and dO of the whole head stay resident (rows of head_dim bf16, SWIZZLE_64B / SWIZZLE_128B so the
// same bytes serve as K-major operand of S^T = K Q^T and as MN-major operand of dK = dS^T Q), K_j / V_j blocks,
// the dS^T tile (double-buffered, MN-major A operand of dQ = dS K), log-sum-exp and delta rows.
// Tensor memory (512 columns, all used):
//   [0,128)    two 64-column buffers  S^T[kv, q]   (fp32), one per softmax group
//   [128,256)  two 64-column buffers  dP^T[kv, q]  (fp32) — once read, overwritten in place by P^T and dS^T (bf16 pairs,
//              the TMEM A operands of dV and dK)
//   [256, ..)  dK_j, dV_j accumulators (head_dim columns each), then dQ_i for EVERY query block of the head
// so dQ is accumulated across the key blocks without atomics and without a second pass.
// Pipeline (measured, tools/umma_probe.cu + the AVS_TC_TRACE timeline): the tensor pipe needs ~1030 cycles per
// 128x128 block (SS MMAs cost 32 + N/4 cycles of operand reads, TS MMAs N/2), the exponentials 1024 MUFU cycles.
// What bounds the kernel is the round trip softmax -> MMA -> softmax, so the S^T product of sub-step t+2 is issued
// as soon as the softmax warps have READ S^T of sub-step t (it does not wait for the dV/dK/dQ products of t), and
// the exponentials of t+2 overlap the tensor work that consumes t.
#include "../../include/avsiam_b200.h"
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int TCB_GROUP_WARPS = 8;                 // softmax warps per ping-pong group (4 lane quarters x 2 column halves)
constexpr int TCB_COMPUTE_WARPS = 2 * TCB_GROUP_WARPS;
constexpr int TCB_FIRST_SOFTMAX_WARP = 4;           // warp 0 loader, warp 1 PV issuer, warp 2 QK issuer, warp 3 dQ issuer
constexpr int TCB_THREADS = 32 * (TCB_FIRST_SOFTMAX_WARP + TCB_COMPUTE_WARPS);
constexpr int DS_TILE_BYTES = 128 * 128 * 2;  // [128 keys][128 queries] bf16 = two SWIZZLE_128B atoms of 64 queries

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// generic shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x1(uint32_t taddr, uint32_t (&r)[1]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r[0]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x1(uint32_t taddr, const uint32_t (&r)[1]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(taddr), "r"(r[0]) : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 32) tmem_ld_32x32b_x32(taddr, r);
  else if constexpr (N == 16) tmem_ld_32x32b_x16(taddr, r);
  else tmem_ld_32x32b_x8(taddr, r);
}
template <int N>
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t (&r)[N]) {
  if constexpr (N == 16) tmem_st_32x32b_x16(taddr, r);
  else tmem_st_32x32b_x8(taddr, r);
}

__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;\n" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// 2^x on the FMA / ALU pipes for -126 < x <= 0 (Cody-Waite split + degree-3 polynomial, relative error 7.7e-5 — well
// below the bf16 rounding of P): x = xi + xf with xi = round(x) taken from the low mantissa bits of x + 1.5 * 2^23,
// 2^xf from the polynomial, and xi added straight into the exponent field.  8 instructions instead of one MUFU.EX2:
// used for a fraction of the elements, because MUFU issues only 16 lanes per clock per SM.
__device__ __forceinline__ float exp2_poly(float x) {
  const float t = x + 12582912.0f;
  const float xf = x - (t - 12582912.0f);
  float p = 0.05508868396282196f;
  p = fmaf(p, xf, 0.24260404706001282f);
  p = fmaf(p, xf, 0.6932762265205383f);
  p = fmaf(p, xf, 0.9999289512634277f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
constexpr float TCF_LAZY = 8.0f;   // log2 of the largest probability tolerated before the running maximum is raised
template <int HD>
struct TcCfg {
  static constexpr int ROWB = HD * 2;                    // bytes per row of a Q/K/V/dO tile
  static constexpr uint32_t LT = (HD == 64) ? 2u : 4u;   // SWIZZLE_128B : SWIZZLE_64B
  static constexpr int SBO = 8 * ROWB;                   // 8-row group pitch
  static constexpr int KSTEPS = HD / 16;                 // k-steps of the head_dim contraction
  static constexpr int BLK_BYTES = 128 * ROWB;           // one 128-row block
  static constexpr int COL_S = 0, COL_DP = 128, COL_DK = 256, COL_DV = 256 + HD, COL_DQ = 256 + 2 * HD;
  static constexpr int MAX_NB = (512 - COL_DQ) / HD;     // query blocks whose dQ fits in TMEM: 6 (hd 32) / 2 (hd 64)
};

// byte offset of 16-byte chunk `chunk` of row `row` in a swizzled [rows][HD] bf16 tile (tile base 1024-aligned)
template <int HD>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  const int sw = (HD == 64) ? (row & 7) : ((row >> 1) & 3);
  return (uint32_t)(row * (HD * 2) + ((chunk ^ sw) << 4));
}

// rows [row0, row0+nrows) of a [*, ld] bf16 matrix (HD columns at src) -> swizzled tile rows [0, nrows);
// rows at or beyond S are zero-filled.  `tid`/`nthr` = the cooperating threads.
template <int HD>
__device__ __forceinline__ void load_rows_async(uint32_t tile, const bf16* __restrict__ src, long long ld, int row0,
                                                int nrows, int S, int tid, int nthr) {
  constexpr int CPR = HD / 8;
  for (int i = tid; i < nrows * CPR; i += nthr) {
    const int r = i / CPR, c = i % CPR;
    const uint32_t dst = tile + tile_off<HD>(r, c);
    if (row0 + r < S) cp_async16(dst, src + (long long)(row0 + r) * ld + c * 8);
    else st_shared_v4(dst, 0u, 0u, 0u, 0u);
  }
}

#ifdef AVS_TC_TRACE
#define TC_TRACE(slot, idx) do { if (a.trace && blockIdx.x == 0 && blockIdx.y == 0) a.trace[(slot) * 512 + (idx)] = clock64(); } while (0)
#else
#define TC_TRACE(slot, idx) do {} while (0)
#endif

struct TcArgs {
  CUtensorMap map_q, map_do;   // [rows, 3D] / [rows, D] bf16, boxes of 128 rows x head_dim columns (Q_i / dO_i blocks i >= 1)
  long long* trace;   // debug timeline (AVS_TC_TRACE builds only)
  const bf16* qkv;
  const bf16* dout;
  const float* lse2;
  const float* delta;
  bf16* dqkv;
  float* dbias;       // optional fp32 [3*D]: += column sums of dQ | dK | dV (gradient of the qkv bias)
  long long ld_qkv, ld_o;
  int S, NB, H, D;
  float scale, scale_log2;
};

template <int HD>
__global__ void __launch_bounds__(TCB_THREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ TcArgs a) {
  using C = TcCfg<HD>;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  const int NB = a.NB, S = a.S;
  const int S_pad = NB * 128;
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sDO = sQ + NB * C::BLK_BYTES;
  const uint32_t sK = sDO + NB * C::BLK_BYTES;   // [2]
  const uint32_t sV = sK + 2 * C::BLK_BYTES;     // [2]
  const uint32_t sDS = sV + 2 * C::BLK_BYTES;    // [2]
  uint8_t* tail = smem + 2 * NB * C::BLK_BYTES + 4 * C::BLK_BYTES + 2 * DS_TILE_BYTES;
  float* s_lse = reinterpret_cast<float*>(tail);
  float* s_delta = s_lse + S_pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_delta + S_pad);
  uint64_t* kv_full = bars;        // [2] loader -> MMA
  uint64_t* kv_empty = bars + 2;   // [2] MMA -> loader
  uint64_t* s_full = bars + 4;     // [2] MMA -> softmax warps: S^T / dP^T of a sub-step are in TMEM
  uint64_t* p_full = bars + 6;     // [2] softmax warps -> MMA: P^T / dS^T written (TMEM + smem)
  uint64_t* ds_empty = bars + 8;   // [2] MMA -> softmax warps: dS^T smem tile consumed by the dQ product
  uint64_t* dkv_full = bars + 10;  // MMA -> softmax warps: dK_j, dV_j complete
  uint64_t* dkv_empty = bars + 11; // softmax warps -> MMA: dK_j, dV_j read out
  uint64_t* s_read = bars + 12;    // [2] softmax warps -> MMA: S^T of a sub-step is in registers, buffer reusable
  uint64_t* dp_full = bars + 14;   // [2] MMA -> softmax warps: dP^T of a sub-step is in TMEM
  uint64_t* dq_done = bars + 16;   // dQ issuer -> softmax warps: every dQ product of the head has retired
  uint64_t* pq_full = bars + 17;   // [2] all 16 softmax warps -> dQ issuer: both halves of dS^T tile n & 1 are written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  float* s_db = reinterpret_cast<float*>(bars + 20);   // [3 * HD] column sums of dQ | dK | dV of this head
  uint64_t* q_full = reinterpret_cast<uint64_t*>(s_db + 3 * 64);   // [8] TMA -> MMA issuers: Q_i / dO_i (i >= 1) landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, seq = blockIdx.y;
  if (threadIdx.x == 0) TC_TRACE(17, 0);
  const long long row_base = (long long)seq * S;
  const bf16* qb = a.qkv + row_base * a.ld_qkv + h * HD;
  const bf16* dob = a.dout + row_base * a.ld_o + h * HD;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 3);      // the three MMA issuers commit
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], TCB_GROUP_WARPS);
      mbar_init(&ds_empty[i], 1);
      mbar_init(&s_read[i], TCB_GROUP_WARPS);
      mbar_init(&dp_full[i], 1);
    }
    mbar_init(dq_done, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&q_full[i], 1);
    mbar_init(&pq_full[0], TCB_COMPUTE_WARPS);
    mbar_init(&pq_full[1], TCB_COMPUTE_WARPS);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_empty, TCB_GROUP_WARPS);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x < 3 * HD) s_db[threadIdx.x] = 0.f;
  // The first 128-query block of Q and dO, log-sum-exp and delta of the whole head come in with the cooperative load; the
  // other Q / dO blocks arrive by TMA (one box each, issued by one thread right after the prologue barrier) while the
  // first key block's sub-steps already run: the whole-head load was a 10 k-cycle prologue per CTA with every pipe idle,
  // 11 % of the CTA's life (AVS_TC_TRACE). A box that reaches past the sequence brings the next sequence's rows (or zero
  // fill past the tensor); those query columns have lse = +inf, i.e. P = dS = 0 exactly.
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.map_q);
    tma_prefetch_desc(&a.map_do);
  }
  load_rows_async<HD>(sQ, qb, a.ld_qkv, 0, 128, S, threadIdx.x, TCB_THREADS);
  load_rows_async<HD>(sDO, dob, a.ld_o, 0, 128, S, threadIdx.x, TCB_THREADS);
  // the first key block comes in with the same cooperative load (one exposed HBM round trip per CTA, not two)
  load_rows_async<HD>(sK, qb + a.D, a.ld_qkv, 0, 128, S, threadIdx.x, TCB_THREADS);
  load_rows_async<HD>(sV, qb + 2 * a.D, a.ld_qkv, 0, 128, S, threadIdx.x, TCB_THREADS);
  {
    const long long sb = ((long long)seq * a.H + h) * S;
    for (int i = threadIdx.x; i < S_pad; i += TCB_THREADS) {
      s_lse[i] = i < S ? a.lse2[sb + i] : INFINITY;  // padded queries: p = exp2(-inf) = 0
      s_delta[i] = i < S ? a.delta[sb + i] : 0.f;
    }
  }
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) TC_TRACE(17, 1);
  const int T = NB * NB * 2;  // 64-query sub-steps

  if (warp == 0) {
    // ============================ loader: K_j, V_j ============================
    if (lane == 0) {
      mbar_arrive(&kv_full[0]);   // block 0 was loaded (and fenced) by the whole CTA above
      for (int i = 1; i < NB; ++i) {
        mbar_arrive_expect_tx(&q_full[i], 2 * C::BLK_BYTES);
        tma_load_2d(smem + i * C::BLK_BYTES, &a.map_q, &q_full[i], h * HD, (int)row_base + i * 128);
        tma_load_2d(smem + (NB + i) * C::BLK_BYTES, &a.map_do, &q_full[i], h * HD, (int)row_base + i * 128);
      }
    }
    __syncwarp();
    for (int j = 1; j < NB; ++j) {
      if (j >= 2) mbar_wait(&kv_empty[j & 1], (uint32_t)(((j >> 1) - 1) & 1));
      load_rows_async<HD>(sK + (j & 1) * C::BLK_BYTES, qb + a.D, a.ld_qkv, j * 128, 128, S, lane, 32);
      load_rows_async<HD>(sV + (j & 1) * C::BLK_BYTES, qb + 2 * a.D, a.ld_qkv, j * 128, 128, S, lane, 32);
      cp_async_wait_all();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&kv_full[j & 1]);
    }
  } else if (warp >= 1 && warp <= 3) {
    // ============================ MMA issuers ============================
    // The MMAs of this kernel are small (16-48 tensor cycles each), so ONE issuing warp's scalar work (descriptor
    // adds, barrier waits: ~4 cycles per dependent instruction) would bound the kernel.  Two warps issue
    // independent streams, each visiting its barriers in the order the events happen:
    //   warp 2 "QK":  S^T(u)  as soon as S^T(u-2) has been read; dP^T(u-2) as soon as dV/dK(u-4) have consumed the
    //                 P^T / dS^T that live in that buffer
    //   warp 1 "PV":  dV(u), dK(u) (A operands in TMEM) and, every second sub-step, dQ
    // Both run convergently (descriptors and TMEM addresses stay in uniform registers); one elected lane issues.
    constexpr uint32_t ID_S = idesc_bf16(128, 64, 0, 0);    // S^T / dP^T : A, B K-major
    constexpr uint32_t ID_TS = idesc_bf16(128, HD, 0, 1);   // dV / dK    : A in TMEM, B MN-major
    constexpr uint32_t ID_DQ = idesc_bf16(128, HD, 1, 1);   // dQ         : A (dS^T tile) and B (K_j) MN-major
    constexpr uint32_t BLK16 = C::BLK_BYTES >> 4, HALF16 = (64 * C::ROWB) >> 4, K16ROWS = (16 * C::ROWB) >> 4;
    auto advance = [&](int& j, int& i, int& hh) {
      if (++hh == 2) {
        hh = 0;
        if (++i == NB) {
          i = 0;
          ++j;
        }
      }
    };
    if (warp == 2) {
      // "QK" issuer: S^T(u) = K_j Q_half^T into buffer u & 1 as soon as the softmax warps have read S^T(u-2)
      const uint64_t kK = make_desc(sK, 0, C::SBO, C::LT);
      const uint64_t kQ = make_desc(sQ, 0, C::SBO, C::LT);
      auto issue_s = [&](uint32_t g, int j, int i, int hh) {
        const uint32_t kvo = (uint32_t)(j & 1) * BLK16, qo = (uint32_t)(i * 2 + hh) * HALF16;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < C::KSTEPS; ++k)
            umma_bf16_ss(tmem + C::COL_S + g * 64, kK + (kvo + 2 * k), kQ + (qo + 2 * k), ID_S, k > 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
          if (i == NB - 1 && hh == 1) umma_commit(&kv_empty[j & 1]);   // last read of K_j by this warp
        }
        __syncwarp();
      };
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0, 0, 0);
      issue_s(1, 0, 0, 1);
      int sj = 0, si = 0, sh = 0;   // position of sub-step u
      advance(sj, si, sh); advance(sj, si, sh);
      for (int u = 2; u < T; ++u) {
        const uint32_t g = (uint32_t)(u & 1);
        if (si == 0 && sh == 0) {
          mbar_wait(&kv_full[sj & 1], (uint32_t)((sj >> 1) & 1));
          tc_fence_after();
        }
        if (sj == 0 && sh == 0 && si > 0) {   // first touch of Q block si
          mbar_wait(&q_full[si], 0);
          tc_fence_after();
        }
        mbar_wait(&s_read[g], (uint32_t)(((u - 2) >> 1) & 1));
        tc_fence_after();
        issue_s(g, sj, si, sh);
        advance(sj, si, sh);
      }
    } else if (warp == 3) {
      // "dQ" issuer: dQ_i += dS K_j once both halves of a 128-query block have their dS^T in shared memory. A warp of its
      // own: the PV issuer's instruction stream (10 small MMAs per sub-step, ~50 issue cycles each) is what paces the
      // kernel (AVS_TC_TRACE), and these 8 MMAs per pair of sub-steps need nothing from it.
      const uint64_t mK = make_desc(sK, C::SBO, C::SBO, C::LT);
      const uint64_t aDS = make_desc(sDS, 16384, 1024, 2u);
      int n = 0;
      for (int j = 0; j < NB; ++j) {
        mbar_wait(&kv_full[j & 1], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        for (int i = 0; i < NB; ++i, ++n) {
          // one barrier per tile buffer, completed by both softmax groups: its next phase needs this warp's ds_empty
          // commit first, so a waiter can never fall two phases behind (the per-group p_full barriers can)
          mbar_wait(&pq_full[n & 1], (uint32_t)((n >> 1) & 1));
          // (holding these MMAs back until the PV issuer's dP^T has retired was measured: 1.67 -> 1.75 ms — the queue is
          //  not what delays that chain)
          tc_fence_after();
          const uint32_t dso = (uint32_t)(n & 1) * (DS_TILE_BYTES >> 4);
          const uint32_t kvo = (uint32_t)(j & 1) * BLK16;
          if (elect_one_sync()) {
            // A = dS^T tile [128 keys][128 queries] read MN-major (two 64-query atoms 16 KB apart)
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_ss(tmem + C::COL_DQ + i * HD, aDS + (dso + k * 128), mK + (kvo + k * K16ROWS), ID_DQ,
                           (j > 0 || k > 0) ? 1u : 0u);
            umma_commit(&ds_empty[n & 1]);
            if (i == NB - 1) umma_commit(&kv_empty[j & 1]);   // last read of K_j by this warp
            if (i == NB - 1 && j == NB - 1) umma_commit(dq_done);
          }
          __syncwarp();
        }
      }
    } else {
      const uint64_t mQ = make_desc(sQ, C::SBO, C::SBO, C::LT), mDO = make_desc(sDO, C::SBO, C::SBO, C::LT);
      // dP^T(u) = V_j dO_half^T lands in the buffer that holds P^T / dS^T of sub-step u-2, so it must follow dV / dK(u-2):
      // issued by THIS warp directly behind them (the tensor pipe runs one thread's MMAs in order) instead of by the QK
      // warp after a commit -> barrier -> poll round trip — the softmax groups wait on exactly this chain
      // (AVS_TC_TRACE: 1350 cycles from S^T to dP^T per sub-step, 500 of compute).
      const uint64_t kV = make_desc(sV, 0, C::SBO, C::LT), kDO = make_desc(sDO, 0, C::SBO, C::LT);
      auto issue_dp = [&](uint32_t g, int j, int i, int hh) {
        const uint32_t kvo = (uint32_t)(j & 1) * BLK16, qo = (uint32_t)(i * 2 + hh) * HALF16;
#pragma unroll
        for (int k = 0; k < C::KSTEPS; ++k)
          umma_bf16_ss(tmem + C::COL_DP + g * 64, kV + (kvo + 2 * k), kDO + (qo + 2 * k), ID_S, k > 0 ? 1u : 0u);
        umma_commit(&dp_full[g]);
      };
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      if (elect_one_sync()) {
        issue_dp(0, 0, 0, 0);
        issue_dp(1, 0, 0, 1);
      }
      __syncwarp();
      int dj = 0, di = 0, dh = 0;   // position of sub-step u + 2 (dP^T stream)
      advance(dj, di, dh); advance(dj, di, dh);
      int cj = 0, ci = 0, ch = 0;
      for (int u = 0; u < T; ++u) {
        const uint32_t g = (uint32_t)(u & 1);
        if (lane == 0) TC_TRACE(1, u);
        mbar_wait(&p_full[g], (uint32_t)((u >> 1) & 1));
        tc_fence_after();
        if (lane == 0) TC_TRACE(2, u);
        const bool first = (ci == 0 && ch == 0);
        if (first && cj > 0) {
          mbar_wait(dkv_empty, (uint32_t)((cj - 1) & 1));
          tc_fence_after();
        }
        const uint32_t qo = (uint32_t)(ci * 2 + ch) * HALF16;
        const int n = u >> 1;
        const uint32_t dso = (uint32_t)(n & 1) * (DS_TILE_BYTES >> 4);
        const uint32_t kvo = (uint32_t)(cj & 1) * BLK16;
        const bool more_dp = (u + 2 < T);
        if (more_dp && di == 0 && dh == 0) {   // sub-step u + 2 opens a key block: its V_j must have landed
          mbar_wait(&kv_full[dj & 1], (uint32_t)((dj >> 1) & 1));
          tc_fence_after();
        }
        if (more_dp && dj == 0 && dh == 0 && di > 0) {   // first touch of dO block di (Q_di is used two sub-steps later)
          mbar_wait(&q_full[di], 0);
          tc_fence_after();
        }
        if (elect_one_sync()) {
          // dV_j += P^T dO_i ,  dK_j += dS^T Q_i : A = bf16 pairs in TMEM (8 columns per 16-query k-step), written by
          // the softmax warps over the dP^T buffer: per 16-column chunk  [P^T (8) | dS^T (8)]
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ts(tmem + C::COL_DV, tmem + C::COL_DP + g * 64 + k * 16, mDO + (qo + k * K16ROWS), ID_TS,
                         (first && k == 0) ? 0u : 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ts(tmem + C::COL_DK, tmem + C::COL_DP + g * 64 + k * 16 + 8, mQ + (qo + k * K16ROWS), ID_TS,
                         (first && k == 0) ? 0u : 1u);
          if (more_dp) issue_dp(g, dj, di, dh);
          if (ch == 1 && ci == NB - 1) {   // last sub-step of key block cj: dK_j / dV_j complete, V_j no longer read
            umma_commit(dkv_full);
            umma_commit(&kv_empty[cj & 1]);
          }
        }
        __syncwarp();
        if (lane == 0) TC_TRACE(12, u);
        advance(cj, ci, ch);
        advance(dj, di, dh);
      }
    }
  } else {
    // ============================ softmax warps ============================
    // Two groups of 8 warps work ping-pong: group g owns TMEM buffer g and therefore every sub-step with
    // (t & 1) == g, i.e. the query half g of each 128-query block.  While one group waits for its tcgen05.ld /
    // fence / barrier round trips the other keeps the MUFU pipe busy.
    const int grp = (warp - TCB_FIRST_SOFTMAX_WARP) >> 3;
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int half = ((warp - TCB_FIRST_SOFTMAX_WARP) & 7) >> 2;     // which 32 of the sub-step's 64 query columns
    const int row = quarter * 32 + lane;        // key row inside the block == TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    bf16* dkb = a.dqkv + row_base * a.ld_qkv + a.D + h * HD;
    bf16* dvb = a.dqkv + row_base * a.ld_qkv + 2 * a.D + h * HD;
    bf16* dqb = a.dqkv + row_base * a.ld_qkv + h * HD;
    constexpr int EC = HD / 2;  // epilogue columns per warp

    auto store_row = [&](bf16* dst, const uint32_t* r, float mul) {
#pragma unroll
      for (int c = 0; c < EC; c += 8) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(r[c]) * mul, __uint_as_float(r[c + 1]) * mul);
        o.y = pack_bf16x2(__uint_as_float(r[c + 2]) * mul, __uint_as_float(r[c + 3]) * mul);
        o.z = pack_bf16x2(__uint_as_float(r[c + 4]) * mul, __uint_as_float(r[c + 5]) * mul);
        o.w = pack_bf16x2(__uint_as_float(r[c + 6]) * mul, __uint_as_float(r[c + 7]) * mul);
        *reinterpret_cast<uint4*>(dst + c) = o;
      }
    };
    auto add_colsum = [&](const uint32_t (&r)[EC], float mul, int col0) {
      float v[EC];
#pragma unroll
      for (int c = 0; c < EC; ++c) v[c] = __uint_as_float(r[c]);
      warp_colsum<EC>(v, lane);
      if (EC == 32) atomicAdd(&s_db[col0 + lane], v[0] * mul);
      else if ((lane & 1) == 0) atomicAdd(&s_db[col0 + ((lane >> 1) & 15)], v[0] * mul);
    };
    // dK_j / dV_j leave through group 0; group 1 only tracks the barrier phase (a waiter may not lag a phase)
    auto epilogue_dkv = [&](int j) {
      mbar_wait(dkv_full, (uint32_t)(j & 1));
      tc_fence_after();
      if (grp != 0) return;
      uint32_t rk[EC], rv[EC];
      tmem_ld_n<EC>(tlane + C::COL_DK + half * EC, rk);
      tmem_ld_n<EC>(tlane + C::COL_DV + half * EC, rv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dkv_empty);
      const int kr = j * 128 + row;
      if (kr < S) {
        store_row(dkb + (long long)kr * a.ld_qkv + half * EC, rk, a.scale);
        store_row(dvb + (long long)kr * a.ld_qkv + half * EC, rv, 1.0f);
      }
      if (a.dbias != nullptr) {   // rows at or beyond S hold exact zeros (P^T / dS^T are masked there)
        add_colsum(rk, a.scale, HD + half * EC);
        add_colsum(rv, 1.0f, 2 * HD + half * EC);
      }
    };

    const uint32_t s_lse_u = smem_u32(s_lse), s_delta_u = smem_u32(s_delta);
    const uint32_t tS = tlane + C::COL_S + grp * 64 + half * 32, tDP = tlane + C::COL_DP + grp * 64 + half * 32;
    auto substep = [&](auto mask_tag, int n, int j, int i) {
      constexpr bool MASK = decltype(mask_tag)::value;   // last key block: rows at or beyond S contribute nothing
      const bool tr = (lane == 0 && ((warp - TCB_FIRST_SOFTMAX_WARP) & 7) == 0);
      const bool kv_ok = !MASK || (j * 128 + row) < S;
      const uint32_t qoff = (uint32_t)((i * 128 + grp * 64 + half * 32) * 4);
      // ---- phase 1: P = exp2(S^T * scale - lse)   (needs only S^T; frees the S^T buffer immediately)
      if (tr) TC_TRACE(3 + grp * 4, n);
      mbar_wait(&s_full[grp], (uint32_t)(n & 1));
      tc_fence_after();
      if (tr) TC_TRACE(4 + grp * 4, n);
      uint32_t pw[16];
      {
        uint32_t sr[32];
        tmem_ld_32x32b_x32(tS, sr);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_read[grp]);
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 ls;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                       : "=f"(ls.x), "=f"(ls.y), "=f"(ls.z), "=f"(ls.w) : "r"(s_lse_u + qoff + c * 4));
          // (moving a quarter of these exponentials to the FMA-pipe polynomial of the forward kernel was measured:
          //  1.585 -> 1.655 ms — the backward's softmax warps are bound by instruction issue, not by the MUFU pipe)
          float p0 = exp2f(fmaf(__uint_as_float(sr[c]), a.scale_log2, -ls.x));
          float p1 = exp2f(fmaf(__uint_as_float(sr[c + 1]), a.scale_log2, -ls.y));
          float p2 = exp2f(fmaf(__uint_as_float(sr[c + 2]), a.scale_log2, -ls.z));
          float p3 = exp2f(fmaf(__uint_as_float(sr[c + 3]), a.scale_log2, -ls.w));
          if (MASK && !kv_ok) p0 = p1 = p2 = p3 = 0.f;
          pw[c / 2] = pack_bf16x2(p0, p1);
          pw[c / 2 + 1] = pack_bf16x2(p2, p3);
        }
      }
      // ---- phase 2: dS = P o (dP^T - delta); P^T and dS^T replace dP^T in TMEM, dS^T also goes to shared memory
      if (n >= 2) mbar_wait(&ds_empty[n & 1], (uint32_t)(((n >> 1) - 1) & 1));
      mbar_wait(&dp_full[grp], (uint32_t)(n & 1));
      tc_fence_after();
      if (tr) TC_TRACE(5 + grp * 4, n);
      const uint32_t ds = sDS + (n & 1) * DS_TILE_BYTES + grp * 16384 + row * 128;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {      // two chunks of 16 query columns
        uint32_t dr[16];
        tmem_ld_32x32b_x16(tDP + cc * 16, dr);
        tmem_ld_wait();
        uint32_t st[16];                   // [P^T (8 words) | dS^T (8 words)] of this chunk
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          float4 dl;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                       : "=f"(dl.x), "=f"(dl.y), "=f"(dl.z), "=f"(dl.w) : "r"(s_delta_u + qoff + (cc * 16 + c) * 4));
          const uint32_t w0 = pw[cc * 8 + c / 2], w1 = pw[cc * 8 + c / 2 + 1];
          // dS = P o (dP - delta) in packed bf16 (HMUL2.BF16): P is already the bf16-rounded probability the dV product
          // uses; the difference is rounded to bf16 once more, the product once — the softmax warps are issue-bound and
          // this is 8 instead of 14 instructions per 4 elements
          const uint32_t e0 = pack_bf16x2(__uint_as_float(dr[c]) - dl.x, __uint_as_float(dr[c + 1]) - dl.y);
          const uint32_t e1 = pack_bf16x2(__uint_as_float(dr[c + 2]) - dl.z, __uint_as_float(dr[c + 3]) - dl.w);
          st[c / 2] = w0;
          st[c / 2 + 1] = w1;
          st[8 + c / 2] = hmul2_bf16(w0, e0);
          st[8 + c / 2 + 1] = hmul2_bf16(w1, e1);
        }
        tmem_st_32x32b_x16(tDP + cc * 16, st);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          st_shared_v4(ds + (((half * 4 + cc * 2 + c) ^ (row & 7)) << 4), st[8 + 4 * c], st[8 + 4 * c + 1],
                       st[8 + 4 * c + 2], st[8 + 4 * c + 3]);
      }
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&p_full[grp]);
        mbar_arrive(&pq_full[n & 1]);
      }
      if (tr) TC_TRACE(6 + grp * 4, n);
    };
    {
      int n = 0;
      for (int j = 0; j < NB; ++j) {
#pragma unroll 1
        for (int i = 0; i < NB; ++i, ++n) {
          if (j == NB - 1) substep(std::true_type{}, n, j, i);
          else substep(std::false_type{}, n, j, i);
          // dK / dV of the previous key block leave one sub-step late, so the tensor pipe's drain is hidden
          if (i == 0 && j > 0) epilogue_dkv(j - 1);
        }
      }
    }
    epilogue_dkv(NB - 1);
    mbar_wait(dq_done, 0);   // every dQ product has retired
    tc_fence_after();
#pragma unroll 1
    for (int i = grp; i < NB; i += 2) {
      uint32_t rq[EC];
      tmem_ld_n<EC>(tlane + C::COL_DQ + i * HD + half * EC, rq);
      tmem_ld_wait();
      const int qr = i * 128 + row;
      if (qr < S) store_row(dqb + (long long)qr * a.ld_qkv + half * EC, rq, a.scale);
      if (a.dbias != nullptr) add_colsum(rq, a.scale, half * EC);
    }
  }

  if (threadIdx.x == 96) TC_TRACE(17, 2);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_TRACE(17, 3);
  if (a.dbias != nullptr && threadIdx.x < 3 * HD)   // one atomic per column per CTA
    atomicAdd(a.dbias + (threadIdx.x / HD) * a.D + h * HD + (threadIdx.x % HD), s_db[threadIdx.x]);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int HD>
int launch_bwd(const TcArgs& a, int n_seq, cudaStream_t stream) {
  using C = TcCfg<HD>;
  const int S_pad = a.NB * 128;
  const int smem = 1024 + 2 * a.NB * C::BLK_BYTES + 4 * C::BLK_BYTES + 2 * DS_TILE_BYTES + 2 * S_pad * 4 + 1024;
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    avs_set_error("avs_attention_bwd(tc): cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
    return (int)e;
  }
  dim3 grid(a.H, n_seq);
  attn_bwd_tc_kernel<HD><<<grid, TCB_THREADS, smem, stream>>>(a);
  return avs_check_launch("attn_bwd_tc_kernel");
}

// =====================================================================================================
// tcgen05 / TMEM attention FORWARD (MAE decoder shape: head_dim 32, 256 <= S <= 768).
//
// The forward of F.scaled_dot_product_attention (cav_mae_base.py:58-77) at head_dim 32 is bound by the exponentials
// (MUFU.EX2 issues 16 lanes per clock per SM — tools/mufu_probe.cu — against S^2 exponentials per head), not by the
// tensor pipe.  Everything else is kept off the softmax warps:
//   * thread = query row (TMEM lane): row max and row sum are per-thread serial reductions — no shuffles.
//   * online softmax with a LAZY running maximum: the maximum a row exponentiates against is only raised when the
//     new block exceeds it by more than 2^8 (probabilities then stay below 256, exact in the final O / l); raising it
//     rescales O in tensor memory (tcgen05.ld / st by the owning thread).  That happens in the first one or two units
//     of a query block and almost never afterwards, so there is no correction warp and no second pass.
//   * when the Cauchy-Schwarz bound |q_i| max_j |k_j| of a query block's scores is small (<= 40 in the log2 domain) it
//     replaces the maximum altogether: exp2(score - bound) cannot underflow, so the maximum pass, the exchange between
//     the two warps of a row and the rescaling all disappear (the common case; the lazy maximum is the general one).
//   * P goes back to TENSOR MEMORY as packed bf16 and is the A operand of the P V product (tcgen05.mma with A in TMEM),
//     so the probabilities never touch shared memory.
//   * row sums on the tensor pipe too: L += P x ones (a 16-column product per k-step), so the softmax warps spend no
//     FADD per element and the sum is taken over exactly the bf16 probabilities that multiply V.
//   * one CTA = one (sequence, head), 10 warps: loader, MMA issuer, 8 softmax warps (the two warps of a TMEM lane
//     quarter split the unit's 128 key columns and agree on the row maximum through shared memory + a 64-thread named
//     barrier).  K and V of the whole head resident in swizzled shared memory (2 x 45 KB), Q blocks streamed (2 x 8 KB),
//     256 TMEM columns: TWO CTAs PER SM (the occupancy API reports 1 for any kernel with tcgen05.alloc; the hardware
//     co-schedules two 256-column CTAs — tools/tmem_occ_probe.cu), i.e. 4 softmax warps per SM sub-partition: one warp
//     alone reaches only 55-70 % of the EX2 rate with this instruction mix, two or more 86-97 % (tools/mufu_probe.cu).
// Measured (S = 708, 4096 heads): 0.775 ms against 0.95 ms for the mma.sync kernel (0.82 ms with every exponential on MUFU).  Timing experiments on the 64-key-unit
// predecessor: without the exponentials it still took 0.72 ms — ~7.5 warp-instructions per score element (barrier
// handling and loop control amortised over 32 elements per thread and unit) made it issue-bound before it is MUFU-bound
// (MUFU floor 0.54 ms); hence 128-key units, read from TMEM twice (maximum, then exponentials) in 32-column chunks.
// Work unit = 128 keys: S (128 fp32 columns) -> P (64 packed columns) -> O, L; S and P single-buffered (the softmax work
// on a unit covers the next S product and the retirement of the previous P V).  The last unit of a sequence only takes
// as many 16-key steps as it has valid keys (S = 708: N = 80).
// =====================================================================================================
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}

constexpr int TCF_SOFTMAX_WARPS = 8;
constexpr int TCF_FIRST_SOFTMAX_WARP = 2;
constexpr int TCF_THREADS = 32 * (TCF_FIRST_SOFTMAX_WARP + TCF_SOFTMAX_WARPS);
constexpr int TCF_COL_S = 0, TCF_COL_P = 128, TCF_COL_O = 192, TCF_COL_L = 224, TCF_TMEM_COLS = 256;   // S 128 | P 64 | O 32 | L 16
#ifndef TCF_POLY_EVERY
#define TCF_POLY_EVERY 4   // every 4th pair of exponentials on the FMA pipe (0 = all on MUFU)
#endif
// largest score bound (log2 domain) for the no-maximum fast path: probabilities stay >= 2^-80, so that even their
// products with small V entries and the fp32 accumulators of O stay far from the denormal range
constexpr float TCF_BOUND_MAX = 40.0f;

struct TcFwdArgs {
  const bf16* qkv;
  bf16* out;
  float* lse2;
  long long ld_qkv, ld_o;
  int S, NB, NU, H, D;   // NB = 128-query blocks, NU = 64-key units
  float scale_log2;
};

template <int HD>
__global__ void __launch_bounds__(TCF_THREADS, 2) attn_fwd_tc_kernel(const TcFwdArgs a) {
  using C = TcCfg<HD>;
  static_assert(TCF_COL_O + HD <= TCF_COL_L, "O does not fit the TMEM allocation");
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = a.S, NB = a.NB, NU = a.NU;                      // NU = 128-key units
  const int n_last = (((S - (NU - 1) * 128) + 15) >> 4) << 4;  // MMA N of the last unit (16 .. 128)
  const int kv_rows = (NU - 1) * 128 + n_last;                  // rows the MMAs touch (multiple of 16: tiles stay 1 KB aligned)
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + kv_rows * C::ROWB;
  const uint32_t sQ = sV + kv_rows * C::ROWB;          // [2] 128-row blocks
  // 2 KB of bf16 1.0: the B operand of the row-sum product L += P 1 (any descriptor that stays inside reads ones)
  const uint32_t sOnes = sQ + 2 * C::BLK_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kv_rows * C::ROWB + 2 * C::BLK_BYTES + 2048);
  uint64_t* q_full = bars;          // [2] loader -> issuer
  uint64_t* q_empty = bars + 2;     // [2] issuer (commit) -> loader
  uint64_t* s_full = bars + 4;      // issuer (commit) -> softmax: S unit in TMEM
  uint64_t* s_read = bars + 5;      // softmax -> issuer: every warp has made its last read of the S unit
  uint64_t* p_full = bars + 6;      // softmax -> issuer: P unit written to TMEM
  uint64_t* p_empty = bars + 7;     // issuer (commit) -> softmax: P V of that unit has retired
  uint64_t* o_full = bars + 8;      // issuer (commit) -> softmax: O of a query block complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* s_mx = reinterpret_cast<float*>(bars + 16);   // [2][2][128] partial row maxima (unit parity, column half, row)
  uint32_t* s_k2max = reinterpret_cast<uint32_t*>(bars + 10);   // bits of max_j |k_j|^2 (non-negative floats order as integers)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, seq = blockIdx.y;
  const long long row_base = (long long)seq * S;
  const bf16* qb = a.qkv + row_base * a.ld_qkv + h * HD;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_read, TCF_SOFTMAX_WARPS);
    mbar_init(p_full, TCF_SOFTMAX_WARPS);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TCF_TMEM_COLS);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) *s_k2max = 0u;
  __syncthreads();   // (orders the initialisation before every warp's atomicMax below)
  // K, V of the whole head and the first Q block: one cooperative load, one exposed HBM round trip
  load_rows_async<HD>(sK, qb + a.D, a.ld_qkv, 0, kv_rows, S, threadIdx.x, TCF_THREADS);
  load_rows_async<HD>(sV, qb + 2 * a.D, a.ld_qkv, 0, kv_rows, S, threadIdx.x, TCF_THREADS);
  load_rows_async<HD>(sQ, qb, a.ld_qkv, 0, 128, S, threadIdx.x, TCF_THREADS);
  for (int i = threadIdx.x; i < 2048 / 16; i += TCF_THREADS)
    st_shared_v4(sOnes + i * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  // Cauchy-Schwarz bound of every score of the head: |q_i . k_j| <= |q_i| max_j |k_j|.  When the bound is small enough
  // that exp2(score - bound) cannot underflow (see the softmax warps), it replaces the running row maximum and the
  // whole maximum pass.  The K rows read here are the ones the cp.async above has just requested (L2 hits).
  {
    float k2 = 0.f;
    for (int r = threadIdx.x; r < S; r += TCF_THREADS) {
      const uint4* kr = reinterpret_cast<const uint4*>(qb + a.D + (long long)r * a.ld_qkv);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 v = __ldg(kr + c);
        float2 f;
        f = unpack_bf16x2(v.x); acc = fmaf(f.x, f.x, fmaf(f.y, f.y, acc));
        f = unpack_bf16x2(v.y); acc = fmaf(f.x, f.x, fmaf(f.y, f.y, acc));
        f = unpack_bf16x2(v.z); acc = fmaf(f.x, f.x, fmaf(f.y, f.y, acc));
        f = unpack_bf16x2(v.w); acc = fmaf(f.x, f.x, fmaf(f.y, f.y, acc));
      }
      k2 = fmaxf(k2, acc);
    }
    k2 = warp_max(k2);
    if (lane == 0) atomicMax(s_k2max, __float_as_uint(k2));
  }
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int U = NB * NU;

  if (warp == 0) {
    // ============================ loader: Q blocks 1 .. NB-1 ============================
    if (lane == 0) mbar_arrive(&q_full[0]);
    for (int i = 1; i < NB; ++i) {
      if (i >= 2) mbar_wait(&q_empty[i & 1], (uint32_t)(((i >> 1) - 1) & 1));
      load_rows_async<HD>(sQ + (i & 1) * C::BLK_BYTES, qb, a.ld_qkv, i * 128, 128, S, lane, 32);
      cp_async_wait_all();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_full[i & 1]);
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    // S and P are single-buffered: with 128-key units the softmax work on unit u (>= 600 cycles) covers both the
    // S(u+1) product, issued as soon as the warps have made their last read of S(u), and the retirement of P(u-1) V.
    constexpr uint32_t ID_PV = idesc_bf16(128, HD, 0, 1);     // A = P in TMEM, B = V rows MN-major
    constexpr uint32_t K16ROWS = (16 * C::ROWB) >> 4, UNIT16 = (128 * C::ROWB) >> 4, BLK16 = C::BLK_BYTES >> 4;
    const uint64_t kQ = make_desc(sQ, 0, C::SBO, C::LT), kK = make_desc(sK, 0, C::SBO, C::LT);
    const uint64_t mV = make_desc(sV, C::SBO, C::SBO, C::LT);
    const uint32_t id_s_full = idesc_bf16(128, 128, 0, 0), id_s_last = idesc_bf16(128, n_last, 0, 0);
    // row sums on the tensor pipe: L[128 x 16] += P[128 x 16 keys] * ones[16 keys x 16] — every column of L is the
    // row sum of the bf16 probabilities that also multiply V (one FADD per element less on the softmax warps)
    constexpr uint32_t ID_L = idesc_bf16(128, 16, 0, 0);
    const uint64_t kOnes = make_desc(sOnes, 128, 128, 0u);
    int si = 0, sj = 0;          // S stream position: query block, key unit
    auto issue_s = [&]() {
      if (sj == 0) {
        mbar_wait(&q_full[si & 1], (uint32_t)((si >> 1) & 1));
        tc_fence_after();
      }
      const uint32_t qo = (uint32_t)(si & 1) * BLK16, ko = (uint32_t)sj * UNIT16;
      if (elect_one_sync()) {
        const uint32_t id = (sj == NU - 1) ? id_s_last : id_s_full;
#pragma unroll
        for (int k = 0; k < C::KSTEPS; ++k)
          umma_bf16_ss(tmem + TCF_COL_S, kQ + (qo + 2 * k), kK + (ko + 2 * k), id, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        if (sj == NU - 1) umma_commit(&q_empty[si & 1]);    // last read of this Q block
      }
      __syncwarp();
      if (++sj == NU) { sj = 0; ++si; }
    };
    issue_s();
    int pj = 0;                  // key unit of the P V stream
    for (int u = 0; u < U; ++u) {
      const uint32_t par = (uint32_t)(u & 1);
      if (u + 1 < U) {
        mbar_wait(s_read, par);
        tc_fence_after();
        issue_s();
      }
      mbar_wait(p_full, par);
      tc_fence_after();
      const int ksteps = (pj == NU - 1) ? (n_last >> 4) : 8;
      const uint32_t vo = (uint32_t)pj * UNIT16;
      if (elect_one_sync()) {
        for (int k = 0; k < ksteps; ++k) {
          umma_bf16_ts(tmem + TCF_COL_O, tmem + TCF_COL_P + k * 8, mV + (vo + k * K16ROWS), ID_PV,
                       (pj > 0 || k > 0) ? 1u : 0u);
          umma_bf16_ts(tmem + TCF_COL_L, tmem + TCF_COL_P + k * 8, kOnes, ID_L, (pj > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(p_empty);
        if (pj == NU - 1) umma_commit(o_full);
      }
      __syncwarp();
      if (++pj == NU) pj = 0;
    }
  } else {
    // ============================ softmax warps ============================
    // 8 warps: the two warps of a TMEM lane quarter split the unit's 128 key columns (64 each), so every SM
    // sub-partition runs 4 softmax warps (2 per CTA x 2 CTAs).  One warp alone cannot keep the MUFU pipe busy with this
    // instruction mix (tools/mufu_probe.cu: 55-70 % of the EX2 rate; two or more reach 86-97 %).
    // Each unit is read from TMEM twice in 32-column chunks — once for the row maximum, once for the exponentials —
    // which keeps the scores out of long-lived registers (tcgen05.ld is one instruction per 32 elements).
    const int sw = warp - TCF_FIRST_SOFTMAX_WARP;
    const int quarter = warp & 3;                   // the TMEM lane quarter this warp may access
    const int half = sw >> 2;                       // which 64 of the unit's 128 key columns
    const int row = quarter * 32 + lane;            // query row inside the block == TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tS = tlane + TCF_COL_S + half * 64, tP = tlane + TCF_COL_P + half * 32;
    bf16* ob = a.out + row_base * a.ld_o + h * HD;
    float* lp = a.lse2 + ((long long)seq * a.H + h) * S;
    const int valid_last = S - (NU - 1) * 128 - half * 64;   // valid keys among this warp's 64 columns of the last unit
    float m_run = -INFINITY;
    const float k2max = __uint_as_float(*s_k2max);
    // Fast path of a query block: every row of this lane quarter has a score bound B = |q| max|k| * scale (log2 domain)
    // of at most TCF_BOUND_MAX.  All scores then lie in [-B, B], so exp2(score - B) >= 2^(-2 B) stays a normal fp32 /
    // bf16 number (with margin for the products with V) and B can stand in for the row maximum: no maximum pass, no exchange, no rescaling.  The decision
    // depends only on the rows of the lane quarter, so the two warps that share them always agree.
    auto block_bound = [&](int blk) {
      const int qr = blk * 128 + row;
      float q2 = 0.f;
      if (qr < S) {
        const uint4* qrow = reinterpret_cast<const uint4*>(qb + (long long)qr * a.ld_qkv);
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          const uint4 v = __ldg(qrow + c);
          float2 f;
          f = unpack_bf16x2(v.x); q2 = fmaf(f.x, f.x, fmaf(f.y, f.y, q2));
          f = unpack_bf16x2(v.y); q2 = fmaf(f.x, f.x, fmaf(f.y, f.y, q2));
          f = unpack_bf16x2(v.z); q2 = fmaf(f.x, f.x, fmaf(f.y, f.y, q2));
          f = unpack_bf16x2(v.w); q2 = fmaf(f.x, f.x, fmaf(f.y, f.y, q2));
        }
      }
      return sqrtf(q2 * k2max) * a.scale_log2 * 1.0001f + 1e-3f;     // margin for the fp32 rounding of the products
    };
    float bound = block_bound(0);
    bool fast = __all_sync(0xffffffffu, bound <= TCF_BOUND_MAX);
    if (fast) m_run = bound;
    int jj = 0, i = 0;
    for (int u = 0; u < U; ++u) {
      const uint32_t par = (uint32_t)(u & 1);
      const bool last = (jj == NU - 1);
      mbar_wait(s_full, par);
      tc_fence_after();
      if (!fast) {
      // ---- pass A: row maximum of this warp's 64 columns
      float mx = -INFINITY;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t sr[32];
        tmem_ld_32x32b_x32(tS + ch * 32, sr);
        tmem_ld_wait();
        if (!last) {
          float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            a0 = fmaxf(a0, fmaxf(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])));
            a1 = fmaxf(a1, fmaxf(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])));
            a2 = fmaxf(a2, fmaxf(__uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5])));
            a3 = fmaxf(a3, fmaxf(__uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7])));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (ch * 32 + c < valid_last) mx = fmaxf(mx, __uint_as_float(sr[c]));
        }
      }
      // ---- running maximum (lazy), agreed between the two warps that share this row
      s_mx[par * 256 + half * 128 + row] = mx;
      named_bar_sync(1 + quarter, 64);                 // the two warps of this lane quarter
      mx = fmaxf(mx, s_mx[par * 256 + (half ^ 1) * 128 + row]);
      const float m_new = mx * a.scale_log2;            // scale > 0 commutes with max
      const bool raise = m_new > m_run + TCF_LAZY;      // always true in the first unit of a query block (m_run = -inf)
      // P(u-1) V (and with it every earlier product into O / L) must have retired before P is overwritten or O rescaled
      if (u >= 1) {
        mbar_wait(p_empty, (uint32_t)((u - 1) & 1));
        tc_fence_after();
      }
      if (__any_sync(0xffffffffu, raise)) {
        const float m_upd = raise ? m_new : m_run;
        if (jj > 0 && half == 0) {
          // O and L were accumulated against the old maximum: rescale (factor 1 for the rows that keep theirs)
          const float f = exp2f(m_run - m_upd);
          uint32_t ro[HD], rl[1];
          tmem_ld_n<HD>(tlane + TCF_COL_O, ro);
          tmem_ld_32x32b_x1(tlane + TCF_COL_L, rl);      // only column 0 of L is ever read back
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < HD; ++c) ro[c] = __float_as_uint(__uint_as_float(ro[c]) * f);
          rl[0] = __float_as_uint(__uint_as_float(rl[0]) * f);
          tmem_st_32x32b_x32(tlane + TCF_COL_O, ro);
          tmem_st_32x32b_x1(tlane + TCF_COL_L, rl);
        }
        m_run = m_upd;
      }
      }   // !fast
      // ---- pass B: P = exp2(S * scale - max) -> bf16 pairs -> TMEM.  P is single-buffered: on the fast path the wait
      // for P(u-1) V to retire comes only after this unit's exponentials, which hide the product's latency.
      uint32_t pw[2][16];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t sr[32];
        tmem_ld_32x32b_x32(tS + ch * 32, sr);
        tmem_ld_wait();
        if (ch == 1) {                                  // last read of S(u): the issuer may overwrite the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_read);
        }
        const bool whole = !last || ch * 32 + 32 <= valid_last;   // every column of this chunk is a valid key
        if (whole && fast && TCF_POLY_EVERY > 0) {
          // fast path: exponents lie in [-80, 0], so every TCF_POLY_EVERY-th pair can take the polynomial
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float x0 = fmaf(__uint_as_float(sr[c]), a.scale_log2, -m_run);
            const float x1 = fmaf(__uint_as_float(sr[c + 1]), a.scale_log2, -m_run);
            if (((c >> 1) % TCF_POLY_EVERY) == TCF_POLY_EVERY - 1) pw[ch][c >> 1] = pack_bf16x2(exp2_poly(x0), exp2_poly(x1));
            else pw[ch][c >> 1] = pack_bf16x2(exp2f(x0), exp2f(x1));
          }
        } else if (whole) {
#pragma unroll
          for (int c = 0; c < 32; c += 2)
            pw[ch][c >> 1] = pack_bf16x2(exp2f(fmaf(__uint_as_float(sr[c]), a.scale_log2, -m_run)),
                                         exp2f(fmaf(__uint_as_float(sr[c + 1]), a.scale_log2, -m_run)));
        } else {
          // last unit of the sequence: 16-column groups without a valid key cost no exponentials (S = 708: the second
          // warp of a row has 4 valid keys among its 64 columns)
#pragma unroll
          for (int g16 = 0; g16 < 2; ++g16) {
            if (ch * 32 + g16 * 16 < valid_last) {
#pragma unroll
              for (int c = g16 * 16; c < g16 * 16 + 16; c += 2) {
                float p0 = exp2f(fmaf(__uint_as_float(sr[c]), a.scale_log2, -m_run));
                float p1 = exp2f(fmaf(__uint_as_float(sr[c + 1]), a.scale_log2, -m_run));
                if (ch * 32 + c >= valid_last) p0 = 0.f;
                if (ch * 32 + c + 1 >= valid_last) p1 = 0.f;
                pw[ch][c >> 1] = pack_bf16x2(p0, p1);
              }
            } else {
#pragma unroll
              for (int c = g16 * 8; c < g16 * 8 + 8; ++c) pw[ch][c] = 0u;
            }
          }
        }
      }
      if (fast && u >= 1) {   // P(u-1) V must have retired before P is overwritten
        mbar_wait(p_empty, (uint32_t)((u - 1) & 1));
        tc_fence_after();
      }
      tmem_st_32x32b_x16(tP, pw[0]);
      tmem_st_32x32b_x16(tP + 16, pw[1]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (last) {
        // ---- epilogue of query block i: O / l, log-sum-exp (each warp stores half of the head_dim columns)
        mbar_wait(o_full, (uint32_t)(i & 1));
        tc_fence_after();
        constexpr int EC = HD / 2;
        uint32_t ro[EC], rl[1];
        tmem_ld_n<EC>(tlane + TCF_COL_O + half * EC, ro);
        tmem_ld_32x32b_x1(tlane + TCF_COL_L, rl);
        tmem_ld_wait();
        const float l = __uint_as_float(rl[0]);
        const float inv = 1.f / l;
        const int qr = i * 128 + row;
        if (qr < S) {
          bf16* dst = ob + (long long)qr * a.ld_o + half * EC;
#pragma unroll
          for (int c = 0; c < EC; c += 8) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(ro[c]) * inv, __uint_as_float(ro[c + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(ro[c + 2]) * inv, __uint_as_float(ro[c + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(ro[c + 4]) * inv, __uint_as_float(ro[c + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(ro[c + 6]) * inv, __uint_as_float(ro[c + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c) = o;
          }
          if (half == 0) lp[qr] = m_run + log2f(l);
        }
        // the first P V of the next query block (accumulate = 0 into O / L) is issued only after all 8 warps have
        // produced its P, i.e. after the tcgen05.ld above has completed in each of them: no extra barrier needed
        jj = 0; ++i;
        m_run = -INFINITY;
        if (i < NB) {
          bound = block_bound(i);
          fast = __all_sync(0xffffffffu, bound <= TCF_BOUND_MAX);
          if (fast) m_run = bound;
        }
      } else {
        ++jj;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, TCF_TMEM_COLS);
  }
}

// =====================================================================================================
// Persistent ping-pong variant of the forward: ONE CTA per SM that walks over (sequence, head) pairs, with TWO
// softmax groups that take alternate query blocks of the head and hand a "MUFU token" back and forth (named barriers),
// so that one group's exponentials always run under the other group's tcgen05.ld / st, barrier and max work.
// Two co-resident CTAs cannot do that: measured (ncu), their softmax warps fall into a convoy — both in the
// exponential phase (sharing the 16-lane MUFU) and then both outside it — and the MUFU pipe idles 40 % of the time.
//   warp 0      K/V loader: TMA (SWIZZLE_64B boxes) of the NEXT head's K and V into the other shared-memory stage
//   warp 1      Q loader of both groups (TMA, double-buffered 128-row blocks per group, in consumption order)
//   warps 2, 3  MMA issuers of group A / B
//   warps 4-7   softmax group A (TMEM columns 0-255), warps 8-11 softmax group B (columns 256-511)
// (12 warps: the register file then allows 168 registers per thread)
// Shared memory: 2 stages x (K | V) x kv_rows x 64 B + 2 groups x 2 x 8 KB of Q  (S = 708: 212 KB).
// =====================================================================================================
constexpr int TCP_FIRST_SOFTMAX_WARP = 4;
constexpr int TCP_THREADS = 32 * (TCP_FIRST_SOFTMAX_WARP + 8);

struct TcPpArgs {
  CUtensorMap map128, map16;   // qkv [rows, 3D] bf16, boxes of 128 / 16 rows x HD columns, SWIZZLE_64B
  bf16* out;
  float* lse2;
  long long ld_o;
  int S, NB, NU, H, D, n_heads;   // n_heads = n_seq * H
  int tokens;                     // 1: the two softmax groups alternate explicitly (MUFU token)
  float scale_log2;
};


template <int HD>
__global__ void __launch_bounds__(TCP_THREADS, 1) attn_fwd_pp_kernel(const __grid_constant__ TcPpArgs a) {
  using C = TcCfg<HD>;
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = a.S, NB = a.NB, NU = a.NU;
  const int n_last = (((S - (NU - 1) * 64) + 15) >> 4) << 4;   // MMA N of the last unit (16 .. 64)
  const int kv_rows = (NU - 1) * 64 + n_last;
  const uint32_t kv_bytes = (uint32_t)kv_rows * C::ROWB;
  const uint32_t sKV = smem_u32(smem);                          // stage s: K at sKV + s*2*kv_bytes, V right after
  const uint32_t sQ = sKV + 4 * kv_bytes;                       // group g, buffer b: sQ + (g*2 + b) * BLK_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * kv_bytes + 4 * C::BLK_BYTES);
  uint64_t* kv_full = bars;           // [2]
  uint64_t* kv_empty = bars + 2;      // [2] both issuers commit
  uint64_t* gb = bars + 4;            // per group 13 barriers
  auto q_full = [&](int g, int b) { return gb + g * 13 + b; };
  auto q_empty = [&](int g, int b) { return gb + g * 13 + 2 + b; };
  auto s_full = [&](int g, int b) { return gb + g * 13 + 4 + b; };
  auto s_read = [&](int g, int b) { return gb + g * 13 + 6 + b; };
  auto p_full = [&](int g, int b) { return gb + g * 13 + 8 + b; };
  auto p_empty = [&](int g, int b) { return gb + g * 13 + 10 + b; };
  auto o_full = [&](int g) { return gb + g * 13 + 12; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
    }
    for (int g = 0; g < 2; ++g) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(q_full(g, b), 1);
        mbar_init(q_empty(g, b), 1);
        mbar_init(s_full(g, b), 1);
        mbar_init(s_read(g, b), 4);
        mbar_init(p_full(g, b), 4);
        mbar_init(p_empty(g, b), 1);
      }
      mbar_init(o_full(g), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // this CTA's heads: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int n_my = (a.n_heads - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // query blocks of a head: group g takes blocks g, g+2, ...
  const int nbg0 = (NB + 1) >> 1, nbg1 = NB >> 1;

  if (warp == 0) {
    // ============================ K/V loader ============================
    if (lane == 0) {
      tma_prefetch_desc(&a.map128);
      tma_prefetch_desc(&a.map16);
      for (int k = 0; k < n_my; ++k) {
        const int hh = blockIdx.x + k * gridDim.x, seq = hh / a.H, h = hh - seq * a.H;
        const int st = k & 1;
        if (k >= 2) mbar_wait_backoff(&kv_empty[st], (uint32_t)(((k >> 1) - 1) & 1), 1000);
        mbar_arrive_expect_tx(&kv_full[st], 2 * kv_bytes);
        const int row0 = seq * S;
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {                           // K then V
          uint8_t* dst = smem + (size_t)(st * 2 + m) * kv_bytes;
          const int col = (m + 1) * a.D + h * HD;
          int r = 0;
          for (; r + 128 <= kv_rows; r += 128) tma_load_2d(dst + r * C::ROWB, &a.map128, &kv_full[st], col, row0 + r);
          for (; r < kv_rows; r += 16) tma_load_2d(dst + r * C::ROWB, &a.map16, &kv_full[st], col, row0 + r);
        }
      }
    }
  } else if (warp == 1) {
    // ============================ Q loader (both groups) ============================
    // Blocks are requested in the order the groups consume them: A(t), B(t), A(t+1), ...  The groups advance in
    // lockstep (token passing), so waiting for A's buffer never starves B of a block it can already use.
    if (lane == 0) {
      int n[2] = {0, 0};   // running count of Q blocks per group
      for (int k = 0; k < n_my; ++k) {
        const int hh = blockIdx.x + k * gridDim.x, seq = hh / a.H, h = hh - seq * a.H;
        for (int i = 0; i < NB; ++i) {
          const int g = i & 1, b = n[g] & 1;
          if (n[g] >= 2) mbar_wait_backoff(q_empty(g, b), (uint32_t)(((n[g] >> 1) - 1) & 1), 500);
          mbar_arrive_expect_tx(q_full(g, b), C::BLK_BYTES);
          tma_load_2d(smem + 4 * kv_bytes + (size_t)(g * 2 + b) * C::BLK_BYTES, &a.map128, q_full(g, b), h * HD,
                      seq * S + i * 128);
          ++n[g];
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ============================ MMA issuers ============================
    const int g = warp - 2;
    const int nbg = g == 0 ? nbg0 : nbg1;
    const int UH = nbg * NU;                                   // units of this group per head
    constexpr uint32_t ID_PV = idesc_bf16(128, HD, 0, 1);
    constexpr uint32_t K16ROWS = (16 * C::ROWB) >> 4, UNIT16 = (64 * C::ROWB) >> 4, BLK16 = C::BLK_BYTES >> 4;
    const uint64_t kQ = make_desc(sQ + g * 2 * C::BLK_BYTES, 0, C::SBO, C::LT);
    const uint64_t kK = make_desc(sKV, 0, C::SBO, C::LT);
    const uint64_t mV = make_desc(sKV + kv_bytes, C::SBO, C::SBO, C::LT);
    const uint32_t stage16 = (2 * kv_bytes) >> 4;
    const uint32_t id_s_full = idesc_bf16(128, 64, 0, 0), id_s_last = idesc_bf16(128, n_last, 0, 0);
    const uint32_t tg = tmem + g * 256;
    const long long U = (long long)n_my * UH;                  // units of this group over the CTA's lifetime
    int sk = 0, sq = 0, sj = 0;   // S stream: head index k, running Q-block count, key unit
    int sqh = 0;                  // Q blocks of the current head already started
    auto issue_s = [&](long long v) {
      const uint32_t b = (uint32_t)(v & 1);
      if (sj == 0) {
        if (sqh == 0) {
          mbar_wait(&kv_full[sk & 1], (uint32_t)((sk >> 1) & 1));
          tc_fence_after();
        }
        mbar_wait(q_full(g, sq & 1), (uint32_t)((sq >> 1) & 1));
        tc_fence_after();
      }
      const uint32_t qo = (uint32_t)(sq & 1) * BLK16, ko = (uint32_t)(sk & 1) * stage16 + (uint32_t)sj * UNIT16;
      if (elect_one_sync()) {
        const uint32_t id = (sj == NU - 1) ? id_s_last : id_s_full;
#pragma unroll
        for (int k = 0; k < C::KSTEPS; ++k)
          umma_bf16_ss(tg + TCF_COL_S + b * 64, kQ + (qo + 2 * k), kK + (ko + 2 * k), id, k > 0 ? 1u : 0u);
        umma_commit(s_full(g, b));
        if (sj == NU - 1) umma_commit(q_empty(g, sq & 1));
      }
      __syncwarp();
      if (++sj == NU) {
        sj = 0; ++sq;
        if (++sqh == nbg) { sqh = 0; ++sk; }
      }
    };
    if (U > 0) issue_s(0);
    if (U > 1) issue_s(1);
    int pk = 0, pqh = 0, pj = 0;
    for (long long u = 0; u < U; ++u) {
      const uint32_t b = (uint32_t)(u & 1);
      if (u + 2 < U) {
        mbar_wait(s_read(g, b), (uint32_t)((u >> 1) & 1));
        tc_fence_after();
        issue_s(u + 2);
      }
      mbar_wait(p_full(g, b), (uint32_t)((u >> 1) & 1));
      tc_fence_after();
      const int ksteps = (pj == NU - 1) ? (n_last >> 4) : 4;
      const uint32_t vo = (uint32_t)(pk & 1) * stage16 + (uint32_t)pj * UNIT16;
      const bool head_done = (pj == NU - 1) && (pqh == nbg - 1);
      if (elect_one_sync()) {
        for (int k = 0; k < ksteps; ++k)
          umma_bf16_ts(tg + TCF_COL_O, tg + TCF_COL_P + b * 32 + k * 8, mV + (vo + k * K16ROWS), ID_PV,
                       (pj > 0 || k > 0) ? 1u : 0u);
        umma_commit(p_empty(g, b));
        if (pj == NU - 1) umma_commit(o_full(g));
        if (head_done) umma_commit(&kv_empty[pk & 1]);          // this group's last read of the stage
      }
      __syncwarp();
      if (++pj == NU) {
        pj = 0;
        if (++pqh == nbg) { pqh = 0; ++pk; }
      }
    }
    if (nbg == 0) {   // a group without query blocks (NB == 1) still has to release the K/V stages
      for (int k = 0; k < n_my; ++k) {
        mbar_wait(&kv_full[k & 1], (uint32_t)((k >> 1) & 1));
        if (lane == 0) mbar_arrive(&kv_empty[k & 1]);
      }
    }
  } else {
    // ============================ softmax groups ============================
    const int g = (warp - TCP_FIRST_SOFTMAX_WARP) >> 2;
    const int nbg = g == 0 ? nbg0 : nbg1;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t tlane = tmem + g * 256 + ((uint32_t)(quarter * 32) << 16);
    const int valid_last = S - (NU - 1) * 64;
    const long long U = (long long)n_my * nbg * NU;
    // token passing: group 0 exponentiates while group 1 does everything else, then they swap.  Both groups make
    // the same number of exchanges (UX per head) even when they own different numbers of query blocks.
    const int UX = nbg0 * NU;
    const int my_bar = 1 + g, other_bar = 2 - g;
    const bool tokens = a.tokens != 0;
    if (tokens && g == 1) named_bar_arrive(1, 256);           // group 0 starts with the token
    float m_run = -INFINITY, l0 = 0.f, l1 = 0.f;
    uint32_t s0[32], s1[32];
    if (U > 0) {
      mbar_wait(s_full(g, 0), 0);
      tc_fence_after();
      tmem_ld_32x32b_x32(tlane + TCF_COL_S, s0);
      tmem_ld_32x32b_x32(tlane + TCF_COL_S + 32, s1);
    }
    auto half_unit = [&](auto mask_tag, const uint32_t (&sr)[32], uint32_t (&pw)[16], int key0) {
      constexpr bool MASKED = decltype(mask_tag)::value;
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        float p0 = exp2f(fmaf(__uint_as_float(sr[c]), a.scale_log2, -m_run));
        float p1 = exp2f(fmaf(__uint_as_float(sr[c + 1]), a.scale_log2, -m_run));
        if (MASKED) {
          if (key0 + c >= valid_last) p0 = 0.f;
          if (key0 + c + 1 >= valid_last) p1 = 0.f;
        }
        l0 += p0;
        l1 += p1;
        pw[c >> 1] = pack_bf16x2(p0, p1);
      }
    };
    long long u = 0;
    for (int k = 0; k < n_my; ++k) {
      const int hh = blockIdx.x + k * gridDim.x, seq = hh / a.H, h = hh - seq * a.H;
      bf16* ob = a.out + (long long)seq * S * a.ld_o + h * HD;
      float* lp = a.lse2 + (long long)hh * S;
      int x = 0;                                   // token exchanges done for this head
      for (int t = 0; t < nbg; ++t) {
        const int i = 2 * t + g;                   // query block
        for (int jj = 0; jj < NU; ++jj, ++u, ++x) {
          const uint32_t b = (uint32_t)(u & 1), bn = b ^ 1u;
          tmem_ld_wait();                          // S(u) is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_read(g, b));
          const bool last = (jj == NU - 1);
          float mx;
          if (!last) {
            float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              a0 = fmaxf(a0, fmaxf(__uint_as_float(s0[c]), __uint_as_float(s0[c + 1])));
              a1 = fmaxf(a1, fmaxf(__uint_as_float(s0[c + 2]), __uint_as_float(s0[c + 3])));
              a2 = fmaxf(a2, fmaxf(__uint_as_float(s1[c]), __uint_as_float(s1[c + 1])));
              a3 = fmaxf(a3, fmaxf(__uint_as_float(s1[c + 2]), __uint_as_float(s1[c + 3])));
            }
            mx = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
          } else {
            mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              if (c < valid_last) mx = fmaxf(mx, __uint_as_float(s0[c]));
              if (c + 32 < valid_last) mx = fmaxf(mx, __uint_as_float(s1[c]));
            }
          }
          const float m_new = mx * a.scale_log2;
          const bool raise = m_new > m_run + TCF_LAZY;
          if (__any_sync(0xffffffffu, raise)) {
            const float m_upd = raise ? m_new : m_run;
            if (jj > 0) {
              const float f = exp2f(m_run - m_upd);
              mbar_wait(p_empty(g, bn), (uint32_t)(((u - 1) >> 1) & 1));
              tc_fence_after();
              uint32_t ro[HD];
              tmem_ld_n<HD>(tlane + TCF_COL_O, ro);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < HD; ++c) ro[c] = __float_as_uint(__uint_as_float(ro[c]) * f);
              tmem_st_32x32b_x32(tlane + TCF_COL_O, ro);
              l0 *= f;
              l1 *= f;
            }
            m_run = m_upd;
          }
          // everything the exponential phase could block on is resolved BEFORE taking the token
          if (u >= 2) {
            mbar_wait(p_empty(g, b), (uint32_t)(((u >> 1) - 1) & 1));
            tc_fence_after();
          }
          if (u + 1 < U) {
            mbar_wait(s_full(g, bn), (uint32_t)(((u + 1) >> 1) & 1));
            tc_fence_after();
          }
          uint32_t pw[16];
          if (tokens) named_bar_sync(my_bar, 256);             // ---- take the MUFU token
          if (last) half_unit(std::true_type{}, s0, pw, 0);
          else half_unit(std::false_type{}, s0, pw, 0);
          tmem_st_32x32b_x16(tlane + TCF_COL_P + b * 32, pw);
          if (u + 1 < U) tmem_ld_32x32b_x32(tlane + TCF_COL_S + bn * 64, s0);
          if (last) half_unit(std::true_type{}, s1, pw, 32);
          else half_unit(std::false_type{}, s1, pw, 32);
          if (tokens) named_bar_arrive(other_bar, 256);        // ---- pass it on
          tmem_st_32x32b_x16(tlane + TCF_COL_P + b * 32 + 16, pw);
          if (u + 1 < U) tmem_ld_32x32b_x32(tlane + TCF_COL_S + bn * 64 + 32, s1);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(g, b));
          if (last) {
            // ---- epilogue of query block i: O / l, log-sum-exp
            const long long nq = (long long)k * nbg + t;
            mbar_wait(o_full(g), (uint32_t)(nq & 1));
            tc_fence_after();
            uint32_t ro[HD];
            tmem_ld_n<HD>(tlane + TCF_COL_O, ro);
            tmem_ld_wait();
            const float l = l0 + l1;
            const float inv = 1.f / l;
            const int qr = i * 128 + row;
            if (qr < S) {
              bf16* dst = ob + (long long)qr * a.ld_o;
#pragma unroll
              for (int c = 0; c < HD; c += 8) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(ro[c]) * inv, __uint_as_float(ro[c + 1]) * inv);
                o.y = pack_bf16x2(__uint_as_float(ro[c + 2]) * inv, __uint_as_float(ro[c + 3]) * inv);
                o.z = pack_bf16x2(__uint_as_float(ro[c + 4]) * inv, __uint_as_float(ro[c + 5]) * inv);
                o.w = pack_bf16x2(__uint_as_float(ro[c + 6]) * inv, __uint_as_float(ro[c + 7]) * inv);
                *reinterpret_cast<uint4*>(dst + c) = o;
              }
              lp[qr] = m_run + log2f(l);
            }
            m_run = -INFINITY; l0 = 0.f; l1 = 0.f;
          }
        }
      }
      for (; tokens && x < UX; ++x) {              // keep the exchange count equal (odd number of query blocks)
        named_bar_sync(my_bar, 256);
        named_bar_arrive(other_bar, 256);
      }
    }
    if (tokens && g == 0) named_bar_sync(1, 256);             // absorb group 1's final hand-over
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int HD>
int launch_fwd(const TcFwdArgs& a, int n_seq, cudaStream_t stream) {
  using C = TcCfg<HD>;
  const int n_last = (((a.S - (a.NU - 1) * 128) + 15) >> 4) << 4;
  // 1 KB of alignment slack; at S = 708: 113.8 KB + 1 KB reserved per CTA, two CTAs = 224.3 KB of the SM's 228 KB
  const int smem = 1024 + 2 * ((a.NU - 1) * 128 + n_last) * C::ROWB + 2 * C::BLK_BYTES + 2048 + 128 + 2048;
  static int smem_set = 0;
  if (smem > smem_set || getenv("AVS_TC_DEBUG")) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      avs_set_error("avs_attention_fwd(tc): cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
    // two CTAs per SM need the full shared-memory carve-out
    cudaFuncSetAttribute(attn_fwd_tc_kernel<HD>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    if (getenv("AVS_TC_DEBUG")) {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, attn_fwd_tc_kernel<HD>, TCF_THREADS, smem);
      fprintf(stderr, "[avs] attn_fwd_tc: smem %d B, %d CTAs per SM\n", smem, nb);
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, attn_fwd_tc_kernel<HD>);
      int dev = 0, smem_sm = 0, smem_res = 0, regs_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
      cudaDeviceGetAttribute(&smem_res, cudaDevAttrReservedSharedMemoryPerBlock, dev);
      cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
      fprintf(stderr, "[avs]   static smem %zu, regs %d, maxThreads %d, maxDyn %d | SM: smem %d, reserved/block %d, regs %d\n",
              fa.sharedSizeBytes, fa.numRegs, fa.maxThreadsPerBlock, fa.maxDynamicSharedSizeBytes, smem_sm, smem_res,
              regs_sm);
      for (int sz : {32768, 65536, 98304, 106496, 110592}) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, attn_fwd_tc_kernel<HD>, TCF_THREADS, sz);
        fprintf(stderr, "[avs]   dyn smem %d -> %d CTAs per SM\n", sz, nb);
      }
    }
    smem_set = smem;
  }
  dim3 grid(a.H, n_seq);
  attn_fwd_tc_kernel<HD><<<grid, TCF_THREADS, smem, stream>>>(a);
  return avs_check_launch("attn_fwd_tc_kernel");
}

long long* g_tc_trace = nullptr;

}  // namespace

extern "C" void avs_debug_set_tc_trace(long long* p) { g_tc_trace = p; }

// Returns -2 when the shape is outside what the TMEM budget covers (the caller then uses the mma.sync kernels of
// attention.cu — both are CUDA paths of this library).  `delta` must already hold rowsum(dO * O).
int avs_attention_bwd_tc(const void* qkv, long long ld_qkv, const void* dout, long long ld_o, const float* lse2,
                         const float* delta, void* dqkv, float* dbias, int n_seq, int S, int H, int head_dim,
                         void* stream) {
  if (head_dim != 32 && head_dim != 64) return -2;
  const int NB = (S + 127) / 128;
  if (NB > (head_dim == 32 ? TcCfg<32>::MAX_NB : TcCfg<64>::MAX_NB)) return -2;
  TcArgs a = {};
  a.trace = g_tc_trace;
  a.qkv = (const bf16*)qkv; a.dout = (const bf16*)dout; a.lse2 = lse2; a.delta = delta; a.dqkv = (bf16*)dqkv; a.dbias = dbias;
  a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.S = S; a.NB = NB; a.H = H; a.D = H * head_dim;
  a.scale = rsqrtf((float)head_dim);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  const long long rows = (long long)n_seq * S;
  const int swz = head_dim == 64 ? 128 : 64;   // one row of the block = one swizzle span
  if (int rc = avs_make_tmap_2d_bf16(&a.map_q, qkv, rows, 3LL * a.D, ld_qkv, head_dim, 128, swz)) return rc;
  if (int rc = avs_make_tmap_2d_bf16(&a.map_do, dout, rows, a.D, ld_o, head_dim, 128, swz)) return rc;
  return head_dim == 64 ? launch_bwd<64>(a, n_seq, (cudaStream_t)stream) : launch_bwd<32>(a, n_seq, (cudaStream_t)stream);
}

// Forward on the tcgen05 path: head_dim 32, 256 <= S <= 768 (K and V of a head resident in shared memory, two CTAs
// per SM).  Returns -2 for other shapes (the caller then uses the mma.sync kernel of attention.cu).
int avs_attention_fwd_tc(const void* qkv, long long ld_qkv, void* out, long long ld_o, float* lse2, int n_seq, int S,
                         int H, int head_dim, void* stream) {
  if (head_dim != 32 || S < 256 || S > 768) return -2;
  static const int variant = getenv("AVS_ATTN_FWD_VARIANT") ? atoi(getenv("AVS_ATTN_FWD_VARIANT")) : 1;
  const int NB = (S + 127) / 128, NU = (S + 63) / 64;
  const float scale_log2 = rsqrtf((float)head_dim) * 1.4426950408889634f;
  if (variant == 2) {
    using C = TcCfg<32>;
    TcPpArgs a = {};
    const long long rows = (long long)n_seq * S;
    int rc = avs_make_tmap_2d_bf16(&a.map128, qkv, rows, 3LL * H * head_dim, ld_qkv, head_dim, 128, 64);
    if (rc) return rc;
    if ((rc = avs_make_tmap_2d_bf16(&a.map16, qkv, rows, 3LL * H * head_dim, ld_qkv, head_dim, 16, 64))) return rc;
    a.out = (bf16*)out; a.lse2 = lse2; a.ld_o = ld_o; a.S = S; a.NB = NB; a.NU = NU; a.H = H; a.D = H * head_dim;
    a.n_heads = n_seq * H;
    a.scale_log2 = scale_log2;
    static const int tokens = getenv("AVS_ATTN_FWD_TOKENS") ? atoi(getenv("AVS_ATTN_FWD_TOKENS")) : 1;
    a.tokens = tokens;
    const int n_last = (((S - (NU - 1) * 64) + 15) >> 4) << 4;
    const int kv_rows = (NU - 1) * 64 + n_last;
    const int smem = 1024 + 4 * kv_rows * C::ROWB + 4 * C::BLK_BYTES + 512;
    if (smem <= 227 * 1024) {
      static int smem_set = 0;
      if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_pp_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
          avs_set_error("avs_attention_fwd(tc): cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
          return (int)e;
        }
        smem_set = smem;
      }
      const int grid = min(avs_num_sms(), a.n_heads);
      attn_fwd_pp_kernel<32><<<grid, TCP_THREADS, smem, (cudaStream_t)stream>>>(a);
      return avs_check_launch("attn_fwd_pp_kernel");
    }
  }
  TcFwdArgs a = {};
  a.qkv = (const bf16*)qkv; a.out = (bf16*)out; a.lse2 = lse2;
  a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.S = S; a.NB = NB; a.NU = (S + 127) / 128; a.H = H; a.D = H * head_dim;   // 128-key units
  a.scale_log2 = scale_log2;
  return launch_fwd<32>(a, n_seq, (cudaStream_t)stream);
}
