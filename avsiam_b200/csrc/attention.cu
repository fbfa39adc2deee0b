// Fused (flash-style) multi-head self-attention forward + backward for the short sequences of the AVSiam
// encoder / MAE decoder (S = 49..913, head_dim 32, 64 or 80).  Reads q,k,v straight out of the packed QKV GEMM
// output [tokens, 3*D] and writes O as [tokens, D] / dQKV as [tokens, 3*D], so there are no permute copies on
// either side, and never materialises the [S,S] score matrix in HBM.
//
// Replaces F.scaled_dot_product_attention + the reshape/permute/transpose around it in Attention.forward
// (cav_mae_base.py:58-77): scale = head_dim^-0.5, no mask, dropout 0.
//
// Design (round 1, v2).  At head_dim 32 (the decoder: 54 % of the MAE FLOPs) attention is bound by the softmax
// exponentials (MUFU, 16/clk/SM), not by the tensor pipe: a 128x128 score block costs 256 tensor cycles on
// tcgen05 but 1024 MUFU cycles.  So the kernel is organised around keeping the SFU/FMA pipes busy:
//   * one CTA per (sequence, head); the WHOLE head's K and V (<= 90 KB) are brought into shared memory once with
//     cp.async (XOR-swizzled 16-byte chunks, conflict-free for ldmatrix) and reused by every query tile — no
//     per-tile reloads, no barriers inside the main loop;
//   * 8 warps, each owning 16-query tiles round-robin; scores / probabilities live in registers
//     (mma.sync.m16n8k16 bf16, fp32 accumulate), B fragments come from ldmatrix.x4 (1 LDSM per 2 MMAs);
//   * online softmax in the exp2 domain with fp32 statistics.
// Backward = delta pre-pass + a dK/dV kernel (warp owns 16 keys, the head's Q and dO resident in smem) + a dQ
// kernel (warp owns 16 queries, K and V resident), each recomputing P from the saved log-sum-exp.
#include "../../include/avsiam_b200.h"
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace {


__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// A head tile in shared memory: rows of HD bf16 (HD*2 bytes, no padding), 16-byte chunks XOR-swizzled by row so
// that the 8 row addresses of an ldmatrix 8x8 fetch hit 8 distinct bank groups.
template <int HD>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  constexpr int CPR = HD / 8;  // 16-byte chunks per row (4, 8 or 10)
  // head_dim 80 (ViT-H): a row is 10 chunks, so 8 consecutive rows start 2 bank groups apart (0,2,4,6,0,2,4,6);
  // flipping the lowest chunk bit for rows 4..7 of every 8 moves them onto the odd groups
  const int sw = (CPR == 8) ? (row & 7) : (CPR == 4) ? ((row >> 1) & 3) : ((row >> 2) & 1);
  return (uint32_t)(row * (HD * 2) + ((chunk ^ sw) << 4));
}

// Whole-head load: rows [0,S) of a [*, ld] bf16 matrix (HD columns starting at src) -> swizzled smem tile; rows
// [S, S_pad) are zero-filled.  Asynchronous (cp.async); caller waits + __syncthreads.
template <int HD, int NTHREADS>
__device__ __forceinline__ void load_head_async(uint32_t tile, const bf16* __restrict__ src, long long ld, int S,
                                                int S_pad) {
  constexpr int CPR = HD / 8;
  for (int i = threadIdx.x; i < S_pad * CPR; i += NTHREADS) {
    const int r = i / CPR, c = i % CPR;
    const uint32_t dst = tile + tile_off<HD>(r, c);
    if (r < S) cp_async16(dst, src + (long long)r * ld + c * 8);
    else asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};\n" ::"r"(dst), "r"(0u) : "memory");
  }
}

// A-operand fragments of 16 rows (row0 + lane/4, +8) x HD straight from global memory; rows >= S read as zero.
template <int HD>
__device__ __forceinline__ void load_a_frags_global(uint32_t (&a)[HD / 16][4], const bf16* __restrict__ base,
                                                    long long ld, int row0, int S, int lane) {
  const int r0 = row0 + (lane >> 2), r1 = r0 + 8, c = (lane & 3) * 2;
  const bf16* p0 = base + (long long)r0 * ld + c;
  const bf16* p1 = base + (long long)r1 * ld + c;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    a[ks][0] = r0 < S ? *reinterpret_cast<const uint32_t*>(p0 + ks * 16) : 0u;
    a[ks][1] = r1 < S ? *reinterpret_cast<const uint32_t*>(p1 + ks * 16) : 0u;
    a[ks][2] = r0 < S ? *reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 8) : 0u;
    a[ks][3] = r1 < S ? *reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 8) : 0u;
  }
}

// acc[nt] (16 x 8*NT) = A(16 x HD) * T[row0 .. row0+8*NT)^T     (T rows are the n index: "col-major B")
template <int HD, int NT>
__device__ __forceinline__ void gemm_a_tT(float (&acc)[NT][4], const uint32_t (&a)[HD / 16][4], uint32_t tile,
                                          int row0, int lane) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
    const int row = row0 + nt * 8 + (lane & 7);
#pragma unroll
    for (int half = 0; half < HD / 32; ++half) {
      uint32_t b[4];
      ldsm_x4(b, tile + tile_off<HD>(row, half * 4 + (lane >> 3)));
      mma16816(acc[nt], a[2 * half], b[0], b[1]);
      mma16816(acc[nt], a[2 * half + 1], b[2], b[3]);
    }
    if constexpr ((HD / 16) % 2 == 1) {   // head_dim 80: the fifth 16-column k-step (chunks 8, 9; lanes 16.. repeat them)
      uint32_t b[4];
      ldsm_x4(b, tile + tile_off<HD>(row, (HD / 32) * 4 + ((lane >> 3) & 1)));
      mma16816(acc[nt], a[HD / 16 - 1], b[0], b[1]);
    }
  }
}

// out[dn] (16 x HD) += P(16 x 16*KK, packed bf16 A-frags) * T[row0 .. row0+16*KK)     (T rows are the k index)
template <int HD, int KK>
__device__ __forceinline__ void gemm_p_t(float (&out)[HD / 8][4], const uint32_t (&p)[KK][4], uint32_t tile, int row0,
                                         int lane) {
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
    const int row = row0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int dp = 0; dp < HD / 16; ++dp) {
      uint32_t b[4];
      ldsm_x4_trans(b, tile + tile_off<HD>(row, dp * 2 + (lane >> 4)));
      mma16816(out[2 * dp], p[kk], b[0], b[1]);
      mma16816(out[2 * dp + 1], p[kk], b[2], b[3]);
    }
  }
}

// pack a 16 x 8*NT fp32 C-fragment set into bf16 A-fragments (the FA2 register trick)
template <int NT>
__device__ __forceinline__ void pack_c_to_a(uint32_t (&p)[NT / 2][4], const float (&s)[NT][4]) {
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    p[kk][0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    p[kk][1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    p[kk][2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    p[kk][3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}

struct AttnArgs {
  const bf16* qkv;     // [rows, ld_qkv]: q | k | v, each D wide, head h at h*HD
  bf16* out;           // fwd: O [rows, ld_o]
  float* lse2;         // [n_seq, H, S] log2-domain logsumexp
  const bf16* dout;    // bwd: dO [rows, ld_o]
  const float* delta;  // bwd: [n_seq, H, S]  rowsum(dO * O)
  bf16* dqkv;          // bwd: [rows, ld_qkv]
  long long ld_qkv, ld_o;
  int S, S_pad, H, D;
  float scale_log2;    // scale * log2(e)
  float scale;
};

// ------------------------------------------------ forward ------------------------------------------------
// WARPS = 8 normally; 4 for sequences of at most 64 tokens (4 x 16-row tiles), so that no warp of a resident CTA idles
template <int HD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, HD == 80 ? 1 : 256 / (WARPS * 32) * 2) attn_fwd_kernel(const AttnArgs a) {
  constexpr int NT = 8;  // 64 keys per inner block
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, seq = blockIdx.y;
  const int S = a.S, S_pad = a.S_pad;
  const long long row_base = (long long)seq * S;
  const bf16* qb = a.qkv + row_base * a.ld_qkv + h * HD;
  const uint32_t sK = smem_u32(att_smem), sV = sK + S_pad * HD * 2;
  load_head_async<HD, WARPS * 32>(sK, qb + a.D, a.ld_qkv, S, S_pad);
  load_head_async<HD, WARPS * 32>(sV, qb + 2 * a.D, a.ld_qkv, S, S_pad);
  cp_async_wait_all();
  __syncthreads();

  bf16* ob = a.out + row_base * a.ld_o + h * HD;
  float* lp = a.lse2 + ((long long)seq * a.H + h) * S;
  const int n_qt = (S + 15) >> 4;
  for (int qt = warp; qt < n_qt; qt += WARPS) {
    uint32_t qf[HD / 16][4];
    load_a_frags_global<HD>(qf, qb, a.ld_qkv, qt * 16, S, lane);
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    // one block of 8 * NTB keys: scores, online softmax, P V.  Full blocks take 64 keys; the last block of the
    // sequence only as many 16-key steps as it has valid keys (S = 708: 4 valid keys in the 12th block -> NTB = 2,
    // a quarter of the work instead of a full block of mostly masked columns).
    auto block = [&](auto ntb_tag, int kv0) {
      constexpr int NTB = decltype(ntb_tag)::value;
      float s[NTB][4];
      gemm_a_tT<HD, NTB>(s, qf, sK, kv0, lane);
      const bool tail = kv0 + 8 * NTB > S;
      float mx0 = -INFINITY, mx1 = -INFINITY;
      // the row max is taken on the raw scores (scale > 0 commutes with max), so that scaling and max subtraction
      // are ONE FFMA per element in the exponent below
#pragma unroll
      for (int nt = 0; nt < NTB; ++nt) {
        if (tail) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((kv0 + nt * 8 + (lane & 3) * 2 + (j & 1)) >= S) s[nt][j] = -INFINITY;
        }
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float nm0 = fmaxf(m0, mx0 * a.scale_log2), nm1 = fmaxf(m1, mx1 * a.scale_log2);
      const float corr0 = exp2f(m0 - nm0), corr1 = exp2f(m1 - nm1);
      m0 = nm0; m1 = nm1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NTB; ++nt) {
        s[nt][0] = exp2f(fmaf(s[nt][0], a.scale_log2, -m0)); s[nt][1] = exp2f(fmaf(s[nt][1], a.scale_log2, -m0));
        s[nt][2] = exp2f(fmaf(s[nt][2], a.scale_log2, -m1)); s[nt][3] = exp2f(fmaf(s[nt][3], a.scale_log2, -m1));
        rs0 += s[nt][0] + s[nt][1];
        rs1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * corr0 + rs0;
      l1 = l1 * corr1 + rs1;
#pragma unroll
      for (int dn = 0; dn < HD / 8; ++dn) {
        o[dn][0] *= corr0; o[dn][1] *= corr0;
        o[dn][2] *= corr1; o[dn][3] *= corr1;
      }
      uint32_t p[NTB / 2][4];
      pack_c_to_a<NTB>(p, s);
      gemm_p_t<HD, NTB / 2>(o, p, sV, kv0, lane);
    };
    int kv0 = 0;
    for (; kv0 + 8 * NT <= S; kv0 += 8 * NT) block(std::integral_constant<int, NT>{}, kv0);
    if (kv0 < S) {
      const int rem = S - kv0;                       // 1 .. 63 valid keys left (tile rows up to S_pad are zero-filled)
      if constexpr (HD == 32) {                      // (head_dim 64 has no registers to spare for more block shapes)
        if (rem <= 16) block(std::integral_constant<int, 2>{}, kv0);
        else if (rem <= 32) block(std::integral_constant<int, 4>{}, kv0);
        else block(std::integral_constant<int, NT>{}, kv0);
      } else {
        block(std::integral_constant<int, NT>{}, kv0);
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    const int r0 = qt * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
      const int c = dn * 8 + (lane & 3) * 2;
      if (r0 < S) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * a.ld_o + c) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
      if (r1 < S) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * a.ld_o + c) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
    if ((lane & 3) == 0) {
      if (r0 < S) lp[r0] = m0 + log2f(l0);
      if (r1 < S) lp[r1] = m1 + log2f(l1);
    }
  }
}

// delta[seq,h,t] = sum_d dO * O.  HBM-bound pre-pass: each thread takes one 16-byte chunk (8 columns) of O and dO,
// HD/8 adjacent lanes reduce with shuffles; fully coalesced 128-bit loads.
template <int HD>
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                                         float* __restrict__ delta, long long ld_o, int S, int H,
                                                         long long rows) {
  constexpr int LPH = HD / 8;                       // lanes per head (4 or 8)
  const int cpr = H * LPH;                          // 16-byte chunks per row
  const long long total = rows * cpr;
  // total is a multiple of LPH and the stride is a multiple of 32, so all lanes of a shuffle group stay together
  // (32-bit index arithmetic: the launcher checks rows * chunks < 2^31; 64-bit divisions were most of this kernel's
  //  instructions)
  const int total_i = (int)total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ((total_i + 31) & ~31); i += gridDim.x * blockDim.x) {
    float acc = 0.f;
    const bool ok = i < total_i;
    int row = 0;
    int c = 0;
    if (ok) {
      row = i / cpr;
      c = i - row * cpr;
      const uint4 x = *reinterpret_cast<const uint4*>(o + (long long)row * ld_o + c * 8);
      const uint4 y = *reinterpret_cast<const uint4*>(dout + (long long)row * ld_o + c * 8);
      float2 a, b;
      a = unpack_bf16x2(x.x); b = unpack_bf16x2(y.x); acc += a.x * b.x + a.y * b.y;
      a = unpack_bf16x2(x.y); b = unpack_bf16x2(y.y); acc += a.x * b.x + a.y * b.y;
      a = unpack_bf16x2(x.z); b = unpack_bf16x2(y.z); acc += a.x * b.x + a.y * b.y;
      a = unpack_bf16x2(x.w); b = unpack_bf16x2(y.w); acc += a.x * b.x + a.y * b.y;
    }
#pragma unroll
    for (int off = LPH / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (ok && (c % LPH) == 0) {
      const int h = c / LPH;
      const int sq = row / S;
      delta[((long long)sq * H + h) * S + (row - sq * S)] = acc;
    }
  }
}

// Same for head dims whose 16-byte chunk count is not a power of two (80 = 10 chunks, ViT-H): one thread per
// (row, head) walks the head's chunks; a warp's loads still cover one contiguous 32 * 2*HD-byte stretch of the row.
template <int HD>
__global__ void __launch_bounds__(256) attn_delta_rowhead_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                                                 float* __restrict__ delta, long long ld_o, int S, int H,
                                                                 long long rows) {
  const long long total = rows * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / H;
    const int h = (int)(i - row * H);
    const bf16* po = o + row * ld_o + h * HD;
    const bf16* pd = dout + row * ld_o + h * HD;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 x = *reinterpret_cast<const uint4*>(po + c * 8);
      const uint4 y = *reinterpret_cast<const uint4*>(pd + c * 8);
      float2 a, b;
      a = unpack_bf16x2(x.x); b = unpack_bf16x2(y.x); acc += a.x * b.x + a.y * b.y;
      a = unpack_bf16x2(x.y); b = unpack_bf16x2(y.y); acc += a.x * b.x + a.y * b.y;
      a = unpack_bf16x2(x.z); b = unpack_bf16x2(y.z); acc += a.x * b.x + a.y * b.y;
      a = unpack_bf16x2(x.w); b = unpack_bf16x2(y.w); acc += a.x * b.x + a.y * b.y;
    }
    const long long sq = row / S;
    delta[(sq * H + h) * S + (row - sq * S)] = acc;
  }
}

// ------------------------------------------------ backward: dQ ------------------------------------------------
// warp owns 16 queries; K and V of the head resident in smem; dQ = sum_blocks dS * K
// WARPS = 8 normally; 4 for sequences of at most 64 tokens (4 x 16-row tiles), so that no warp of a resident CTA idles
template <int HD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, HD == 80 ? 1 : 256 / (WARPS * 32) * 2) attn_bwd_dq_kernel(const AttnArgs a) {
  constexpr int NT = (HD >= 64) ? 4 : 8;  // keys per inner block / 8 (register budget: 128/thread)
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, seq = blockIdx.y;
  const int S = a.S, S_pad = a.S_pad;
  const long long row_base = (long long)seq * S;
  const bf16* qb = a.qkv + row_base * a.ld_qkv + h * HD;
  const bf16* dob = a.dout + row_base * a.ld_o + h * HD;
  const uint32_t sK = smem_u32(att_smem), sV = sK + S_pad * HD * 2;
  load_head_async<HD, WARPS * 32>(sK, qb + a.D, a.ld_qkv, S, S_pad);
  load_head_async<HD, WARPS * 32>(sV, qb + 2 * a.D, a.ld_qkv, S, S_pad);
  cp_async_wait_all();
  __syncthreads();

  const long long sb = ((long long)seq * a.H + h) * S;
  bf16* dqb = a.dqkv + row_base * a.ld_qkv + h * HD;
  const int n_qt = (S + 15) >> 4;
  for (int qt = warp; qt < n_qt; qt += WARPS) {
    uint32_t qf[HD / 16][4], dof[HD / 16][4];
    load_a_frags_global<HD>(qf, qb, a.ld_qkv, qt * 16, S, lane);
    load_a_frags_global<HD>(dof, dob, a.ld_o, qt * 16, S, lane);
    const int r0 = qt * 16 + (lane >> 2), r1 = r0 + 8;
    const float lse0 = r0 < S ? a.lse2[sb + r0] : 0.f, lse1 = r1 < S ? a.lse2[sb + r1] : 0.f;
    const float dl0 = r0 < S ? a.delta[sb + r0] : 0.f, dl1 = r1 < S ? a.delta[sb + r1] : 0.f;
    float dq[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;

    for (int kv0 = 0; kv0 < S_pad; kv0 += 8 * NT) {
      float s[NT][4], dp[NT][4];
      gemm_a_tT<HD, NT>(s, qf, sK, kv0, lane);
      gemm_a_tT<HD, NT>(dp, dof, sV, kv0, lane);
      const bool tail = kv0 + 8 * NT > S;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float lse = (j < 2) ? lse0 : lse1, dl = (j < 2) ? dl0 : dl1;
          float p = exp2f(s[nt][j] * a.scale_log2 - lse);
          if (tail && (kv0 + nt * 8 + (lane & 3) * 2 + (j & 1)) >= S) p = 0.f;
          s[nt][j] = p * (dp[nt][j] - dl) * a.scale;  // dS * scale
        }
      }
      uint32_t ds[NT / 2][4];
      pack_c_to_a<NT>(ds, s);
      gemm_p_t<HD, NT / 2>(dq, ds, sK, kv0, lane);
    }
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
      const int c = dn * 8 + (lane & 3) * 2;
      if (r0 < S) *reinterpret_cast<uint32_t*>(dqb + (long long)r0 * a.ld_qkv + c) = pack_bf16x2(dq[dn][0], dq[dn][1]);
      if (r1 < S) *reinterpret_cast<uint32_t*>(dqb + (long long)r1 * a.ld_qkv + c) = pack_bf16x2(dq[dn][2], dq[dn][3]);
    }
  }
}

// ------------------------------------------------ backward: dK, dV ------------------------------------------------
// warp owns 16 keys; Q and dO of the head (+ lse, delta) resident in smem
// WARPS = 8 normally; 4 for sequences of at most 64 tokens (4 x 16-row tiles), so that no warp of a resident CTA idles
template <int HD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, HD == 80 ? 1 : 256 / (WARPS * 32) * 2) attn_bwd_dkv_kernel(const AttnArgs a) {
  constexpr int NT = (HD >= 64) ? 4 : 8;  // queries per inner block / 8
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, seq = blockIdx.y;
  const int S = a.S, S_pad = a.S_pad;
  const long long row_base = (long long)seq * S;
  const bf16* qb = a.qkv + row_base * a.ld_qkv + h * HD;
  const bf16* dob = a.dout + row_base * a.ld_o + h * HD;
  const uint32_t sQ = smem_u32(att_smem), sDO = sQ + S_pad * HD * 2;
  float* s_lse = reinterpret_cast<float*>(att_smem + 2 * S_pad * HD * 2);
  float* s_delta = s_lse + S_pad;
  load_head_async<HD, WARPS * 32>(sQ, qb, a.ld_qkv, S, S_pad);
  load_head_async<HD, WARPS * 32>(sDO, dob, a.ld_o, S, S_pad);
  const long long sb = ((long long)seq * a.H + h) * S;
  for (int i = threadIdx.x; i < S_pad; i += WARPS * 32) {
    s_lse[i] = i < S ? a.lse2[sb + i] : INFINITY;  // padded queries: p = exp2(-inf) = 0
    s_delta[i] = i < S ? a.delta[sb + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  bf16* dkb = a.dqkv + row_base * a.ld_qkv + a.D + h * HD;
  bf16* dvb = a.dqkv + row_base * a.ld_qkv + 2 * a.D + h * HD;
  const int n_kt = (S + 15) >> 4;
  for (int kt = warp; kt < n_kt; kt += WARPS) {
    uint32_t kf[HD / 16][4], vf[HD / 16][4];
    load_a_frags_global<HD>(kf, qb + a.D, a.ld_qkv, kt * 16, S, lane);
    load_a_frags_global<HD>(vf, qb + 2 * a.D, a.ld_qkv, kt * 16, S, lane);
    float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dk[i][j] = 0.f, dv[i][j] = 0.f;

    for (int q0 = 0; q0 < S_pad; q0 += 8 * NT) {
      float st[NT][4], dpt[NT][4];
      gemm_a_tT<HD, NT>(st, kf, sQ, q0, lane);     // S^T  [keys x q]
      gemm_a_tT<HD, NT>(dpt, vf, sDO, q0, lane);   // dP^T [keys x q]
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int q = q0 + nt * 8 + (lane & 3) * 2;
        const float2 ls = *reinterpret_cast<const float2*>(s_lse + q);
        const float2 dl = *reinterpret_cast<const float2*>(s_delta + q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p = exp2f(st[nt][j] * a.scale_log2 - ((j & 1) ? ls.y : ls.x));
          st[nt][j] = p;
          dpt[nt][j] = p * (dpt[nt][j] - ((j & 1) ? dl.y : dl.x)) * a.scale;  // dS^T * scale
        }
      }
      uint32_t pt[NT / 2][4];
      pack_c_to_a<NT>(pt, st);
      gemm_p_t<HD, NT / 2>(dv, pt, sDO, q0, lane);  // dV += P^T dO
      pack_c_to_a<NT>(pt, dpt);
      gemm_p_t<HD, NT / 2>(dk, pt, sQ, q0, lane);   // dK += dS^T Q
    }
    const int r0 = kt * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
      const int c = dn * 8 + (lane & 3) * 2;
      if (r0 < S) {
        *reinterpret_cast<uint32_t*>(dkb + (long long)r0 * a.ld_qkv + c) = pack_bf16x2(dk[dn][0], dk[dn][1]);
        *reinterpret_cast<uint32_t*>(dvb + (long long)r0 * a.ld_qkv + c) = pack_bf16x2(dv[dn][0], dv[dn][1]);
      }
      if (r1 < S) {
        *reinterpret_cast<uint32_t*>(dkb + (long long)r1 * a.ld_qkv + c) = pack_bf16x2(dk[dn][2], dk[dn][3]);
        *reinterpret_cast<uint32_t*>(dvb + (long long)r1 * a.ld_qkv + c) = pack_bf16x2(dv[dn][2], dv[dn][3]);
      }
    }
  }
}

constexpr int ATT_SMEM_MAX = 227 * 1024;

template <typename K>
int set_smem(K kernel, int bytes, const char* who) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    avs_set_error("%s: cudaFuncSetAttribute(smem=%d): %s", who, bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int check_common(const void* qkv, long long ld_qkv, long long ld_o, int n_seq, int S, int H, int HD, const char* who) {
  AVS_REQUIRE(qkv != nullptr, "%s: null pointer", who);
  AVS_REQUIRE(HD == 32 || HD == 64 || HD == 80, "%s: head_dim must be 32, 64 or 80 (got %d)", who, HD);
  AVS_REQUIRE(S > 0 && H > 0 && n_seq >= 0 && n_seq <= 65535, "%s: bad shape", who);
  AVS_REQUIRE(ld_qkv % 8 == 0 && ld_o % 8 == 0 && ((uintptr_t)qkv & 15) == 0, "%s: 16-byte alignment", who);
  const int S_pad = (S + 63) & ~63;
  AVS_REQUIRE(2 * S_pad * HD * 2 + 8 * S_pad <= ATT_SMEM_MAX,
              "%s: sequence too long for the head-resident kernel (S=%d, head_dim=%d)", who, S, HD);
  return 0;
}

// head_dim 80 runs at one 8-warp or two 4-warp CTAs per SM (registers): with 16-row tiles dealt round-robin to the warps,
// S = 164 (ViT-H's kept audio tokens: 11 tiles) leaves 5 of 16 warp-slots idle with 8 warps and 1 of 12 with 4.
bool few_warps_balance_better(int S) {
  const int tiles = (S + 15) / 16;
  const int waste4 = (tiles + 3) / 4 * 4 - tiles, waste8 = (tiles + 7) / 8 * 8 - tiles;
  return waste4 * (long long)((tiles + 7) / 8 * 8) < waste8 * (long long)((tiles + 3) / 4 * 4);   // idle fraction
}

}  // namespace

bool avs_attention_tc_enabled() {
  static const bool on = [] {
    const char* e = getenv("AVS_ATTN_TC");
    return !(e && e[0] == '0');
  }();
  return on;
}

extern "C" int avs_attention_fwd(const void* qkv, long long ld_qkv, void* out, long long ld_o, float* lse2, int n_seq,
                                 int S, int H, int head_dim, void* stream) {
  if (check_common(qkv, ld_qkv, ld_o, n_seq, S, H, head_dim, "avs_attention_fwd")) return -1;
  AVS_REQUIRE(out && lse2, "avs_attention_fwd: null pointer");
  if (n_seq == 0) return 0;
  if (avs_attention_tc_enabled()) {   // encoder shapes (head_dim 64, S <= 128): persistent single-tile tcgen05 kernel
    const int rc = avs_attention_small_fwd(qkv, ld_qkv, out, ld_o, lse2, n_seq, S, H, head_dim, stream);
    if (rc != -2) return rc;
  }
  static const bool fwd_tc = avs_attention_tc_enabled() && !(getenv("AVS_ATTN_FWD_TC") && atoi(getenv("AVS_ATTN_FWD_TC")) == 0);
  if (fwd_tc) {   // long head_dim-32 sequences (MAE decoder): tcgen05 kernel, exponential-bound instead of issue-bound
    const int rc = avs_attention_fwd_tc(qkv, ld_qkv, out, ld_o, lse2, n_seq, S, H, head_dim, stream);
    if (rc != -2) return rc;
  }
  AttnArgs a = {};
  a.qkv = (const bf16*)qkv; a.out = (bf16*)out; a.lse2 = lse2;
  a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.S = S; a.S_pad = (S + 63) & ~63; a.H = H; a.D = H * head_dim;
  a.scale = rsqrtf((float)head_dim);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  const int smem = 2 * a.S_pad * head_dim * 2;
  dim3 grid(H, n_seq);
  int rc;
  const bool small = S <= 64;                       // 4 query tiles: 4-warp CTAs, twice as many resident per SM
#define AVS_LAUNCH_FWD(HD_, W_)                                                                      \
  do {                                                                                               \
    if ((rc = set_smem(attn_fwd_kernel<HD_, W_>, smem, "avs_attention_fwd"))) return rc;             \
    attn_fwd_kernel<HD_, W_><<<grid, W_ * 32, smem, (cudaStream_t)stream>>>(a);                      \
  } while (0)
  if (head_dim == 80) {   // ViT-H/14: 5 MMA k-steps per row, ~230 registers per thread
    if (small || few_warps_balance_better(S)) AVS_LAUNCH_FWD(80, 4); else AVS_LAUNCH_FWD(80, 8);
  } else if (head_dim == 64) {
    if (small) AVS_LAUNCH_FWD(64, 4); else AVS_LAUNCH_FWD(64, 8);
  } else {
    if (small) AVS_LAUNCH_FWD(32, 4); else AVS_LAUNCH_FWD(32, 8);
  }
#undef AVS_LAUNCH_FWD
  return avs_check_launch("attn_fwd_kernel");
}

// delta: scratch [n_seq, H, S] fp32
extern "C" int avs_attention_bwd(const void* qkv, long long ld_qkv, const void* out, const void* dout, long long ld_o,
                                 const float* lse2, float* delta, void* dqkv, float* dbias, int n_seq, int S, int H,
                                 int head_dim, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (check_common(qkv, ld_qkv, ld_o, n_seq, S, H, head_dim, "avs_attention_bwd")) return -1;
  AVS_REQUIRE(out && dout && lse2 && delta && dqkv, "avs_attention_bwd: null pointer");
  if (n_seq == 0) return 0;
  if (avs_attention_tc_enabled()) {   // encoder shapes: one kernel, delta taken inside it (no pre-pass over O / dO)
    const int rc = avs_attention_small_bwd(qkv, ld_qkv, dout, ld_o, lse2, dqkv, dbias, n_seq, S, H, head_dim, stream_);
    if (rc != -2) return rc;
  }
  AttnArgs a = {};
  a.qkv = (const bf16*)qkv; a.dout = (const bf16*)dout; a.lse2 = const_cast<float*>(lse2); a.delta = delta;
  a.dqkv = (bf16*)dqkv;
  a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.S = S; a.S_pad = (S + 63) & ~63; a.H = H; a.D = H * head_dim;
  a.scale = rsqrtf((float)head_dim);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  const long long rows = (long long)n_seq * S;
  const long long chunks = rows * H * (head_dim / 8);
  AVS_REQUIRE(chunks + 32 < (1ll << 31), "avs_attention_bwd: too many rows for the delta pre-pass");
  const int dblocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll(chunks, 256));
  if (head_dim == 80) attn_delta_rowhead_kernel<80><<<dblocks, 256, 0, stream>>>((const bf16*)out, (const bf16*)dout, delta, ld_o, S, H, rows);
  else if (head_dim == 64) attn_delta_kernel<64><<<dblocks, 256, 0, stream>>>((const bf16*)out, (const bf16*)dout, delta, ld_o, S, H, rows);
  else attn_delta_kernel<32><<<dblocks, 256, 0, stream>>>((const bf16*)out, (const bf16*)dout, delta, ld_o, S, H, rows);
  int rc = avs_check_launch("attn_delta_kernel");
  if (rc) return rc;
  if (avs_attention_tc_enabled() && S >= 96) {  // short sequences (video, S = 49): the mma.sync kernels win
    rc = avs_attention_bwd_tc(qkv, ld_qkv, dout, ld_o, lse2, delta, dqkv, dbias, n_seq, S, H, head_dim, stream_);
    if (rc != -2) return rc;
  }
  const int smem_dq = 2 * a.S_pad * head_dim * 2;
  const int smem_dkv = smem_dq + 2 * a.S_pad * 4;
  dim3 grid(H, n_seq);
  const bool small = S <= 64;                       // 4 tiles: 4-warp CTAs, twice as many resident per SM
#define AVS_LAUNCH_BWD(HD_, W_)                                                                      \
  do {                                                                                               \
    if ((rc = set_smem(attn_bwd_dkv_kernel<HD_, W_>, smem_dkv, "avs_attention_bwd"))) return rc;     \
    if ((rc = set_smem(attn_bwd_dq_kernel<HD_, W_>, smem_dq, "avs_attention_bwd"))) return rc;       \
    attn_bwd_dkv_kernel<HD_, W_><<<grid, W_ * 32, smem_dkv, stream>>>(a);                            \
    if ((rc = avs_check_launch("attn_bwd_dkv_kernel"))) return rc;                                   \
    attn_bwd_dq_kernel<HD_, W_><<<grid, W_ * 32, smem_dq, stream>>>(a);                              \
  } while (0)
  if (head_dim == 80) {
    if (small || few_warps_balance_better(S)) AVS_LAUNCH_BWD(80, 4); else AVS_LAUNCH_BWD(80, 8);
  } else if (head_dim == 64) {
    if (small) AVS_LAUNCH_BWD(64, 4); else AVS_LAUNCH_BWD(64, 8);
  } else {
    if (small) AVS_LAUNCH_BWD(32, 4); else AVS_LAUNCH_BWD(32, 8);
  }
#undef AVS_LAUNCH_BWD
  if ((rc = avs_check_launch("attn_bwd_dq_kernel"))) return rc;
  // qkv-bias gradient: the tcgen05 kernel accumulates it in its epilogues; this path takes one more pass over dQKV
  if (dbias != nullptr) return avs_colsum_bf16(dqkv, ld_qkv, dbias, n_seq * S, 3 * a.D, 1.0f, stream_);
  return 0;
}
