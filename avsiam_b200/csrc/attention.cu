// Fused (flash-style) multi-head self-attention forward + backward for the short sequences of the AVSiam
// encoder / MAE decoder (S = 49..708, head_dim 32 or 64).  Reads q,k,v straight out of the packed QKV GEMM
// output [tokens, 3*D] and writes O as [tokens, D] / dQKV as [tokens, 3*D], so there are no permute copies on
// either side.  Online softmax in fp32 (exp2 domain), bf16 mma.sync.m16n8k16 tensor-core tiles, never
// materialises the [S,S] score matrix in HBM.
//
// Replaces F.scaled_dot_product_attention + the reshape/permute/transpose around it in Attention.forward
// (cav_mae_base.py:58-77): scale = head_dim^-0.5, no mask, dropout 0.
//
// Round-1 note: this kernel uses the legacy mma.sync path (HMMA); attention is ~11 % of the step FLOPs. A
// tcgen05/TMEM version is the next optimisation.
#include "../../include/avsiam_b200.h"
#include "common.cuh"

namespace {

constexpr int BM = 64;  // query rows per CTA (4 warps x 16)
constexpr int BN = 64;  // keys per inner tile
constexpr int ATT_THREADS = 128;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r0), "=r"(r1)
               : "r"(smem_u32(p)));
}

template <int HD>
struct Tile {
  static constexpr int LDS = HD + 8;  // padded row pitch (elements): conflict-free fragment loads
  bf16 d[BM][LDS];
};

// cooperative load of a [64 x HD] tile; rows >= valid_rows are zero-filled
template <int HD>
__device__ __forceinline__ void load_tile(Tile<HD>& t, const bf16* __restrict__ src, long long ld, int valid_rows) {
  constexpr int VPR = HD / 8;  // 16-byte vectors per row
  for (int i = threadIdx.x; i < BM * VPR; i += ATT_THREADS) {
    const int r = i / VPR, c = (i % VPR) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < valid_rows) v = *reinterpret_cast<const uint4*>(src + (long long)r * ld + c);
    *reinterpret_cast<uint4*>(&t.d[r][c]) = v;
  }
}

// A fragments (16 rows of this warp x HD) from a smem tile
template <int HD>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[HD / 16][4], const Tile<HD>& t, int warp, int lane) {
  const int r = warp * 16 + (lane >> 2), c = (lane & 3) * 2;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    a[ks][0] = *reinterpret_cast<const uint32_t*>(&t.d[r][ks * 16 + c]);
    a[ks][1] = *reinterpret_cast<const uint32_t*>(&t.d[r + 8][ks * 16 + c]);
    a[ks][2] = *reinterpret_cast<const uint32_t*>(&t.d[r][ks * 16 + c + 8]);
    a[ks][3] = *reinterpret_cast<const uint32_t*>(&t.d[r + 8][ks * 16 + c + 8]);
  }
}

// acc[nt] (16 x 64 over 8 n-tiles) = A(16 x HD) * T^T, T = [64 x HD] row-major tile ("col-major B")
template <int HD>
__device__ __forceinline__ void gemm_a_tT(float (&acc)[8][4], const uint32_t (&a)[HD / 16][4], const Tile<HD>& t,
                                          int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
    const int n = nt * 8 + (lane >> 2), c = (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&t.d[n][ks * 16 + c]);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&t.d[n][ks * 16 + c + 8]);
      mma16816(acc[nt], a[ks], b0, b1);
    }
  }
}

// out[dn] (16 x HD) += P(16 x 64, packed bf16 A-frags) * T,  T = [64 x HD] row-major (B via ldmatrix.trans)
template <int HD>
__device__ __forceinline__ void gemm_p_t(float (&out)[HD / 8][4], const uint32_t (&p)[4][4], const Tile<HD>& t,
                                         int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
      uint32_t b0, b1;
      ldsm_x2_trans(b0, b1, &t.d[kk * 16 + (lane & 15)][dn * 8]);
      mma16816(out[dn], p[kk], b0, b1);
    }
  }
}

// pack a 16x64 fp32 C-fragment set into bf16 A-fragments (the FA2 register trick)
__device__ __forceinline__ void pack_c_to_a(uint32_t (&p)[4][4], const float (&s)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    p[kk][0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    p[kk][1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    p[kk][2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    p[kk][3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}

struct AttnArgs {
  const bf16* qkv;   // [rows, ld_qkv]: q | k | v, each D wide, head h at h*HD
  bf16* out;         // fwd: O [rows, ld_o]
  float* lse2;       // [n_seq, H, S] log2-domain logsumexp
  const bf16* dout;  // bwd: dO [rows, ld_o]
  const float* delta;  // bwd: [n_seq, H, S]  rowsum(dO * O)
  bf16* dqkv;        // bwd: [rows, ld_qkv]
  long long ld_qkv, ld_o;
  int S, H, D;
  float scale_log2;  // scale * log2(e)
  float scale;
};

// ------------------------------------------------ forward ------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const AttnArgs a) {
  __shared__ __align__(16) Tile<HD> sQ, sK, sV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM, h = blockIdx.y, seq = blockIdx.z;
  const long long row0 = (long long)seq * a.S;
  const bf16* qb = a.qkv + row0 * a.ld_qkv + h * HD;
  const bf16* kb = qb + a.D;
  const bf16* vb = qb + 2 * a.D;

  load_tile<HD>(sQ, qb + (long long)q0 * a.ld_qkv, a.ld_qkv, min(BM, a.S - q0));
  __syncthreads();
  uint32_t qf[HD / 16][4];
  load_a_frags<HD>(qf, sQ, warp, lane);

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kv0 = 0; kv0 < a.S; kv0 += BN) {
    __syncthreads();  // previous tile fully consumed
    const int valid = min(BN, a.S - kv0);
    load_tile<HD>(sK, kb + (long long)kv0 * a.ld_qkv, a.ld_qkv, valid);
    load_tile<HD>(sV, vb + (long long)kv0 * a.ld_qkv, a.ld_qkv, valid);
    __syncthreads();
    float s[8][4];
    gemm_a_tT<HD>(s, qf, sK, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (c + (j & 1)) < valid;
        s[nt][j] = ok ? s[nt][j] * a.scale_log2 : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float nm0 = fmaxf(m0, mx0), nm1 = fmaxf(m1, mx1);
    const float corr0 = exp2f(m0 - nm0), corr1 = exp2f(m1 - nm1);
    m0 = nm0; m1 = nm1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - m0); s[nt][1] = exp2f(s[nt][1] - m0);
      s[nt][2] = exp2f(s[nt][2] - m1); s[nt][3] = exp2f(s[nt][3] - m1);
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
    uint32_t p[4][4];
    pack_c_to_a(p, s);
    gemm_p_t<HD>(o, p, sV, lane);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  bf16* ob = a.out + row0 * a.ld_o + h * HD;
#pragma unroll
  for (int dn = 0; dn < HD / 8; ++dn) {
    const int c = dn * 8 + (lane & 3) * 2;
    if (r0 < a.S) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * a.ld_o + c) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
    if (r1 < a.S) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * a.ld_o + c) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
  }
  if ((lane & 3) == 0) {
    float* lp = a.lse2 + ((long long)seq * a.H + h) * a.S;
    if (r0 < a.S) lp[r0] = m0 + log2f(l0);
    if (r1 < a.S) lp[r1] = m1 + log2f(l1);
  }
}

// delta[seq,h,t] = sum_d dO * O   — one warp per (token, head)
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                                  long long ld_o, int S, int H, int HD, long long total) {
  const int lane = threadIdx.x & 31;
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total;
       w += ((long long)gridDim.x * blockDim.x) >> 5) {
    const long long row = w / H;
    const int h = (int)(w % H);
    float acc = 0.f;
    for (int c = lane * 2; c < HD; c += 64) {
      const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + row * ld_o + h * HD + c));
      const float2 y = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + row * ld_o + h * HD + c));
      acc += x.x * y.x + x.y * y.y;
    }
    acc = warp_sum(acc);
    if (lane == 0) delta[((row / S) * H + h) * S + (row % S)] = acc;
  }
}

// ------------------------------------------------ backward: dQ ------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_kernel(const AttnArgs a) {
  __shared__ __align__(16) Tile<HD> sQ, sK, sV;  // sQ is reused for dO
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM, h = blockIdx.y, seq = blockIdx.z;
  const long long row0 = (long long)seq * a.S;
  const bf16* qb = a.qkv + row0 * a.ld_qkv + h * HD;
  const bf16* kb = qb + a.D;
  const bf16* vb = qb + 2 * a.D;
  const int vq = min(BM, a.S - q0);

  uint32_t qf[HD / 16][4], dof[HD / 16][4];
  load_tile<HD>(sQ, qb + (long long)q0 * a.ld_qkv, a.ld_qkv, vq);
  __syncthreads();
  load_a_frags<HD>(qf, sQ, warp, lane);
  __syncthreads();
  load_tile<HD>(sQ, a.dout + (row0 + q0) * a.ld_o + h * HD, a.ld_o, vq);
  __syncthreads();
  load_a_frags<HD>(dof, sQ, warp, lane);

  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  const long long sb = ((long long)seq * a.H + h) * a.S;
  const float lse0 = r0 < a.S ? a.lse2[sb + r0] : 0.f, lse1 = r1 < a.S ? a.lse2[sb + r1] : 0.f;
  const float dl0 = r0 < a.S ? a.delta[sb + r0] : 0.f, dl1 = r1 < a.S ? a.delta[sb + r1] : 0.f;

  float dq[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;

  for (int kv0 = 0; kv0 < a.S; kv0 += BN) {
    __syncthreads();
    const int valid = min(BN, a.S - kv0);
    load_tile<HD>(sK, kb + (long long)kv0 * a.ld_qkv, a.ld_qkv, valid);
    load_tile<HD>(sV, vb + (long long)kv0 * a.ld_qkv, a.ld_qkv, valid);
    __syncthreads();
    float s[8][4], dp[8][4];
    gemm_a_tT<HD>(s, qf, sK, lane);
    gemm_a_tT<HD>(dp, dof, sV, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (c + (j & 1)) < valid;
        const float lse = (j < 2) ? lse0 : lse1, dl = (j < 2) ? dl0 : dl1;
        const float p = ok ? exp2f(s[nt][j] * a.scale_log2 - lse) : 0.f;
        s[nt][j] = p * (dp[nt][j] - dl) * a.scale;  // dS * scale
      }
    }
    uint32_t ds[4][4];
    pack_c_to_a(ds, s);
    gemm_p_t<HD>(dq, ds, sK, lane);
  }
  bf16* dqb = a.dqkv + row0 * a.ld_qkv + h * HD;
#pragma unroll
  for (int dn = 0; dn < HD / 8; ++dn) {
    const int c = dn * 8 + (lane & 3) * 2;
    if (r0 < a.S) *reinterpret_cast<uint32_t*>(dqb + (long long)r0 * a.ld_qkv + c) = pack_bf16x2(dq[dn][0], dq[dn][1]);
    if (r1 < a.S) *reinterpret_cast<uint32_t*>(dqb + (long long)r1 * a.ld_qkv + c) = pack_bf16x2(dq[dn][2], dq[dn][3]);
  }
}

// ------------------------------------------------ backward: dK, dV ------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkv_kernel(const AttnArgs a) {
  __shared__ __align__(16) Tile<HD> sA, sQ, sDO;  // sA stages K then V for fragment extraction
  __shared__ float s_lse[BM], s_delta[BM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * BN, h = blockIdx.y, seq = blockIdx.z;
  const long long row0 = (long long)seq * a.S;
  const bf16* qb = a.qkv + row0 * a.ld_qkv + h * HD;
  const bf16* kb = qb + a.D;
  const bf16* vb = qb + 2 * a.D;
  const int vk = min(BN, a.S - k0);

  uint32_t kf[HD / 16][4], vf[HD / 16][4];
  load_tile<HD>(sA, kb + (long long)k0 * a.ld_qkv, a.ld_qkv, vk);
  __syncthreads();
  load_a_frags<HD>(kf, sA, warp, lane);
  __syncthreads();
  load_tile<HD>(sA, vb + (long long)k0 * a.ld_qkv, a.ld_qkv, vk);
  __syncthreads();
  load_a_frags<HD>(vf, sA, warp, lane);

  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dk[i][j] = 0.f, dv[i][j] = 0.f;
  const long long sb = ((long long)seq * a.H + h) * a.S;

  for (int q0 = 0; q0 < a.S; q0 += BM) {
    __syncthreads();
    const int vq = min(BM, a.S - q0);
    load_tile<HD>(sQ, qb + (long long)q0 * a.ld_qkv, a.ld_qkv, vq);
    load_tile<HD>(sDO, a.dout + (row0 + q0) * a.ld_o + h * HD, a.ld_o, vq);
    if (threadIdx.x < BM) {
      const bool ok = threadIdx.x < vq;
      s_lse[threadIdx.x] = ok ? a.lse2[sb + q0 + threadIdx.x] : 0.f;
      s_delta[threadIdx.x] = ok ? a.delta[sb + q0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    float st[8][4], dpt[8][4];
    gemm_a_tT<HD>(st, kf, sQ, lane);     // S^T  [keys x q]
    gemm_a_tT<HD>(dpt, vf, sDO, lane);   // dP^T [keys x q]
    uint32_t pt[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = c + (j & 1);
        const float p = (q < vq) ? exp2f(st[nt][j] * a.scale_log2 - s_lse[q]) : 0.f;
        st[nt][j] = p;
        dpt[nt][j] = p * (dpt[nt][j] - s_delta[q]) * a.scale;  // dS^T * scale
      }
    }
    pack_c_to_a(pt, st);
    gemm_p_t<HD>(dv, pt, sDO, lane);  // dV += P^T dO
    pack_c_to_a(pt, dpt);
    gemm_p_t<HD>(dk, pt, sQ, lane);   // dK += dS^T Q
  }
  const int r0 = k0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  bf16* dkb = a.dqkv + row0 * a.ld_qkv + a.D + h * HD;
  bf16* dvb = a.dqkv + row0 * a.ld_qkv + 2 * a.D + h * HD;
#pragma unroll
  for (int dn = 0; dn < HD / 8; ++dn) {
    const int c = dn * 8 + (lane & 3) * 2;
    if (r0 < a.S) {
      *reinterpret_cast<uint32_t*>(dkb + (long long)r0 * a.ld_qkv + c) = pack_bf16x2(dk[dn][0], dk[dn][1]);
      *reinterpret_cast<uint32_t*>(dvb + (long long)r0 * a.ld_qkv + c) = pack_bf16x2(dv[dn][0], dv[dn][1]);
    }
    if (r1 < a.S) {
      *reinterpret_cast<uint32_t*>(dkb + (long long)r1 * a.ld_qkv + c) = pack_bf16x2(dk[dn][2], dk[dn][3]);
      *reinterpret_cast<uint32_t*>(dvb + (long long)r1 * a.ld_qkv + c) = pack_bf16x2(dv[dn][2], dv[dn][3]);
    }
  }
}

int check_common(const void* qkv, long long ld_qkv, long long ld_o, int n_seq, int S, int H, int HD, const char* who) {
  AVS_REQUIRE(qkv != nullptr, "%s: null pointer", who);
  AVS_REQUIRE(HD == 32 || HD == 64, "%s: head_dim must be 32 or 64 (got %d)", who, HD);
  AVS_REQUIRE(S > 0 && H > 0 && n_seq >= 0 && n_seq <= 65535 && H <= 65535, "%s: bad shape", who);
  AVS_REQUIRE(ld_qkv % 8 == 0 && ld_o % 8 == 0 && ((uintptr_t)qkv & 15) == 0, "%s: 16-byte alignment", who);
  return 0;
}

}  // namespace

extern "C" int avs_attention_fwd(const void* qkv, long long ld_qkv, void* out, long long ld_o, float* lse2, int n_seq,
                                 int S, int H, int head_dim, void* stream) {
  if (check_common(qkv, ld_qkv, ld_o, n_seq, S, H, head_dim, "avs_attention_fwd")) return -1;
  AVS_REQUIRE(out && lse2, "avs_attention_fwd: null pointer");
  if (n_seq == 0) return 0;
  AttnArgs a = {};
  a.qkv = (const bf16*)qkv; a.out = (bf16*)out; a.lse2 = lse2;
  a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.S = S; a.H = H; a.D = H * head_dim;
  a.scale = rsqrtf((float)head_dim);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  dim3 grid(ceil_div(S, BM), H, n_seq);
  if (head_dim == 64) attn_fwd_kernel<64><<<grid, ATT_THREADS, 0, (cudaStream_t)stream>>>(a);
  else attn_fwd_kernel<32><<<grid, ATT_THREADS, 0, (cudaStream_t)stream>>>(a);
  return avs_check_launch("attn_fwd_kernel");
}

// delta: scratch [n_seq, H, S] fp32
extern "C" int avs_attention_bwd(const void* qkv, long long ld_qkv, const void* out, const void* dout, long long ld_o,
                                 const float* lse2, float* delta, void* dqkv, int n_seq, int S, int H, int head_dim,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (check_common(qkv, ld_qkv, ld_o, n_seq, S, H, head_dim, "avs_attention_bwd")) return -1;
  AVS_REQUIRE(out && dout && lse2 && delta && dqkv, "avs_attention_bwd: null pointer");
  if (n_seq == 0) return 0;
  AttnArgs a = {};
  a.qkv = (const bf16*)qkv; a.dout = (const bf16*)dout; a.lse2 = const_cast<float*>(lse2); a.delta = delta;
  a.dqkv = (bf16*)dqkv;
  a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.S = S; a.H = H; a.D = H * head_dim;
  a.scale = rsqrtf((float)head_dim);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  const long long total = (long long)n_seq * S * H;
  const int dblocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll(total * 32, 256));
  attn_delta_kernel<<<dblocks, 256, 0, stream>>>((const bf16*)out, (const bf16*)dout, delta, ld_o, S, H, head_dim,
                                                 total);
  int rc = avs_check_launch("attn_delta_kernel");
  if (rc) return rc;
  dim3 grid(ceil_div(S, BM), H, n_seq);
  if (head_dim == 64) {
    attn_bwd_dkv_kernel<64><<<grid, ATT_THREADS, 0, stream>>>(a);
    if ((rc = avs_check_launch("attn_bwd_dkv_kernel"))) return rc;
    attn_bwd_dq_kernel<64><<<grid, ATT_THREADS, 0, stream>>>(a);
  } else {
    attn_bwd_dkv_kernel<32><<<grid, ATT_THREADS, 0, stream>>>(a);
    if ((rc = avs_check_launch("attn_bwd_dkv_kernel"))) return rc;
    attn_bwd_dq_kernel<32><<<grid, ATT_THREADS, 0, stream>>>(a);
  }
  return avs_check_launch("attn_bwd_dq_kernel");
}
