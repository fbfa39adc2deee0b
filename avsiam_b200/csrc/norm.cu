// LayerNorm forward / backward (one warp per row, fp32 statistics, bf16 activations, fp32 affine parameters),
// and the per-sequence token mean that feeds InfoNCE.  HBM-bound: each kernel touches every element once.
//
// Replaces nn.LayerNorm inside Block (cav_mae_base.py:151-152,169-170,190-191, eps 1e-5), decoder_norm (:631),
// the final vit norm / norm_a (:492-495, :563-566, eps 1e-6) and `.mean(dim=1)` (:563-566, :729).
//
// "row map": the y / dy side may live in a per-sample concatenated layout (torch.cat((ca,cv),dim=1) at :503):
//   y_row(r) = (r / S) * y_seq_stride + y_off + (r % S)     (S == 0 => identity)
#include "../../include/avsiam_b200.h"
#include "common.cuh"
#include <stdlib.h>

// Two affine sets in one launch: rows [0, split) use set 0, rows [split, M) set 1 (audio | video tokens of the shared
// encoder, Block.forward's modality switch at cav_mae_base.py:151-152,169-170,190-191). CTAs [0, cta_split) work on the
// first range, the rest on the second, so every thread still keeps ITS affine chunk in registers for the whole kernel.
struct LnSeg {
  int lo, hi;        // row range of this CTA
  int bi, nb;        // index of this CTA among the nb CTAs of the range
  int which;         // 0 / 1
};
__device__ __forceinline__ LnSeg ln_segment(int M, int split, int cta_split) {
  LnSeg s;
  s.which = ((int)blockIdx.x >= cta_split) ? 1 : 0;
  s.lo = s.which ? split : 0;
  s.hi = s.which ? M : split;
  s.bi = s.which ? (int)blockIdx.x - cta_split : (int)blockIdx.x;
  s.nb = s.which ? (int)gridDim.x - cta_split : cta_split;
  return s;
}
static int ln_cta_split(int blocks, int M, int split) {
  if (split >= M) return blocks;
  if (split <= 0) return 0;
  if (blocks < 2) return -1;   // caller must provide >= 2 CTAs for two non-empty ranges
  int c = (int)(((long long)blocks * split + M / 2) / M);
  return c < 1 ? 1 : (c > blocks - 1 ? blocks - 1 : c);
}

// rows are counted in int (M < 2^31): the division runs in 32 bits (a 64-bit division costs ~80 instructions, and
// these kernels are issue-bound)
__device__ __forceinline__ long long map_row(long long r, int S, int stride, int off) {
  if (!(S > 0 && stride > 0)) return r;
  const int ri = (int)r, q = ri / S;
  return (long long)q * stride + off + (ri - q * S);
}

__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

__device__ __forceinline__ void unpack8u(const uint4& u, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}

// One warp per row, software-pipelined: the 16-byte loads of the warp's NEXT row are in flight while the current
// row is reduced and written (one row per warp in flight left the kernel latency-bound at ~3.5 TB/s).
template <int NCH>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const bf16* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            bf16* __restrict__ y, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int M_all, int D, int S,
                                                            int x_stride, int x_off, int y_stride, int y_off,
                                                            const float* __restrict__ gamma1,
                                                            const float* __restrict__ beta1, int split, int cta_split) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = D / 8;
  const LnSeg sg = ln_segment(M_all, split, cta_split);
  if (sg.which) { gamma = gamma1; beta = beta1; }
  const int M = sg.hi;
  const long long stride = (long long)sg.nb * 8;
  long long r = (long long)sg.lo + (long long)sg.bi * 8 + warp;
  uint4 cur[NCH], nxt[NCH];
  auto issue = [&](long long row, uint4 (&dst)[NCH]) {
    const long long xr = map_row(row, S, x_stride, x_off);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int ci = lane + 32 * k;
      dst[k] = (ci < nchunks) ? __ldg(reinterpret_cast<const uint4*>(x + xr * D + ci * 8)) : make_uint4(0, 0, 0, 0);
    }
  };
  if (r < M) issue(r, cur);
  for (; r < M; r += stride) {
    const bool more = (r + stride) < M;
    if (more) issue(r + stride, nxt);
    float v[NCH][8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      unpack8u(cur[k], v[k]);   // chunks beyond D were loaded as zeros
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[k][j];
    }
    const float mean = warp_sum(sum) / D;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunks) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[k][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / D + eps);
    if (lane == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
    const long long yr = map_row(r, S, y_stride, y_off);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunks) {
        float o[8];
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ci * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + ci * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + ci * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + ci * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * g[j] + b[j];
        store8(y + yr * D + ci * 8, o);
      }
    }
    if (more) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) cur[k] = nxt[k];
    }
  }
}

// Backward.  Thread mapping chosen for memory-level parallelism at low register cost: every thread owns ONE
// 16-byte chunk (8 columns) of a row, a row is covered by TPR = D/8 consecutive threads ("row group"), a CTA holds
// RG row groups and each group keeps LNB_R rows in flight.  The per-column accumulators (dgamma, dbeta, and the
// column sum of the produced dx = bias gradient of the Linear that feeds this residual stream) are therefore
// 3 x 8 registers per thread instead of 3 x D/32, and x / dy / resid stay packed in bf16 until used.
#ifndef AVS_LNB_R
#define AVS_LNB_R 2
#endif
#ifndef AVS_LNB_STAGES
#define AVS_LNB_STAGES 3
#endif
constexpr int LNB_R = AVS_LNB_R;
constexpr int LNB_STAGES = AVS_LNB_STAGES;   // iterations of x / dy / resid chunks in flight per thread (cp.async ring)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}

template <int TPR, int RG>
__global__ void __launch_bounds__(TPR * RG, 2) layernorm_bwd_kernel(
    const bf16* __restrict__ dy, const float* __restrict__ dpool, float pool_scale, const bf16* __restrict__ x,
    const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
    const bf16* __restrict__ resid, bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
    float* __restrict__ dbias, int M_all, int S, int x_stride, int x_off, int y_stride, int y_off,
    const float* __restrict__ gamma1, float* __restrict__ dgamma1, float* __restrict__ dbeta1, int split,
    int cta_split) {
  constexpr int NT = TPR * RG, NW = (NT + 31) / 32, D = TPR * 8;
  const LnSeg sg = ln_segment(M_all, split, cta_split);
  if (sg.which) { gamma = gamma1; dgamma = dgamma1; dbeta = dbeta1; }
  const int M = sg.hi;
  constexpr int WPG = (TPR >= 32) ? TPR / 32 : 1;   // warps per row group
  __shared__ float red[2][LNB_R][2][NW];
  __shared__ float sacc[3][D];
  const int t = threadIdx.x, rg = t / TPR, ci = t % TPR, warp = t >> 5, lane = t & 31;
  for (int i = t; i < 3 * D; i += NT) (&sacc[0][0])[i] = 0.f;

  float g[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + ci * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(gamma + ci * 8 + 4);
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
  }
  float ag[8], ab[8], ad[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ag[j] = ab[j] = ad[j] = 0.f;
  const float inv_d = 1.0f / D;

  // The x / dy / resid chunks travel through a per-thread cp.async ring in shared memory, LNB_STAGES iterations deep:
  // every thread copies and later reads only ITS OWN 16-byte slots, so the pipeline needs no barrier — each thread waits
  // on its own cp.async groups — and the loads of the next iterations stay in flight during the reductions of this
  // one (with plain loads the kernel alternated between a burst of loads and a compute phase: ~3 TB/s).
  extern __shared__ uint4 ln_ring[];    // [LNB_STAGES][LNB_R][3][NT]
  const long long step = (long long)sg.nb * (RG * LNB_R);
  auto issue = [&](long long base, int stage) {
#pragma unroll
    for (int k = 0; k < LNB_R; ++k) {
      const long long r = base + k * RG + rg;
      if (r < M) {
        const long long xr = map_row(r, S, x_stride, x_off);
        uint4* slot = ln_ring + ((size_t)(stage * LNB_R + k) * 3) * NT + t;
        cp_async16(slot, x + xr * D + ci * 8);
        if (dy != nullptr) cp_async16(slot + NT, dy + map_row(r, S, y_stride, y_off) * D + ci * 8);
        if (resid != nullptr) cp_async16(slot + 2 * NT, resid + xr * D + ci * 8);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  const long long base0 = (long long)sg.lo + (long long)sg.bi * (RG * LNB_R);
#pragma unroll
  for (int st = 0; st < LNB_STAGES - 1; ++st) issue(base0 + st * step, st);   // (groups past M are empty)

  int it = 0;
  for (long long base = base0; base < M; base += step, ++it) {
    issue(base + (LNB_STAGES - 1) * step, (it + LNB_STAGES - 1) % LNB_STAGES);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(LNB_STAGES - 1) : "memory");
    const int stage = it % LNB_STAGES;
    uint4 xv[LNB_R], dv[LNB_R], rv[LNB_R];
    float mean[LNB_R], rstd[LNB_R], s1[LNB_R], s2[LNB_R];
    long long xrow[LNB_R], seq[LNB_R];
    bool ok[LNB_R];
#pragma unroll
    for (int k = 0; k < LNB_R; ++k) {
      const long long r = base + k * RG + rg;
      ok[k] = r < M;
      xv[k] = dv[k] = rv[k] = make_uint4(0, 0, 0, 0);
      mean[k] = 0.f; rstd[k] = 0.f; xrow[k] = 0; seq[k] = 0;
      if (ok[k]) {
        xrow[k] = map_row(r, S, x_stride, x_off);
        seq[k] = S > 0 ? (int)r / S : 0;
        const uint4* slot = ln_ring + ((size_t)(stage * LNB_R + k) * 3) * NT + t;
        xv[k] = slot[0];
        if (dy != nullptr) dv[k] = slot[NT];
        if (resid != nullptr) rv[k] = slot[2 * NT];
        mean[k] = mean_in[r];
        rstd[k] = rstd_in[r];
      }
    }
#pragma unroll
    for (int k = 0; k < LNB_R; ++k) {
      float xf[8], d[8];
      unpack8(xv[k], xf);
      unpack8(dv[k], d);
      if (dpool != nullptr && ok[k]) {
        const float4 p0 = *reinterpret_cast<const float4*>(dpool + seq[k] * D + ci * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(dpool + seq[k] * D + ci * 8 + 4);
        d[0] += p0.x * pool_scale; d[1] += p0.y * pool_scale; d[2] += p0.z * pool_scale; d[3] += p0.w * pool_scale;
        d[4] += p1.x * pool_scale; d[5] += p1.y * pool_scale; d[6] += p1.z * pool_scale; d[7] += p1.w * pool_scale;
      }
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h = (xf[j] - mean[k]) * rstd[k];
        const float gd = g[j] * d[j];
        ag[j] += d[j] * h;
        ab[j] += d[j];
        a1 += gd;
        a2 += gd * h;
      }
      s1[k] = a1; s2[k] = a2;
    }
    // row reductions over the TPR threads of the group
    if (TPR >= 32) {
#pragma unroll
      for (int k = 0; k < LNB_R; ++k) {
        s1[k] = warp_sum(s1[k]);
        s2[k] = warp_sum(s2[k]);
      }
      if (WPG > 1) {
        const int b = it & 1;
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < LNB_R; ++k) { red[b][k][0][warp] = s1[k]; red[b][k][1][warp] = s2[k]; }
        }
        // only the WPG warps of this row group exchange partial sums: a named barrier per group lets the RG groups of
        // the CTA drift apart instead of all waiting for each other in every iteration
        asm volatile("bar.sync %0, %1;\n" ::"r"(1 + rg), "r"(TPR) : "memory");
#pragma unroll
        for (int k = 0; k < LNB_R; ++k) {
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int w = 0; w < WPG; ++w) { t1 += red[b][k][0][rg * WPG + w]; t2 += red[b][k][1][rg * WPG + w]; }
          s1[k] = t1; s2[k] = t2;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < LNB_R; ++k) {
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) {
          s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
          s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < LNB_R; ++k) {
      if (!ok[k]) continue;
      const float c1 = s1[k] * inv_d, c2 = s2[k] * inv_d;
      float xf[8], d[8], rr[8], o[8];
      unpack8(xv[k], xf);
      unpack8(dv[k], d);
      unpack8(rv[k], rr);
      if (dpool != nullptr) {
        const float4 p0 = *reinterpret_cast<const float4*>(dpool + seq[k] * D + ci * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(dpool + seq[k] * D + ci * 8 + 4);
        d[0] += p0.x * pool_scale; d[1] += p0.y * pool_scale; d[2] += p0.z * pool_scale; d[3] += p0.w * pool_scale;
        d[4] += p1.x * pool_scale; d[5] += p1.y * pool_scale; d[6] += p1.z * pool_scale; d[7] += p1.w * pool_scale;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h = (xf[j] - mean[k]) * rstd[k];
        o[j] = rstd[k] * (g[j] * d[j] - c1 - h * c2) + rr[j];
        ad[j] += o[j];
      }
      store8(dx + xrow[k] * D + ci * 8, o);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sacc[0][ci * 8 + j], ag[j]);
    atomicAdd(&sacc[1][ci * 8 + j], ab[j]);
    if (dbias != nullptr) atomicAdd(&sacc[2][ci * 8 + j], ad[j]);
  }
  __syncthreads();
  for (int i = t; i < D; i += NT) {
    atomicAdd(dgamma + i, sacc[0][i]);
    atomicAdd(dbeta + i, sacc[1][i]);
    if (dbias != nullptr) atomicAdd(dbias + i, sacc[2][i]);
  }
}

// Forward for the transformer widths (512 / 768 / 1024), with the backward's thread mapping: every thread owns ONE
// 16-byte chunk of a row (gamma / beta for its 8 columns stay in registers for the whole kernel), a row is covered by
// TPR consecutive threads, LNF_R rows per row group and iteration, x travels through a per-thread cp.async ring
// LNF_STAGES iterations deep, and the two reductions (mean, centred second moment) synchronise only the WPG warps of a
// row group through a named barrier.
#ifndef AVS_LNF_R
#define AVS_LNF_R 4   // rows in flight per row group: 2 -> 4 measured +5 % (decoder) ... +10 % (encoder shapes), profiles/r02_ln_ab.log
#endif
#ifndef AVS_LNF_STAGES
#define AVS_LNF_STAGES 4
#endif
constexpr int LNF_R = AVS_LNF_R, LNF_STAGES = AVS_LNF_STAGES;
template <int TPR, int RG>
__global__ void __launch_bounds__(TPR * RG, 2) layernorm_fwd_rg_kernel(
    const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M_all, int S, int x_stride,
    int x_off, int y_stride, int y_off, const float* __restrict__ gamma1, const float* __restrict__ beta1, int split,
    int cta_split) {
  constexpr int NT = TPR * RG, NW = NT / 32, D = TPR * 8, WPG = TPR / 32;
  const LnSeg sg = ln_segment(M_all, split, cta_split);
  if (sg.which) { gamma = gamma1; beta = beta1; }
  const int M = sg.hi;
  static_assert(TPR % 32 == 0 && WPG >= 2, "row groups must be whole warps");
  extern __shared__ uint4 lnf_ring[];   // [LNF_STAGES][LNF_R][NT]
  __shared__ float red[2][LNF_R][NW];
  const int t = threadIdx.x, rg = t / TPR, ci = t % TPR, warp = t >> 5, lane = t & 31;
  float g[8], b[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + ci * 8), g1 = *reinterpret_cast<const float4*>(gamma + ci * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + ci * 8), b1 = *reinterpret_cast<const float4*>(beta + ci * 8 + 4);
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  }
  const long long step = (long long)sg.nb * (RG * LNF_R);
  auto issue = [&](long long base, int stage) {
#pragma unroll
    for (int k = 0; k < LNF_R; ++k) {
      const long long r = base + k * RG + rg;
      if (r < M) cp_async16(lnf_ring + (size_t)(stage * LNF_R + k) * NT + t, x + map_row(r, S, x_stride, x_off) * D + ci * 8);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  const long long base0 = (long long)sg.lo + (long long)sg.bi * (RG * LNF_R);
#pragma unroll
  for (int st = 0; st < LNF_STAGES - 1; ++st) issue(base0 + st * step, st);
  const float inv_d = 1.0f / D;
  int it = 0;
  for (long long base = base0; base < M; base += step, ++it) {
    issue(base + (LNF_STAGES - 1) * step, (it + LNF_STAGES - 1) % LNF_STAGES);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(LNF_STAGES - 1) : "memory");
    const int stage = it % LNF_STAGES;
    float v[LNF_R][8], s[LNF_R];
#pragma unroll
    for (int k = 0; k < LNF_R; ++k) {
      const long long r = base + k * RG + rg;
      const uint4 u = (r < M) ? lnf_ring[(size_t)(stage * LNF_R + k) * NT + t] : make_uint4(0, 0, 0, 0);
      unpack8(u, v[k]);
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) a += v[k][j];
      s[k] = warp_sum(a);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < LNF_R; ++k) red[0][k][warp] = s[k];
    }
    asm volatile("bar.sync %0, %1;\n" ::"r"(1 + rg), "r"(TPR) : "memory");
    float mean[LNF_R], q[LNF_R];
#pragma unroll
    for (int k = 0; k < LNF_R; ++k) {
      float tsum = 0.f;
#pragma unroll
      for (int w = 0; w < WPG; ++w) tsum += red[0][k][rg * WPG + w];
      mean[k] = tsum * inv_d;
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[k][j] - mean[k];
        a = fmaf(d, d, a);
      }
      q[k] = warp_sum(a);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < LNF_R; ++k) red[1][k][warp] = q[k];
    }
    asm volatile("bar.sync %0, %1;\n" ::"r"(1 + rg), "r"(TPR) : "memory");
#pragma unroll
    for (int k = 0; k < LNF_R; ++k) {
      const long long r = base + k * RG + rg;
      if (r >= M) continue;
      float tq = 0.f;
#pragma unroll
      for (int w = 0; w < WPG; ++w) tq += red[1][k][rg * WPG + w];
      const float rstd = rsqrtf(tq * inv_d + eps);
      if (ci == 0) {
        mean_out[r] = mean[k];
        rstd_out[r] = rstd;
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean[k]) * rstd * g[j] + b[j];
      store8(y + map_row(r, S, y_stride, y_off) * D + ci * 8, o);
    }
  }
}

template <int TPR, int RG>
static void launch_ln_fwd_rg(const void* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                             float* rstd, int M, int S, int x_stride, int x_off, int y_stride, int y_off,
                             const float* gamma1, const float* beta1, int split, cudaStream_t stream) {
  const int ring = LNF_STAGES * LNF_R * TPR * RG * 16;
  static bool ring_set = false;
  if (!ring_set) {
    cudaFuncSetAttribute(layernorm_fwd_rg_kernel<TPR, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring);
    ring_set = true;
  }
  const int blocks = max(split < M && split > 0 ? 2 : 1, min(avs_num_sms() * 2, ceil_div(M, RG * LNF_R)));
  layernorm_fwd_rg_kernel<TPR, RG><<<blocks, TPR * RG, ring, stream>>>((const bf16*)x, gamma, beta, eps, (bf16*)y, mean,
                                                                      rstd, M, S, x_stride, x_off, y_stride, y_off,
                                                                      gamma1, beta1, split, ln_cta_split(blocks, M, split));
}

static int ln_nch(int D) { return (D + 255) / 256; }

static int ln_fwd_impl(const void* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                       float* rstd, int M, int D, int seq_len, int x_seq_stride, int x_off, int y_seq_stride, int y_off,
                       const float* gamma1, const float* beta1, int split, cudaStream_t stream) {
  AVS_REQUIRE(x && gamma && beta && y && mean && rstd, "avs_layernorm_fwd: null pointer");
  AVS_REQUIRE(D % 8 == 0 && D <= 2048, "avs_layernorm_fwd: D must be a multiple of 8 and <= 2048 (got %d)", D);
  AVS_REQUIRE(((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0, "avs_layernorm_fwd: gamma/beta 16-byte alignment");
  if (M == 0) return 0;
  static const bool rg_path = !(getenv("AVS_LN_FWD_RG") && atoi(getenv("AVS_LN_FWD_RG")) == 0);
  if (rg_path && (D == 512 || D == 768 || D == 1024 || D == 1280)) {
    if (D == 512) launch_ln_fwd_rg<64, 4>(x, gamma, beta, eps, y, mean, rstd, M, seq_len, x_seq_stride, x_off, y_seq_stride, y_off, gamma1, beta1, split, stream);
    else if (D == 1280) launch_ln_fwd_rg<160, 2>(x, gamma, beta, eps, y, mean, rstd, M, seq_len, x_seq_stride, x_off, y_seq_stride, y_off, gamma1, beta1, split, stream);   // ViT-H
    else if (D == 768) launch_ln_fwd_rg<96, 4>(x, gamma, beta, eps, y, mean, rstd, M, seq_len, x_seq_stride, x_off, y_seq_stride, y_off, gamma1, beta1, split, stream);
    else launch_ln_fwd_rg<128, 2>(x, gamma, beta, eps, y, mean, rstd, M, seq_len, x_seq_stride, x_off, y_seq_stride, y_off, gamma1, beta1, split, stream);
    return avs_check_launch("layernorm_fwd_rg_kernel");
  }
  const int blocks = max(split < M && split > 0 ? 2 : 1, min(avs_num_sms() * 8, ceil_div(M, 8)));
  const int cta_split = ln_cta_split(blocks, M, split);
#define LN_FWD(N)                                                                                              \
  layernorm_fwd_kernel<N><<<blocks, 256, 0, stream>>>((const bf16*)x, gamma, beta, eps, (bf16*)y, mean, rstd, M, D, \
                                                      seq_len, x_seq_stride, x_off, y_seq_stride, y_off, gamma1, beta1, \
                                                      split, cta_split)
  switch (ln_nch(D)) {
    case 1: LN_FWD(1); break;
    case 2: LN_FWD(2); break;
    case 3: LN_FWD(3); break;
    case 4: LN_FWD(4); break;
    case 5: LN_FWD(5); break;
    default: LN_FWD(8); break;
  }
#undef LN_FWD
  return avs_check_launch("layernorm_fwd_kernel");
}

extern "C" int avs_layernorm_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y,
                                 float* mean, float* rstd, int M, int D, int seq_len, int x_seq_stride, int x_off,
                                 int y_seq_stride, int y_off, void* stream_) {
  return ln_fwd_impl(x, gamma, beta, eps, y, mean, rstd, M, D, seq_len, x_seq_stride, x_off, y_seq_stride, y_off,
                     nullptr, nullptr, M, (cudaStream_t)stream_);
}

extern "C" int avs_layernorm_fwd2(const void* x, const float* gamma0, const float* beta0, int split_row,
                                  const float* gamma1, const float* beta1, float eps, void* y, float* mean,
                                  float* rstd, int M, int D, void* stream_) {
  AVS_REQUIRE(split_row >= 0 && split_row <= M, "avs_layernorm_fwd2: split_row out of range");
  AVS_REQUIRE(split_row == M || (gamma1 && beta1 && ((uintptr_t)gamma1 & 15) == 0 && ((uintptr_t)beta1 & 15) == 0),
              "avs_layernorm_fwd2: second affine set missing / misaligned");
  return ln_fwd_impl(x, gamma0, beta0, eps, y, mean, rstd, M, D, 0, 0, 0, 0, 0, gamma1, beta1, split_row,
                     (cudaStream_t)stream_);
}

template <int TPR, int RG>
static void launch_ln_bwd(const void* dy, const float* dpool, float pool_scale, const void* x, const float* mean,
                          const float* rstd, const float* gamma, const void* resid, void* dx, float* dgamma,
                          float* dbeta, float* dbias, int M, int seq_len, int x_seq_stride, int x_off,
                          int y_seq_stride, int y_off, const float* gamma1, float* dgamma1, float* dbeta1, int split,
                          cudaStream_t stream) {
  const int rows_per_cta = RG * LNB_R;
  const int ctas_per_sm = 2;
  const int blocks = max(split < M && split > 0 ? 2 : 1, min(avs_num_sms() * ctas_per_sm, ceil_div(M, rows_per_cta)));
  const int ring = LNB_STAGES * LNB_R * 3 * TPR * RG * 16;
  static bool ring_set = false;   // per instantiation (static + dynamic shared memory can exceed the 48 KB default)
  if (!ring_set) {
    cudaFuncSetAttribute(layernorm_bwd_kernel<TPR, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring);
    ring_set = true;
  }
  layernorm_bwd_kernel<TPR, RG><<<blocks, TPR * RG, ring, stream>>>(
      (const bf16*)dy, dpool, pool_scale, (const bf16*)x, mean, rstd, gamma, (const bf16*)resid, (bf16*)dx, dgamma,
      dbeta, dbias, M, seq_len, x_seq_stride, x_off, y_seq_stride, y_off, gamma1, dgamma1, dbeta1, split,
      ln_cta_split(blocks, M, split));
}

static int ln_bwd_impl(const void* dy, const float* dpool, float pool_scale, const void* x, const float* mean,
                       const float* rstd, const float* gamma, const void* resid, void* dx, float* dgamma, float* dbeta,
                       float* dbias, int M, int D, int seq_len, int x_seq_stride, int x_off, int y_seq_stride, int y_off,
                       const float* gamma1, float* dgamma1, float* dbeta1, int split, cudaStream_t stream) {
  AVS_REQUIRE(x && mean && rstd && gamma && dx && dgamma && dbeta, "avs_layernorm_bwd: null pointer");
  AVS_REQUIRE(dy != nullptr || dpool != nullptr, "avs_layernorm_bwd: need dy and/or dpool");
  AVS_REQUIRE(dpool == nullptr || seq_len > 0, "avs_layernorm_bwd: dpool needs seq_len");
  AVS_REQUIRE(((uintptr_t)gamma & 15) == 0 && (dpool == nullptr || ((uintptr_t)dpool & 15) == 0),
              "avs_layernorm_bwd: gamma / dpool 16-byte alignment");
  if (M == 0) return 0;
#define LN_BWD(TPR, RG)                                                                                          \
  launch_ln_bwd<TPR, RG>(dy, dpool, pool_scale, x, mean, rstd, gamma, resid, dx, dgamma, dbeta, dbias, M, seq_len, \
                         x_seq_stride, x_off, y_seq_stride, y_off, gamma1, dgamma1, dbeta1, split, stream)
  switch (D) {
    case 64: LN_BWD(8, 32); break;
    case 128: LN_BWD(16, 16); break;
    case 256: LN_BWD(32, 8); break;
    case 512: LN_BWD(64, 4); break;
    case 768: LN_BWD(96, 2); break;
    case 1024: LN_BWD(128, 2); break;
    case 1280: LN_BWD(160, 1); break;
    default:
      avs_set_error("avs_layernorm_bwd: unsupported width D=%d (instantiated: 64,128,256,512,768,1024,1280)", D);
      return -1;
  }
#undef LN_BWD
  return avs_check_launch("layernorm_bwd_kernel");
}

extern "C" int avs_layernorm_bwd(const void* dy, const float* dpool, float pool_scale, const void* x,
                                 const float* mean, const float* rstd, const float* gamma, const void* resid,
                                 void* dx, float* dgamma, float* dbeta, float* dbias, int M, int D, int seq_len,
                                 int x_seq_stride, int x_off, int y_seq_stride, int y_off, void* stream_) {
  return ln_bwd_impl(dy, dpool, pool_scale, x, mean, rstd, gamma, resid, dx, dgamma, dbeta, dbias, M, D, seq_len,
                     x_seq_stride, x_off, y_seq_stride, y_off, nullptr, nullptr, nullptr, M, (cudaStream_t)stream_);
}

extern "C" int avs_layernorm_bwd2(const void* dy, const void* x, const float* mean, const float* rstd,
                                  const float* gamma0, float* dgamma0, float* dbeta0, int split_row,
                                  const float* gamma1, float* dgamma1, float* dbeta1, const void* resid, void* dx,
                                  float* dbias, int M, int D, void* stream_) {
  AVS_REQUIRE(split_row >= 0 && split_row <= M, "avs_layernorm_bwd2: split_row out of range");
  AVS_REQUIRE(split_row == M || (gamma1 && dgamma1 && dbeta1 && ((uintptr_t)gamma1 & 15) == 0),
              "avs_layernorm_bwd2: second affine set missing / misaligned");
  AVS_REQUIRE(dy != nullptr, "avs_layernorm_bwd2: dy required");
  return ln_bwd_impl(dy, nullptr, 0.f, x, mean, rstd, gamma0, resid, dx, dgamma0, dbeta0, dbias, M, D, 0, 0, 0, 0, 0,
                     gamma1, dgamma1, dbeta1, split_row, (cudaStream_t)stream_);
}

// out[s, :] = (1/S) * sum_t y[y_row(s*S + t), :]   (fp32 output)
__global__ void seq_mean_fwd_kernel(const bf16* __restrict__ y, float* __restrict__ out, int S, int D, int y_stride,
                                    int y_off) {
  const int s = blockIdx.x;
  for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < S; ++t) {
      const long long yr = map_row((long long)s * S + t, S, y_stride, y_off);
      float v[8];
      load8(y + yr * D + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    const float inv = 1.0f / S;
#pragma unroll
    for (int j = 0; j < 8; ++j) out[(size_t)s * D + c + j] = acc[j] * inv;
  }
}

extern "C" int avs_seq_mean_fwd(const void* y, float* out, int n_seq, int seq_len, int D, int y_seq_stride,
                                int y_off, void* stream) {
  AVS_REQUIRE(y && out, "avs_seq_mean_fwd: null pointer");
  AVS_REQUIRE(D % 8 == 0 && seq_len > 0, "avs_seq_mean_fwd: bad shape");
  if (n_seq == 0) return 0;
  const int threads = min(256, ((D / 8 + 31) / 32) * 32);
  seq_mean_fwd_kernel<<<n_seq, threads, 0, (cudaStream_t)stream>>>((const bf16*)y, out, seq_len, D, y_seq_stride,
                                                                  y_off);
  return avs_check_launch("seq_mean_fwd_kernel");
}
