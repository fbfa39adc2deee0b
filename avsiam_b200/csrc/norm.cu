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

__device__ __forceinline__ long long map_row(long long r, int S, int stride, int off) {
  return (S > 0 && stride > 0) ? (r / S) * (long long)stride + off + (r % S) : r;
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

template <int NCH>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const bf16* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            bf16* __restrict__ y, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int M, int D, int S,
                                                            int x_stride, int x_off, int y_stride, int y_off) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = D / 8;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < M; r += (long long)gridDim.x * 8) {
    float v[NCH][8];
    float sum = 0.f;
    const long long xr = map_row(r, S, x_stride, x_off);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunks) {
        load8(x + xr * D + ci * 8, v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[k][j];
      }
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
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + ci * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gamma + ci * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + ci * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(beta + ci * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * g[j] + b[j];
        store8(y + yr * D + ci * 8, o);
      }
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(
    const bf16* __restrict__ dy, const float* __restrict__ dpool, float pool_scale, const bf16* __restrict__ x,
    const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
    const bf16* __restrict__ resid, bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
    int M, int D, int S, int x_stride, int x_off, int y_stride, int y_off) {
  extern __shared__ float sred[];  // [2*D]: dgamma partials, dbeta partials
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = D / 8;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float ag[NCH][8], ab[NCH][8];
#pragma unroll
  for (int k = 0; k < NCH; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[k][j] = 0.f, ab[k][j] = 0.f;

  for (long long r = (long long)blockIdx.x * 8 + warp; r < M; r += (long long)gridDim.x * 8) {
    const float mean = mean_in[r], rstd = rstd_in[r];
    const long long yr = map_row(r, S, y_stride, y_off);
    const long long xr = map_row(r, S, x_stride, x_off);
    const long long seq = S > 0 ? r / S : 0;
    float xh[NCH][8], gd[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunks) {
        float xv[8], d[8];
        load8(x + xr * D + ci * 8, xv);
        if (dy != nullptr) {
          load8(dy + yr * D + ci * 8, d);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = 0.f;
        }
        if (dpool != nullptr) {
          const float4 p0 = *reinterpret_cast<const float4*>(dpool + seq * D + ci * 8);
          const float4 p1 = *reinterpret_cast<const float4*>(dpool + seq * D + ci * 8 + 4);
          d[0] += p0.x * pool_scale; d[1] += p0.y * pool_scale; d[2] += p0.z * pool_scale; d[3] += p0.w * pool_scale;
          d[4] += p1.x * pool_scale; d[5] += p1.y * pool_scale; d[6] += p1.z * pool_scale; d[7] += p1.w * pool_scale;
        }
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + ci * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gamma + ci * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = (xv[j] - mean) * rstd;
          xh[k][j] = h;
          ag[k][j] += d[j] * h;
          ab[k][j] += d[j];
          const float gdy = g[j] * d[j];
          gd[k][j] = gdy;
          s1 += gdy;
          s2 += gdy * h;
        }
      }
    }
    const float c1 = warp_sum(s1) / D, c2 = warp_sum(s2) / D;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunks) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (gd[k][j] - c1 - xh[k][j] * c2);
        if (resid != nullptr) {
          float rr[8];
          load8(resid + xr * D + ci * 8, rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rr[j];
        }
        store8(dx + xr * D + ci * 8, o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int ci = lane + 32 * k;
    if (ci < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sred[ci * 8 + j], ag[k][j]);
        atomicAdd(&sred[D + ci * 8 + j], ab[k][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, sred[i]);
    atomicAdd(dbeta + i, sred[D + i]);
  }
}

static int ln_nch(int D) { return (D + 255) / 256; }

extern "C" int avs_layernorm_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y,
                                 float* mean, float* rstd, int M, int D, int seq_len, int x_seq_stride, int x_off,
                                 int y_seq_stride, int y_off, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AVS_REQUIRE(x && gamma && beta && y && mean && rstd, "avs_layernorm_fwd: null pointer");
  AVS_REQUIRE(D % 8 == 0 && D <= 2048, "avs_layernorm_fwd: D must be a multiple of 8 and <= 2048 (got %d)", D);
  AVS_REQUIRE(((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0, "avs_layernorm_fwd: gamma/beta 16-byte alignment");
  if (M == 0) return 0;
  const int blocks = min(avs_num_sms() * 8, ceil_div(M, 8));
#define LN_FWD(N)                                                                                              \
  layernorm_fwd_kernel<N><<<blocks, 256, 0, stream>>>((const bf16*)x, gamma, beta, eps, (bf16*)y, mean, rstd, M, D, \
                                                      seq_len, x_seq_stride, x_off, y_seq_stride, y_off)
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

extern "C" int avs_layernorm_bwd(const void* dy, const float* dpool, float pool_scale, const void* x,
                                 const float* mean, const float* rstd, const float* gamma, const void* resid,
                                 void* dx, float* dgamma, float* dbeta, int M, int D, int seq_len, int x_seq_stride,
                                 int x_off, int y_seq_stride, int y_off, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AVS_REQUIRE(x && mean && rstd && gamma && dx && dgamma && dbeta, "avs_layernorm_bwd: null pointer");
  AVS_REQUIRE(dy != nullptr || dpool != nullptr, "avs_layernorm_bwd: need dy and/or dpool");
  AVS_REQUIRE(dpool == nullptr || seq_len > 0, "avs_layernorm_bwd: dpool needs seq_len");
  AVS_REQUIRE(D % 8 == 0 && D <= 2048, "avs_layernorm_bwd: D must be a multiple of 8 and <= 2048 (got %d)", D);
  if (M == 0) return 0;
  const int blocks = min(avs_num_sms() * 2, ceil_div(M, 8));
  const size_t smem = 2 * (size_t)D * sizeof(float);
#define LN_BWD(N)                                                                                                  \
  layernorm_bwd_kernel<N><<<blocks, 256, smem, stream>>>((const bf16*)dy, dpool, pool_scale, (const bf16*)x, mean, \
                                                         rstd, gamma, (const bf16*)resid, (bf16*)dx, dgamma, dbeta, \
                                                         M, D, seq_len, x_seq_stride, x_off, y_seq_stride, y_off)
  switch (ln_nch(D)) {
    case 1: LN_BWD(1); break;
    case 2: LN_BWD(2); break;
    case 3: LN_BWD(3); break;
    case 4: LN_BWD(4); break;
    case 5: LN_BWD(5); break;
    default: LN_BWD(8); break;
  }
#undef LN_BWD
  return avs_check_launch("layernorm_bwd_kernel");
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
