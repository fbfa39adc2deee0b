// Finetune / evaluation operators outside the transformer trunk:
//   * classification losses of the finetune loop (traintest_ft_base.py:106-109,148-158): nn.BCEWithLogitsLoss() and
//     nn.CrossEntropyLoss() with probability targets (the loader's label vectors, dataloader.py:497-503), mean
//     reduction, forward and gradient in one pass over the fp32 logits;
//   * retrieval (retrieval.py:27-52): the cosine-similarity matrix that the reference fills with a Python double loop,
//     and, per query row, how many candidates score strictly higher than / exactly equal to the matching (diagonal)
//     candidate — everything compute_metrics derives from its sort.
// HBM-/latency-bound fp32 SIMT kernels; no tensor cores (the similarity is ranked, so it stays in fp32).
#include "../../include/avsiam_b200.h"
#include "common.cuh"

namespace {

// loss += sum_i [max(x,0) - x*y + log1p(exp(-|x|))] / n ;  dx = (sigmoid(x) - y) / n        (torch's stable form)
__global__ void bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ loss,
                                  float* __restrict__ dx, long long n, float inv_n) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i], t = y[i];
    const float e = __expf(-fabsf(v));
    acc += fmaxf(v, 0.f) - v * t + log1pf(e);
    if (dx) {
      const float sig = v >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
      dx[i] = (sig - t) * inv_n;
    }
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss, v * inv_n);
  }
}

// one warp per row: loss += -sum_c y_c * log_softmax(x)_c / B ;  dx_c = (softmax_c * sum(y) - y_c) / B
__global__ void ce_prob_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ loss,
                               float* __restrict__ dx, int B, int C, float inv_b) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + (size_t)row * C;
  const float* yr = y + (size_t)row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, xr[c]);
  mx = warp_max(mx);
  float se = 0.f, sy = 0.f, sxy = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = xr[c] - mx, t = yr[c];
    se += __expf(v);
    sy += t;
    sxy += t * v;
  }
  se = warp_sum(se); sy = warp_sum(sy); sxy = warp_sum(sxy);
  const float lse = logf(se);
  if (lane == 0) atomicAdd(loss, (sy * lse - sxy) * inv_b);
  if (dx) {
    const float inv_se = 1.f / se;
    for (int c = lane; c < C; c += 32)
      dx[(size_t)row * C + c] = (__expf(xr[c] - mx) * inv_se * sy - yr[c]) * inv_b;
  }
}

// ||row||_2 of fp32 [n, d]; one warp per row
__global__ void row_norm_kernel(const float* __restrict__ x, float* __restrict__ nrm, int n, int d) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = x[(size_t)row * d + c];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) nrm[row] = sqrtf(s);
}

// sim[i, j] = (a_i . b_j) / (|a_i| * |b_j|)  (retrieval.py:27-29). 64 x 64 tile per CTA, 16 x 16 threads, each a
// 4 x 4 register tile, K staged 16 at a time through shared memory (transposed so the inner loop reads float4).
constexpr int SIM_T = 64, SIM_K = 16;
__global__ void __launch_bounds__(256) sim_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                  const float* __restrict__ na, const float* __restrict__ nb,
                                                  float* __restrict__ sim, int n, int m, int d) {
  __shared__ __align__(16) float sa[SIM_K][SIM_T + 4];
  __shared__ __align__(16) float sb[SIM_K][SIM_T + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * SIM_T, j0 = blockIdx.x * SIM_T;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int k0 = 0; k0 < d; k0 += SIM_K) {
    for (int e = threadIdx.x; e < SIM_T * SIM_K; e += 256) {
      const int r = e / SIM_K, k = e % SIM_K;
      const bool kin = k0 + k < d;
      sa[k][r] = (kin && i0 + r < n) ? a[(size_t)(i0 + r) * d + k0 + k] : 0.f;
      sb[k][r] = (kin && j0 + r < m) ? b[(size_t)(j0 + r) * d + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SIM_K; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&sa[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sb[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(ar[r], br[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= n) continue;
    const float ni = na[i];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j < m) sim[(size_t)i * m + j] = acc[r][c] / (ni * nb[j]);
    }
  }
}

// per query row i of sim [n, n]: greater[i] = #{j : sim[i,j] > sim[i,i]}, equal[i] = #{j : sim[i,j] == sim[i,i]} (>= 1).
// compute_metrics (retrieval.py:39-52) sorts -x along each row and lists every position holding the diagonal's value:
// those positions are greater[i] .. greater[i] + equal[i] - 1.
__global__ void rank_kernel(const float* __restrict__ sim, int* __restrict__ greater, int* __restrict__ equal, int n) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* r = sim + (size_t)row * n;
  const float dg = r[row];
  int g = 0, e = 0;
  for (int j = lane; j < n; j += 32) {
    const float v = r[j];
    g += v > dg;
    e += v == dg;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    g += __shfl_xor_sync(0xffffffffu, g, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
  }
  if (lane == 0) {
    greater[row] = g;
    equal[row] = e;
  }
}

}  // namespace

extern "C" int avs_bce_with_logits(const float* logits, const float* target, float* loss, float* dlogits, long long n,
                                   void* stream) {
  AVS_REQUIRE(logits && target && loss, "avs_bce_with_logits: null pointer");
  AVS_REQUIRE(n >= 0, "avs_bce_with_logits: bad size");
  if (n == 0) return 0;
  const int blocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll(n, 256));
  bce_logits_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, target, loss, dlogits, n, 1.0f / (float)n);
  return avs_check_launch("bce_logits_kernel");
}

extern "C" int avs_cross_entropy_prob(const float* logits, const float* target, float* loss, float* dlogits, int B,
                                      int C, void* stream) {
  AVS_REQUIRE(logits && target && loss, "avs_cross_entropy_prob: null pointer");
  AVS_REQUIRE(B >= 0 && C > 0, "avs_cross_entropy_prob: bad shape");
  if (B == 0) return 0;
  ce_prob_kernel<<<ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(logits, target, loss, dlogits, B, C, 1.0f / B);
  return avs_check_launch("ce_prob_kernel");
}

extern "C" int avs_cosine_sim(const float* a, const float* b, int n, int m, int d, float* norm_scratch, float* sim,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AVS_REQUIRE(a && b && norm_scratch && sim, "avs_cosine_sim: null pointer");
  AVS_REQUIRE(n >= 0 && m >= 0 && d > 0, "avs_cosine_sim: bad shape");
  if (n == 0 || m == 0) return 0;
  AVS_REQUIRE(ceil_div(n, SIM_T) <= 65535, "avs_cosine_sim: too many query rows");
  row_norm_kernel<<<ceil_div(n, 8), 256, 0, stream>>>(a, norm_scratch, n, d);
  int rc = avs_check_launch("row_norm_kernel");
  if (rc) return rc;
  row_norm_kernel<<<ceil_div(m, 8), 256, 0, stream>>>(b, norm_scratch + n, m, d);
  if ((rc = avs_check_launch("row_norm_kernel"))) return rc;
  sim_kernel<<<dim3(ceil_div(m, SIM_T), ceil_div(n, SIM_T)), 256, 0, stream>>>(a, b, norm_scratch, norm_scratch + n,
                                                                             sim, n, m, d);
  return avs_check_launch("sim_kernel");
}

extern "C" int avs_retrieval_ranks(const float* sim, int n, int* greater, int* equal, void* stream) {
  AVS_REQUIRE(sim && greater && equal, "avs_retrieval_ranks: null pointer");
  AVS_REQUIRE(n >= 0, "avs_retrieval_ranks: bad shape");
  if (n == 0) return 0;
  rank_kernel<<<ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(sim, greater, equal, n);
  return avs_check_launch("rank_kernel");
}
