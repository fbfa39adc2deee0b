// Classification heads of the finetune model (CAVMAEFT_BASE, cav_mae_base.py:813-816,842,875,1020-1035):
//   logits = Linear(LayerNorm(pooled))      pooled fp32 [B, Din] (Din = 768 or 1536), Linear [C, Din], C = 527 / 309
// plus the backward of the segment means that feed them (av[:, :512].mean(1) | av[:, 512:].mean(1), :1025-1028).
// < 0.1 % of the step's FLOPs and an odd class count (527 is not a multiple of 8, so the TMA GEMM does not apply):
// plain fp32 SIMT kernels, parameters read straight from the fp32 arena, no bf16 rounding anywhere.
#include "../../include/avsiam_b200.h"
#include "common.cuh"

namespace {

// one warp per row: xhat = (x - mean) * rstd (saved), y = xhat * gamma + beta
__global__ void head_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float* __restrict__ xhat,
                                   float* __restrict__ rstd_out, float* __restrict__ y, int B, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + (size_t)row * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += xr[c];
  const float mean = warp_sum(s) / D;
  float v = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float d = xr[c] - mean;
    v += d * d;
  }
  const float rstd = rsqrtf(warp_sum(v) / D + eps);
  for (int c = lane; c < D; c += 32) {
    const float h = (xr[c] - mean) * rstd;
    xhat[(size_t)row * D + c] = h;
    y[(size_t)row * D + c] = h * gamma[c] + beta[c];
  }
  if (lane == 0) rstd_out[row] = rstd;
}

// logits[b, c] = y[b, :] . W[c, :] + bias[c]   — one warp per (b, c)
__global__ void head_linear_fwd_kernel(const float* __restrict__ y, const float* __restrict__ W,
                                       const float* __restrict__ bias, float* __restrict__ out, int B, int C, int D) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * C) return;
  const int b = (int)(w / C), c = (int)(w % C);
  const float* yr = y + (size_t)b * D;
  const float* wr = W + (size_t)c * D;
  float acc = 0.f;
  for (int k = lane; k < D; k += 32) acc += yr[k] * wr[k];
  acc = warp_sum(acc);
  if (lane == 0) out[(size_t)b * C + c] = acc + bias[c];
}

// dy_ln[b, k] = sum_c dlogits[b, c] W[c, k]   — thread per (b, k), W rows read coalesced
__global__ void head_linear_dgrad_kernel(const float* __restrict__ dl, const float* __restrict__ W,
                                         float* __restrict__ dy, int B, int C, int D) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (k >= D) return;
  const float* dr = dl + (size_t)b * C;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc += dr[c] * W[(size_t)c * D + k];
  dy[(size_t)b * D + k] = acc;
}

// dW[c, k] += sum_b dlogits[b, c] y[b, k] ; dbias[c] += sum_b dlogits[b, c]   — thread per (c, k)
__global__ void head_linear_wgrad_kernel(const float* __restrict__ dl, const float* __restrict__ y,
                                         float* __restrict__ dW, float* __restrict__ dbias, int B, int C, int D) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (k >= D) return;
  float acc = 0.f, accb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = dl[(size_t)b * C + c];
    acc += g * y[(size_t)b * D + k];
    accb += g;
  }
  dW[(size_t)c * D + k] += acc;
  if (k == 0) dbias[c] += accb;
}

// LayerNorm backward on [B, D] fp32: dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)); one warp per row.
// dgamma / dbeta are accumulated with atomics (B rows, D columns: tiny).
__global__ void head_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xhat,
                                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                                   float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                   int B, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* dr = dy + (size_t)row * D;
  const float* hr = xhat + (size_t)row * D;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float gd = gamma[c] * dr[c];
    s1 += gd;
    s2 += gd * hr[c];
  }
  s1 = warp_sum(s1) / D;
  s2 = warp_sum(s2) / D;
  const float r = rstd[row];
  for (int c = lane; c < D; c += 32) {
    const float d = dr[c], h = hr[c];
    dx[(size_t)row * D + c] = r * (gamma[c] * d - s1 - h * s2);
    atomicAdd(dgamma + c, d * h);
    atomicAdd(dbeta + c, d);
  }
}

// dx[row(s, t), c] (bf16, overwrite) = scale * dpool[s, c]  for t in [0, S): backward of a token mean over a segment
// of each sequence (rows s*stride + off + t)
__global__ void seq_mean_bwd_kernel(const float* __restrict__ dpool, bf16* __restrict__ dx, int S, int D, int stride,
                                    int off, float scale, long long total_chunks) {
  const int cpr = D / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cpr;
    const int c = (int)(i % cpr) * 8;
    const long long s = r / S;
    const int t = (int)(r % S);
    const float4 p0 = *reinterpret_cast<const float4*>(dpool + s * D + c);
    const float4 p1 = *reinterpret_cast<const float4*>(dpool + s * D + c + 4);
    uint4 o;
    o.x = pack_bf16x2(p0.x * scale, p0.y * scale);
    o.y = pack_bf16x2(p0.z * scale, p0.w * scale);
    o.z = pack_bf16x2(p1.x * scale, p1.y * scale);
    o.w = pack_bf16x2(p1.z * scale, p1.w * scale);
    *reinterpret_cast<uint4*>(dx + (s * stride + off + t) * D + c) = o;
  }
}

}  // namespace

extern "C" int avs_head_fwd(const float* x, const float* gamma, const float* beta, float eps, const float* W,
                            const float* bias, float* xhat, float* rstd, float* y, float* logits, int B, int C, int D,
                            void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AVS_REQUIRE(x && gamma && beta && W && bias && xhat && rstd && y && logits, "avs_head_fwd: null pointer");
  AVS_REQUIRE(B >= 0 && C > 0 && D > 0, "avs_head_fwd: bad shape");
  if (B == 0) return 0;
  head_ln_fwd_kernel<<<ceil_div(B, 8), 256, 0, stream>>>(x, gamma, beta, eps, xhat, rstd, y, B, D);
  int rc = avs_check_launch("head_ln_fwd_kernel");
  if (rc) return rc;
  const long long warps = (long long)B * C;
  head_linear_fwd_kernel<<<(unsigned)ceil_div_ll(warps * 32, 256), 256, 0, stream>>>(y, W, bias, logits, B, C, D);
  return avs_check_launch("head_linear_fwd_kernel");
}

// dW, dbias, dgamma, dbeta are ACCUMULATED; dx is overwritten. `dy_scratch` [B, D] fp32.
extern "C" int avs_head_bwd(const float* dlogits, const float* xhat, const float* rstd, const float* y,
                            const float* gamma, const float* W, float* dW, float* dbias, float* dgamma, float* dbeta,
                            float* dy_scratch, float* dx, int B, int C, int D, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AVS_REQUIRE(dlogits && xhat && rstd && y && gamma && W && dW && dbias && dgamma && dbeta && dy_scratch && dx,
              "avs_head_bwd: null pointer");
  AVS_REQUIRE(B >= 0 && C > 0 && D > 0 && B <= 65535 && C <= 65535, "avs_head_bwd: bad shape");
  if (B == 0) return 0;
  head_linear_dgrad_kernel<<<dim3(ceil_div(D, 128), B), 128, 0, stream>>>(dlogits, W, dy_scratch, B, C, D);
  int rc = avs_check_launch("head_linear_dgrad_kernel");
  if (rc) return rc;
  head_linear_wgrad_kernel<<<dim3(ceil_div(D, 128), C), 128, 0, stream>>>(dlogits, y, dW, dbias, B, C, D);
  if ((rc = avs_check_launch("head_linear_wgrad_kernel"))) return rc;
  head_ln_bwd_kernel<<<ceil_div(B, 8), 256, 0, stream>>>(dy_scratch, xhat, rstd, gamma, dx, dgamma, dbeta, B, D);
  return avs_check_launch("head_ln_bwd_kernel");
}

extern "C" int avs_seq_mean_bwd(const float* dpool, void* dx, int n_seq, int seg_len, int D, int seq_stride, int off,
                                void* stream) {
  AVS_REQUIRE(dpool && dx, "avs_seq_mean_bwd: null pointer");
  AVS_REQUIRE(D % 8 == 0 && seg_len > 0 && seq_stride >= seg_len && off >= 0, "avs_seq_mean_bwd: bad shape");
  if (n_seq == 0) return 0;
  const long long chunks = (long long)n_seq * seg_len * (D / 8);
  const int blocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll(chunks, 256));
  seq_mean_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dpool, (bf16*)dx, seg_len, D, seq_stride, off,
                                                                 1.0f / seg_len, chunks);
  return avs_check_launch("seq_mean_bwd_kernel");
}
