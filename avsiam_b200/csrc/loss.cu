// Fused losses: masked-MSE reconstruction (target read straight from the raw fbank / frame, never materialised)
// and global-batch InfoNCE over L2-normalised pooled embeddings, forward + backward.
//
// Replaces patchify + forward_mae_loss (cav_mae_base.py:343-351,663-683) and forward_contrastive (:641-661).
#include "../../include/avsiam_b200.h"
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// masked MSE.   pred bf16 [B*T, P];   kind 0: audio [B, Tlen, F] -> P = p*p, token (f,t), vec (pf,pt)
//                                     kind 1: video [B, C, H, W] -> P = p*p*C, token (h,w), vec (p,q,c)
//   fwd : loss += sum_{masked rows} mean_e (pred-target)^2 / n_masked
//   bwd : dpred = mask * 2 (pred-target) / (P * n_masked) * upstream        (zeros on kept rows)
// One warp per (b, token) row.
// ---------------------------------------------------------------------------------------------------------
struct MaeGeom {
  int kind, p, C;
  int d0, d1;  // audio: Tlen, F   video: H, W
  int gw;      // tokens per row of the token grid (audio: ta, video: W/p)
  int T;       // tokens per sample
  int P;       // elements per patch
};

// PS = patch size known at compile time (16 for every AVSiam model: divisions become shifts / multiply-shifts, which is
// what this kernel's time went into) or 0 for the generic path.  (r, c) = the token's position in the token grid.
template <int PS>
__device__ __forceinline__ float mae_target(const float* __restrict__ in, const MaeGeom& g, int b, int r, int c, int e) {
  const int p = PS > 0 ? PS : g.p;
  if (g.kind == 0) {
    const int pf = e / p, pt = e % p;  // r = f block, c = t block
    return __ldg(in + ((size_t)b * g.d0 + (size_t)c * p + pt) * g.d1 + r * p + pf);
  } else if (PS > 0 && g.C == 3) {
    const int ch = e % 3, q = (e / 3) % p, pp = e / (3 * p);
    return __ldg(in + (((size_t)b * 3 + ch) * g.d0 + (size_t)r * p + pp) * g.d1 + (size_t)c * p + q);
  } else {
    const int ch = e % g.C, q = (e / g.C) % p, pp = e / (g.C * p);
    return __ldg(in + (((size_t)b * g.C + ch) * g.d0 + (size_t)r * p + pp) * g.d1 + (size_t)c * p + q);
  }
}

template <bool BWD, int PS>
__global__ void __launch_bounds__(256) mae_loss_kernel(const bf16* __restrict__ pred, const float* __restrict__ in,
                                                       const float* __restrict__ mask, MaeGeom g, int rows,
                                                       float inv_n_masked, float* __restrict__ loss,
                                                       const float* __restrict__ upstream, bf16* __restrict__ dpred) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float local = 0.f;
  const float up = BWD ? (upstream ? *upstream : 1.0f) * 2.0f * inv_n_masked / g.P : 0.f;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const float m = mask[r];
    const int b = r / g.T, tok = r - b * g.T;
    const int tr = tok / g.gw, tc = tok - tr * g.gw;
    if (m == 0.f) {
      if (BWD)
        for (int e = lane * 8; e < g.P; e += 256) *reinterpret_cast<uint4*>(dpred + (size_t)r * g.P + e) = make_uint4(0, 0, 0, 0);
      continue;
    }
    float acc = 0.f;
    for (int e = lane * 8; e < g.P; e += 256) {
      const uint4 u = *reinterpret_cast<const uint4*>(pred + (size_t)r * g.P + e);
      float pv[8];
      float2 f;
      f = unpack_bf16x2(u.x); pv[0] = f.x; pv[1] = f.y;
      f = unpack_bf16x2(u.y); pv[2] = f.x; pv[3] = f.y;
      f = unpack_bf16x2(u.z); pv[4] = f.x; pv[5] = f.y;
      f = unpack_bf16x2(u.w); pv[6] = f.x; pv[7] = f.y;
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d[j] = pv[j] - mae_target<PS>(in, g, b, tr, tc, e + j);
        acc += d[j] * d[j];
      }
      if (BWD) {
        uint4 o;
        o.x = pack_bf16x2(d[0] * up * m, d[1] * up * m); o.y = pack_bf16x2(d[2] * up * m, d[3] * up * m);
        o.z = pack_bf16x2(d[4] * up * m, d[5] * up * m); o.w = pack_bf16x2(d[6] * up * m, d[7] * up * m);
        *reinterpret_cast<uint4*>(dpred + (size_t)r * g.P + e) = o;
      }
    }
    if (!BWD) {
      acc = warp_sum(acc);
      local += acc / g.P * m;
    }
  }
  if (!BWD) {
    __shared__ float s[8];
    if (lane == 0) s[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += s[i];
      atomicAdd(loss, t * inv_n_masked);
    }
  }
}

static int mae_geom(MaeGeom& g, int kind, int patch, int C, int d0, int d1) {
  g.kind = kind; g.p = patch;
  if (kind == 0) {
    g.C = 1; g.d0 = d0; g.d1 = d1;   // Tlen, F
    g.gw = d0 / patch;               // ta
    g.T = (d0 / patch) * (d1 / patch);
    g.P = patch * patch;
  } else {
    g.C = C; g.d0 = d0; g.d1 = d1;   // H, W
    g.gw = d1 / patch;
    g.T = (d0 / patch) * (d1 / patch);
    g.P = patch * patch * C;
  }
  AVS_REQUIRE(g.P % 8 == 0, "mae loss: patch vector length must be a multiple of 8");
  return 0;
}

extern "C" int avs_mae_loss_fwd(const void* pred, const float* input, const float* mask, int kind, int B, int patch,
                                int C, int d0, int d1, float n_masked, float* loss_accum, void* stream) {
  AVS_REQUIRE(pred && input && mask && loss_accum, "avs_mae_loss_fwd: null pointer");
  AVS_REQUIRE(n_masked > 0, "avs_mae_loss_fwd: n_masked must be > 0");
  MaeGeom g;
  if (mae_geom(g, kind, patch, C, d0, d1)) return -1;
  const int rows = B * g.T;
  if (rows == 0) return 0;
  const int blocks = min(avs_num_sms() * 8, ceil_div(rows, 8));
  if (patch == 16)
    mae_loss_kernel<false, 16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)pred, input, mask, g, rows,
                                                                         1.0f / n_masked, loss_accum, nullptr, nullptr);
  else
    mae_loss_kernel<false, 0><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)pred, input, mask, g, rows,
                                                                        1.0f / n_masked, loss_accum, nullptr, nullptr);
  return avs_check_launch("mae_loss_kernel<fwd>");
}

extern "C" int avs_mae_loss_bwd(const void* pred, const float* input, const float* mask, int kind, int B, int patch,
                                int C, int d0, int d1, float n_masked, const float* upstream, void* dpred,
                                void* stream) {
  AVS_REQUIRE(pred && input && mask && dpred, "avs_mae_loss_bwd: null pointer");
  AVS_REQUIRE(n_masked > 0, "avs_mae_loss_bwd: n_masked must be > 0");
  MaeGeom g;
  if (mae_geom(g, kind, patch, C, d0, d1)) return -1;
  const int rows = B * g.T;
  if (rows == 0) return 0;
  const int blocks = min(avs_num_sms() * 8, ceil_div(rows, 8));
  if (patch == 16)
    mae_loss_kernel<true, 16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)pred, input, mask, g, rows,
                                                                        1.0f / n_masked, nullptr, upstream, (bf16*)dpred);
  else
    mae_loss_kernel<true, 0><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)pred, input, mask, g, rows,
                                                                       1.0f / n_masked, nullptr, upstream, (bf16*)dpred);
  return avs_check_launch("mae_loss_kernel<bwd>");
}

// ---------------------------------------------------------------------------------------------------------
// InfoNCE (fp32 throughout; N = global batch <= a few thousand, so the [N,N] logits are tiny).
// workspace (floats): ah[N*D] vh[N*D] inv_na[N] inv_nv[N] S[N*N] row_lse[N] col_lse[N] row_arg[N] col_arg[N]
// ---------------------------------------------------------------------------------------------------------
struct NceWs {
  float *ah, *vh, *ina, *inv, *S, *row_lse, *col_lse;
  int *row_arg, *col_arg;
};
static NceWs nce_carve(float* w, int N, int D) {
  NceWs s;
  s.ah = w; w += (size_t)N * D;
  s.vh = w; w += (size_t)N * D;
  s.ina = w; w += N;
  s.inv = w; w += N;
  s.S = w; w += (size_t)N * N;
  s.row_lse = w; w += N;
  s.col_lse = w; w += N;
  s.row_arg = (int*)w; w += N;
  s.col_arg = (int*)w; w += N;
  return s;
}
extern "C" size_t avs_infonce_workspace_bytes(int N, int D) {
  return ((size_t)2 * N * D + (size_t)N * N + 6 * (size_t)N) * sizeof(float);
}

__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ xh, float* __restrict__ inv_norm,
                                  int D) {
  const int r = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = x[(size_t)r * D + c];
    s += v * v;
  }
  __shared__ float sh[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) sh[0] = 1.0f / fmaxf(sqrtf(t), 1e-12f);  // F.normalize eps
  }
  __syncthreads();
  const float inv = sh[0];
  if (threadIdx.x == 0) inv_norm[r] = inv;
  for (int c = threadIdx.x; c < D; c += blockDim.x) xh[(size_t)r * D + c] = x[(size_t)r * D + c] * inv;
}

// C[i,j] = alpha * sum_k A[i*ai + k*ak] * B[j*bj + k*bk]   (32x32 tiles, fp32 SIMT; only used on tiny problems)
__global__ void sgemm_strided_kernel(const float* __restrict__ A, long long ai, long long ak,
                                     const float* __restrict__ B, long long bj, long long bk, float* __restrict__ C,
                                     long long ldc, int M, int N, int K, float alpha) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[4] = {0, 0, 0, 0};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = ty + 8 * u;
      const int k = k0 + tx;
      const int i = i0 + rr, j = j0 + rr;
      // pick the faster-varying thread index along the contiguous memory direction
      sa[rr][tx] = (i < M && k < K) ? A[i * ai + k * ak] : 0.f;
      sb[rr][tx] = (j < N && k < K) ? B[j * bj + k * bk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float b = sb[tx][kk];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += sa[ty + 8 * u][kk] * b;
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty + 8 * u, j = j0 + tx;
    if (i < M && j < N) C[i * ldc + j] = alpha * acc[u];
  }
}
// Same contraction with 64x64 tiles and 4x4 outputs per thread (16 FMA per 2 float4 shared-memory reads): the global
// InfoNCE logits are N^2 D fp32 multiply-adds with N = world x batch — 6.4 GFLOP at 8 x 256 — and every rank computes
// all of them, so this product is what grows with W^2 in the 8-GPU step (1.1 ms with the 32x32 kernel above).
__global__ void __launch_bounds__(256) sgemm64_kernel(const float* __restrict__ A, long long ai, long long ak,
                                                      const float* __restrict__ B, long long bj, long long bk,
                                                      float* __restrict__ C, long long ldc, int M, int N, int K,
                                                      float alpha) {
  __shared__ __align__(16) float sa[16][68], sb[16][68];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = tid + 256 * u;
      // the index that is contiguous in memory varies fastest across the threads
      const int r = (ak == 1) ? idx >> 4 : idx & 63, k = (ak == 1) ? idx & 15 : idx >> 6;
      const int rb = (bk == 1) ? idx >> 4 : idx & 63, kb = (bk == 1) ? idx & 15 : idx >> 6;
      sa[k][r] = (i0 + r < M && k0 + k < K) ? A[(long long)(i0 + r) * ai + (long long)(k0 + k) * ak] : 0.f;
      sb[kb][rb] = (j0 + rb < N && k0 + kb < K) ? B[(long long)(j0 + rb) * bj + (long long)(k0 + kb) * bk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sa[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sb[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty * 4 + u;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = j0 + tx * 4 + v;
      if (i < M && j < N) C[(long long)i * ldc + j] = alpha * acc[u][v];
    }
  }
}
static void sgemm(const float* A, long long ai, long long ak, const float* B, long long bj, long long bk, float* C,
                  long long ldc, int M, int N, int K, float alpha, cudaStream_t st) {
  if ((long long)M * N >= 64 * 64 * 16) {   // enough 64x64 tiles to matter; the small kernel keeps more CTAs busy below
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
    sgemm64_kernel<<<grid, 256, 0, st>>>(A, ai, ak, B, bj, bk, C, ldc, M, N, K, alpha);
    return;
  }
  dim3 grid(ceil_div(N, 32), ceil_div(M, 32)), block(32, 8);
  sgemm_strided_kernel<<<grid, block, 0, st>>>(A, ai, ak, B, bj, bk, C, ldc, M, N, K, alpha);
}

// block r < N: row r statistics; block r >= N: column (r-N) statistics. lse = log sum exp, arg = first argmax.
__global__ void nce_stats_kernel(const float* __restrict__ S, int N, float* __restrict__ row_lse,
                                 float* __restrict__ col_lse, int* __restrict__ row_arg, int* __restrict__ col_arg) {
  const bool is_col = blockIdx.x >= N;
  const int r = is_col ? blockIdx.x - N : blockIdx.x;
  const long long stride = is_col ? N : 1;
  const float* base = is_col ? S + r : S + (size_t)r * N;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float v = base[i * stride];
    if (v > mx || (v == mx && i < arg)) mx = v, arg = i;
  }
  __shared__ float smx[32];
  __shared__ int sarg[32];
  __shared__ float ssum[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) mx = om, arg = oa;
  }
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) smx[w] = mx, sarg[w] = arg;
  __syncthreads();
  mx = smx[0]; arg = sarg[0];
  for (int i = 1; i < nw; ++i)
    if (smx[i] > mx || (smx[i] == mx && sarg[i] < arg)) mx = smx[i], arg = sarg[i];
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += expf(base[i * stride] - mx);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ssum[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += ssum[i];
    const float lse = mx + logf(t);
    if (is_col) col_lse[r] = lse, col_arg[r] = arg;
    else row_lse[r] = lse, row_arg[r] = arg;
  }
}

// single block: loss and accuracy.  nce_1 (dim=0 softmax = column-wise) always; nce_2 (row-wise) if bidirect.
__global__ void nce_finalize_kernel(const float* __restrict__ S, int N, const float* __restrict__ row_lse,
                                    const float* __restrict__ col_lse, const int* __restrict__ row_arg,
                                    const int* __restrict__ col_arg, int bidirect, float* __restrict__ loss,
                                    float* __restrict__ acc) {
  float l = 0.f, a = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float d = S[(size_t)i * N + i];
    l += col_lse[i] - d;
    a += (col_arg[i] == i) ? 1.f : 0.f;
    if (bidirect) {
      l += row_lse[i] - d;
      a += (row_arg[i] == i) ? 1.f : 0.f;
    }
  }
  __shared__ float sl[32], sa[32];
  l = warp_sum(l); a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sl[threadIdx.x >> 5] = l, sa[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tl = 0.f, ta = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) tl += sl[i], ta += sa[i];
    const float denom = bidirect ? 2.f * N : (float)N;
    *loss = tl / denom;
    *acc = ta / denom;
  }
}

extern "C" int avs_infonce_fwd(const float* ea, const float* ev, int N, int D, float temperature, int bidirect,
                               float* workspace, float* loss_out, float* acc_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  AVS_REQUIRE(ea && ev && workspace && loss_out && acc_out, "avs_infonce_fwd: null pointer");
  AVS_REQUIRE(N > 0 && D > 0 && temperature > 0, "avs_infonce_fwd: bad shape");
  NceWs w = nce_carve(workspace, N, D);
  l2norm_fwd_kernel<<<N, 128, 0, st>>>(ea, w.ah, w.ina, D);
  int rc = avs_check_launch("l2norm_fwd_kernel");
  if (rc) return rc;
  l2norm_fwd_kernel<<<N, 128, 0, st>>>(ev, w.vh, w.inv, D);
  if ((rc = avs_check_launch("l2norm_fwd_kernel"))) return rc;
  sgemm(w.ah, D, 1, w.vh, D, 1, w.S, N, N, N, D, 1.0f / temperature, st);
  if ((rc = avs_check_launch("sgemm_strided_kernel"))) return rc;
  nce_stats_kernel<<<2 * N, 256, 0, st>>>(w.S, N, w.row_lse, w.col_lse, w.row_arg, w.col_arg);
  if ((rc = avs_check_launch("nce_stats_kernel"))) return rc;
  nce_finalize_kernel<<<1, 256, 0, st>>>(w.S, N, w.row_lse, w.col_lse, w.row_arg, w.col_arg, bidirect, loss_out,
                                         acc_out);
  return avs_check_launch("nce_finalize_kernel");
}

// G = dL/dS (in place over S):  w/denom * [softmax_col(S) - I  (+ softmax_row(S) - I)]
__global__ void nce_grad_logits_kernel(float* __restrict__ S, int N, const float* __restrict__ row_lse,
                                       const float* __restrict__ col_lse, int bidirect, float weight,
                                       const float* __restrict__ upstream) {
  const float w = weight * (upstream ? *upstream : 1.0f) / (bidirect ? 2.f * N : (float)N);
  const long long total = (long long)N * N;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(g / N), j = (int)(g % N);
    const float s = S[g];
    float v = expf(s - col_lse[j]) - (i == j ? 1.f : 0.f);
    if (bidirect) v += expf(s - row_lse[i]) - (i == j ? 1.f : 0.f);
    S[g] = v * w;
  }
}

// dx = (dxh - xh * (xh . dxh)) * inv_norm   for rows [row0, row0+rows)
__global__ void l2norm_bwd_kernel(const float* __restrict__ xh, const float* __restrict__ inv_norm,
                                  const float* __restrict__ dxh, float* __restrict__ dx, int row0, int D) {
  const int lr = blockIdx.x, r = row0 + lr;
  float s = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) s += xh[(size_t)r * D + c] * dxh[(size_t)lr * D + c];
  __shared__ float sh[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    sh[0] = t;
  }
  __syncthreads();
  const float dot = sh[0], inv = inv_norm[r];
  for (int c = threadIdx.x; c < D; c += blockDim.x)
    dx[(size_t)lr * D + c] = (dxh[(size_t)lr * D + c] - xh[(size_t)r * D + c] * dot) * inv;
}

// Gradients w.r.t. the un-normalised embeddings of rows [row0, row0+rows) (this rank's slice of the global batch).
// `scratch`: 2*rows*D floats. Consumes (overwrites) the logits held in the workspace by avs_infonce_fwd.
extern "C" int avs_infonce_bwd(int N, int D, float temperature, int bidirect, float weight, const float* upstream,
                               float* workspace, int row0, int rows, float* scratch, float* d_ea, float* d_ev,
                               void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  AVS_REQUIRE(workspace && scratch && d_ea && d_ev, "avs_infonce_bwd: null pointer");
  AVS_REQUIRE(row0 >= 0 && rows > 0 && row0 + rows <= N, "avs_infonce_bwd: bad row range");
  NceWs w = nce_carve(workspace, N, D);
  const int blocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll((long long)N * N, 256));
  nce_grad_logits_kernel<<<blocks, 256, 0, st>>>(w.S, N, w.row_lse, w.col_lse, bidirect, weight, upstream);
  int rc = avs_check_launch("nce_grad_logits_kernel");
  if (rc) return rc;
  float* dah = scratch;
  float* dvh = scratch + (size_t)rows * D;
  // dAh[i,:] = (1/T) sum_j G[i,j] vh[j,:]      i in slice
  sgemm(w.S + (size_t)row0 * N, N, 1, w.vh, 1, D, dah, D, rows, D, N, 1.0f / temperature, st);
  if ((rc = avs_check_launch("sgemm_strided_kernel"))) return rc;
  // dVh[j,:] = (1/T) sum_i G[i,j] ah[i,:]      j in slice
  sgemm(w.S + row0, 1, N, w.ah, 1, D, dvh, D, rows, D, N, 1.0f / temperature, st);
  if ((rc = avs_check_launch("sgemm_strided_kernel"))) return rc;
  l2norm_bwd_kernel<<<rows, 128, 0, st>>>(w.ah, w.ina, dah, d_ea, row0, D);
  if ((rc = avs_check_launch("l2norm_bwd_kernel"))) return rc;
  l2norm_bwd_kernel<<<rows, 128, 0, st>>>(w.vh, w.inv, dvh, d_ev, row0, D);
  return avs_check_launch("l2norm_bwd_kernel");
}
