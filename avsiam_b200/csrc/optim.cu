// Flat-buffer optimizer and parameter plumbing kernels (HBM-bound, 128-bit accesses):
//   fused Adam / AdamW step over one contiguous fp32 parameter arena (+ bf16 shadow write, + GradScaler unscale /
//   inf-check), fp32->bf16 cast, bias-gradient column sums, and an inf/nan scan.
// Replaces torch.optim.Adam.step + GradScaler.unscale_ (traintest_cavmae_base.py:64-66,138-140,149-152).
#include "../../include/avsiam_b200.h"
#include "common.cuh"

// torch.optim.Adam semantics (non-amsgrad): coupled L2 (g += wd*p) or decoupled (AdamW: p *= 1 - lr*wd).
// found_inf (optional, device float flag != 0) => skip the whole step (GradScaler.step behaviour).
// inv_scale (optional, device float) multiplies every gradient first (GradScaler.unscale_).
// active (optional, one byte per 64-element chunk of the arena): 0 => the chunk belongs to a parameter whose
// .grad is None in the reference or that was not handed to the optimizer (torch.optim skips those entirely — no
// weight decay, no moment update); k > 0 => the chunk belongs to param_group k-1, whose lr / weight_decay apply
// (the finetune loop builds three groups, traintest_ft_base.py:78-83).
// tick (optional, device {int step; float bc1; float bc2_sqrt}): the step counter lives on the device and advances only
// when the step is not skipped, so the bias corrections stay torch's after a GradScaler overflow.
#define AVS_ADAM_MAX_GROUPS 8
struct AdamGroups {
  float lr[AVS_ADAM_MAX_GROUPS];
  float wd[AVS_ADAM_MAX_GROUPS];
};
struct AdamTick {
  int step;
  float bc1, bc2_sqrt;
};

__global__ void adam_tick_kernel(AdamTick* __restrict__ tick, const float* __restrict__ found_inf, float beta1,
                                 float beta2) {
  if (found_inf != nullptr && *found_inf != 0.f) return;
  const int step = tick->step + 1;
  tick->step = step;
  tick->bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  tick->bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   bf16* __restrict__ shadow, long long n, AdamGroups groups,
                                                   float beta1, float beta2, float eps, float bc1_host,
                                                   float bc2_sqrt_host, int decoupled,
                                                   const float* __restrict__ inv_scale,
                                                   const float* __restrict__ found_inf,
                                                   const uint8_t* __restrict__ active,
                                                   const AdamTick* __restrict__ tick) {
  if (found_inf != nullptr && *found_inf != 0.f) return;
  const float gs = inv_scale ? *inv_scale : 1.0f;
  const float bc1 = tick ? tick->bc1 : bc1_host, bc2_sqrt = tick ? tick->bc2_sqrt : bc2_sqrt_host;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    int grp = 0;
    if (active != nullptr) {
      const int a = active[i >> 4];
      if (a == 0) continue;  // parameter received no gradient / is not the optimizer's: Adam skips it
      grp = min(a - 1, AVS_ADAM_MAX_GROUPS - 1);
    }
    const float lr = groups.lr[grp], wd = groups.wd[grp];
    const float step_size = lr / bc1;
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x * gs, gg.y * gs, gg.z * gs, gg.w * gs};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gj = ga[j];
      if (decoupled) pa[j] *= 1.0f - lr * wd;
      else gj += wd * pa[j];
      ma[j] = beta1 * ma[j] + (1.0f - beta1) * gj;
      va[j] = beta2 * va[j] + (1.0f - beta2) * gj * gj;
      const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
      pa[j] -= step_size * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (shadow != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(pa[0], pa[1]);
      o.y = pack_bf16x2(pa[2], pa[3]);
      reinterpret_cast<uint2*>(shadow)[i] = o;
    }
  }
}

extern "C" int avs_adam_step_groups(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n,
                                    const float* lr, const float* weight_decay, int n_groups, float beta1, float beta2,
                                    float eps, int step, int decoupled, const float* inv_scale, const float* found_inf,
                                    const uint8_t* group_chunks, void* tick_state, void* stream) {
  AVS_REQUIRE(p && g && m && v && lr && weight_decay, "avs_adam_step: null pointer");
  AVS_REQUIRE(n % 4 == 0, "avs_adam_step: n must be a multiple of 4 (pad the arena)");
  AVS_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "avs_adam_step: 16-byte alignment");
  AVS_REQUIRE(n_groups >= 1 && n_groups <= AVS_ADAM_MAX_GROUPS, "avs_adam_step: 1..%d parameter groups", AVS_ADAM_MAX_GROUPS);
  AVS_REQUIRE(n_groups == 1 || group_chunks, "avs_adam_step: several groups need the per-chunk group map");
  AVS_REQUIRE(tick_state != nullptr || step >= 1, "avs_adam_step: step must be >= 1");
  if (n == 0) return 0;
  AdamGroups groups;
  for (int i = 0; i < AVS_ADAM_MAX_GROUPS; ++i) {
    groups.lr[i] = lr[min(i, n_groups - 1)];
    groups.wd[i] = weight_decay[min(i, n_groups - 1)];
  }
  float bc1 = 1.f, bc2s = 1.f;
  if (tick_state != nullptr) {
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((AdamTick*)tick_state, found_inf, beta1, beta2);
    if (int rc = avs_check_launch("adam_tick_kernel")) return rc;
  } else {
    bc1 = (float)(1.0 - pow((double)beta1, step));
    bc2s = (float)sqrt(1.0 - pow((double)beta2, step));
  }
  const int blocks = (int)min((long long)avs_num_sms() * 16, ceil_div_ll(n / 4, 256));
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)shadow_bf16, n, groups, beta1, beta2, eps, bc1,
                                                        bc2s, decoupled, inv_scale, found_inf, group_chunks,
                                                        (const AdamTick*)tick_state);
  return avs_check_launch("adam_kernel");
}

extern "C" int avs_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, int decoupled,
                             const float* inv_scale, const float* found_inf, const uint8_t* active_chunks,
                             void* stream) {
  return avs_adam_step_groups(p, g, m, v, shadow_bf16, n, &lr, &weight_decay, 1, beta1, beta2, eps, step, decoupled,
                              inv_scale, found_inf, active_chunks, nullptr, stream);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long n8 = n / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  for (long long i = n8 * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

extern "C" int avs_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  AVS_REQUIRE(src && dst, "avs_cast_f32_to_bf16: null pointer");
  AVS_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "avs_cast_f32_to_bf16: 16-byte alignment");
  if (n == 0) return 0;
  const int blocks = (int)min((long long)avs_num_sms() * 16, ceil_div_ll(n / 8 + 1, 256));
  cast_f32_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  return avs_check_launch("cast_f32_bf16_kernel");
}

// out[n] += alpha * sum_m dy[m, n]  (bias gradients of fc1 / qkv).  HBM-bound: each CTA owns 256 columns x a strip of
// rows; 8 warps stride the rows, every lane reads 16 bytes (8 bf16) per row, so a warp moves 512 contiguous bytes per
// load instruction and keeps 4 rows in flight; smem reduce; one fp32 atomic per column per CTA.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ out, int M,
                                                     int N, long long ld, int rows_per_cta, float alpha) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < N) {
    const bf16* p = dy + c0;
    int r = r0 + warp;
    for (; r + 24 < r1; r += 32) {   // 4 independent 16-byte loads in flight per lane
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(p + (size_t)(r + 8 * k) * ld));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f;
        f = unpack_bf16x2(u[k].x); a[0] += f.x; a[1] += f.y;
        f = unpack_bf16x2(u[k].y); a[2] += f.x; a[3] += f.y;
        f = unpack_bf16x2(u[k].z); a[4] += f.x; a[5] += f.y;
        f = unpack_bf16x2(u[k].w); a[6] += f.x; a[7] += f.y;
      }
    }
    for (; r < r1; r += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + (size_t)r * ld));
      float2 f;
      f = unpack_bf16x2(u.x); a[0] += f.x; a[1] += f.y;
      f = unpack_bf16x2(u.y); a[2] += f.x; a[3] += f.y;
      f = unpack_bf16x2(u.z); a[4] += f.x; a[5] += f.y;
      f = unpack_bf16x2(u.w); a[6] += f.x; a[7] += f.y;
    }
  }
  __shared__ float s[8][256];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[warp][lane * 8 + j] = a[j];
  __syncthreads();
  {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s[w][threadIdx.x];
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < N) atomicAdd(out + c, t * alpha);
  }
}

extern "C" int avs_colsum_bf16(const void* dy, long long ld, float* out, int M, int N, float alpha, void* stream) {
  AVS_REQUIRE(dy && out, "avs_colsum_bf16: null pointer");
  AVS_REQUIRE(N % 8 == 0 && ld % 8 == 0 && ((uintptr_t)dy & 15) == 0,
              "avs_colsum_bf16: N and ld must be multiples of 8 and dy 16-byte aligned");
  if (M == 0 || N == 0) return 0;
  const int col_blocks = ceil_div(N, 256);
  int row_blocks = max(1, (avs_num_sms() * 6) / col_blocks);
  int rows_per_cta = max(64, ceil_div(M, row_blocks));
  row_blocks = ceil_div(M, rows_per_cta);
  dim3 grid(col_blocks, row_blocks);
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, out, M, N, ld, rows_per_cta, alpha);
  return avs_check_launch("colsum_kernel");
}

// flag = 1.0 if any element is inf/nan (GradScaler found_inf); caller zeroes the flag first.
__global__ void found_inf_kernel(const float* __restrict__ g, long long n, float* __restrict__ flag) {
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    bad |= !isfinite(g[i]);
  if (__syncthreads_or(bad) && threadIdx.x == 0) *flag = 1.0f;
}

extern "C" int avs_found_inf(const float* g, long long n, float* flag, void* stream) {
  AVS_REQUIRE(g && flag, "avs_found_inf: null pointer");
  if (n == 0) return 0;
  const int blocks = (int)min((long long)avs_num_sms() * 16, ceil_div_ll(n, 256));
  found_inf_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, n, flag);
  return avs_check_launch("found_inf_kernel");
}
