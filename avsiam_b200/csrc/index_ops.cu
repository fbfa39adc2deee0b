// Bit-exact index kernels: random-mask argsort, kept-row gather, patch extraction (im2col-free, fused with the
// kept-token gather), decoder restore (mask-token fill + un-shuffle + pos/modality add) and its backward.
// All are HBM-bound: 128-bit accesses, one pass over the data.
#include "../../include/avsiam_b200.h"
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// avs_mask_argsort: per-row stable ascending argsort of noise[N,L] (ties -> lower index first) producing
// ids_shuffle, ids_restore (its inverse) and the binary mask (0 keep / 1 remove).
// Replaces the two torch.argsort + gather at cav_mae_base.py:377-388 / :426-437.
// One CTA per row; bitonic sort of (key,index) pairs in shared memory (L <= 2048).
// ---------------------------------------------------------------------------------------------------------
__global__ void mask_argsort_kernel(const float* __restrict__ noise, int L, int Lp2, int len_keep,
                                    int* __restrict__ ids_shuffle, int* __restrict__ ids_restore,
                                    float* __restrict__ mask) {
  extern __shared__ unsigned long long skeys[];  // (orderable float bits << 32) | index
  const int row = blockIdx.x;
  const float* nrow = noise + (size_t)row * L;
  for (int i = threadIdx.x; i < Lp2; i += blockDim.x) {
    unsigned long long key;
    if (i < L) {
      unsigned int u = __float_as_uint(nrow[i]);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // total order on floats
      key = ((unsigned long long)u << 32) | (unsigned int)i;
    } else {
      key = ~0ull;
    }
    skeys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= Lp2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < Lp2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = skeys[i], b = skeys[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) {
            skeys[i] = b;
            skeys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const int src = (int)(skeys[i] & 0xffffffffu);
    ids_shuffle[(size_t)row * L + i] = src;
    ids_restore[(size_t)row * L + src] = i;
    mask[(size_t)row * L + src] = (i >= len_keep) ? 1.0f : 0.0f;
  }
}

extern "C" int avs_mask_argsort(const float* noise, int N, int L, int len_keep, int32_t* ids_shuffle,
                                int32_t* ids_restore, float* mask, void* stream) {
  AVS_REQUIRE(noise && ids_shuffle && ids_restore && mask, "avs_mask_argsort: null pointer");
  AVS_REQUIRE(N >= 0 && L > 0 && L <= 2048 && len_keep >= 0 && len_keep <= L, "avs_mask_argsort: bad shape N=%d L=%d",
              N, L);
  if (N == 0) return 0;
  int Lp2 = 1;
  while (Lp2 < L) Lp2 <<= 1;
  mask_argsort_kernel<<<N, 256, Lp2 * sizeof(unsigned long long), (cudaStream_t)stream>>>(
      noise, L, Lp2, len_keep, ids_shuffle, ids_restore, mask);
  return avs_check_launch("mask_argsort_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// avs_mask_force_noise: the structured-mask pattern of random_masking_structured (cav_mae_base.py:404-423): the noise
// of whole time columns / frequency rows of the [f, t] patch grid is overwritten with `value` (1.1: "large value will
// be removed") before the argsort. The column / row lists are the host's random.sample draws (the reference draws
// them with Python's `random`, so the draws stay on the host to remain seed-compatible); one launch applies all of
// them for the whole batch instead of the reference's N x k Python slice writes.
//   cols int32 [N, kt] (time indices, may be NULL / kt = 0), rows int32 [N, kf] (frequency indices, may be NULL).
// ---------------------------------------------------------------------------------------------------------
__global__ void mask_force_noise_kernel(float* __restrict__ noise, const int* __restrict__ cols, int kt,
                                        const int* __restrict__ rows, int kf, int f, int t, float value) {
  float* nz = noise + (size_t)blockIdx.x * f * t;
  const int* c = cols ? cols + (size_t)blockIdx.x * kt : nullptr;
  const int* r = rows ? rows + (size_t)blockIdx.x * kf : nullptr;
  for (int i = threadIdx.x; i < kt * f; i += blockDim.x) nz[(i / kt) * t + c[i % kt]] = value;
  for (int i = threadIdx.x; i < kf * t; i += blockDim.x) nz[r[i / t] * t + (i % t)] = value;
}

extern "C" int avs_mask_force_noise(float* noise, int N, int f, int t, const int32_t* cols, int kt,
                                    const int32_t* rows, int kf, float value, void* stream) {
  AVS_REQUIRE(noise, "avs_mask_force_noise: null pointer");
  AVS_REQUIRE(N >= 0 && f > 0 && t > 0 && kt >= 0 && kf >= 0 && kt <= t && kf <= f,
              "avs_mask_force_noise: bad shape N=%d f=%d t=%d kt=%d kf=%d", N, f, t, kt, kf);
  AVS_REQUIRE((kt == 0 || cols) && (kf == 0 || rows), "avs_mask_force_noise: index list missing");
  if (N == 0 || (kt == 0 && kf == 0)) return 0;
  mask_force_noise_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(noise, kt ? cols : nullptr, kt, kf ? rows : nullptr, kf,
                                                              f, t, value);
  return avs_check_launch("mask_force_noise_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// avs_gather_rows: out[n, i, :] = x[n, ids[n, i], :] for i < keep — byte-exact row copies (16-byte vectors).
// Replaces torch.gather(x, 1, ids_keep.unsqueeze(-1).repeat(1,1,D)) at cav_mae_base.py:382,431.
// ---------------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const uint4* __restrict__ x, const int* __restrict__ ids, uint4* __restrict__ out,
                                   int L, int keep, int ids_ld, int row_vec, long long total_vec) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total_vec;
       g += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(g % row_vec);
    const long long r = g / row_vec;  // output row = n*keep + i
    const int n = (int)(r / keep), i = (int)(r % keep);
    const int src = ids[(size_t)n * ids_ld + i];
    out[g] = x[((size_t)n * L + src) * row_vec + c];
  }
}

extern "C" int avs_gather_rows(const void* x, const int32_t* ids, void* out, int N, int L, int keep, int ids_ld,
                               int row_bytes, void* stream) {
  if (N == 0 || keep == 0) return 0;  // empty selection: nothing to copy
  AVS_REQUIRE(x && ids && out, "avs_gather_rows: null pointer");
  AVS_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0, "avs_gather_rows: row_bytes must be a multiple of 16");
  AVS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "avs_gather_rows: 16-byte alignment");
  AVS_REQUIRE(keep >= 0 && keep <= L && ids_ld >= keep, "avs_gather_rows: bad keep/ids_ld");
  const long long total = (long long)N * keep * (row_bytes / 16);
  if (total == 0) return 0;
  const int blocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll(total, 256));
  gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, ids, (uint4*)out, L, keep, ids_ld,
                                                               row_bytes / 16, total);
  return avs_check_launch("gather_rows_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// Patch extraction fused with the kept-token gather (only kept patches are ever materialised):
//   audio [B, T, F] fp32 -> rows [B*keep, p*p] bf16, token = f*ta + t, vector order (pf, pt)
//     (PatchEmbed on a.unsqueeze(1).transpose(2,3), cav_mae_base.py:444-448; SURVEY §8a identity 3)
//   video [B, C, H, W] fp32 -> rows [B*keep, C*p*p] bf16, token = h*g + w, vector order (c, p, q)  (identity 4)
// ids == NULL => all tokens in natural order (keep == number of tokens).
// sample_idx (int32 [B], optional): output sample b reads input sample sample_idx[b] (the chunk permutation of
// forward_encoder_mmixed, cav_mae_base.py:533-549) — B is the number of OUTPUT samples.
// ---------------------------------------------------------------------------------------------------------
__global__ void patchify_audio_kernel(const float* __restrict__ audio, const int* __restrict__ ids,
                                      const int* __restrict__ sample_idx, bf16* __restrict__ out, int T, int F, int p, int ta, int keep, int ids_ld,
                                      int ld_out, long long total8) {
  const int vec = p * p;  // elements per patch
  const int per_row = vec / 8;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total8;
       g += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(g % per_row);
    const long long r = g / per_row;
    const int bl = (int)(r / keep), i = (int)(r % keep);
    const int tok = ids ? ids[(size_t)bl * ids_ld + i] : i;
    const int b = sample_idx ? sample_idx[bl] : bl;
    const int f = tok / ta, t = tok % ta;
    const int e0 = c8 * 8;
    const int pf = e0 / p, pt0 = e0 % p;  // p is a multiple of 8 => the 8 elements share pf
    const float* src = audio + ((size_t)b * T + (size_t)t * p + pt0) * F + f * p + pf;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (size_t)j * F);
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + (size_t)r * ld_out + e0) = o;
  }
}

__global__ void patchify_video_kernel(const float* __restrict__ img, const int* __restrict__ ids,
                                      const int* __restrict__ sample_idx, bf16* __restrict__ out, int C, int H, int W, int p, int gw, int keep,
                                      int ids_ld, int ld_out, long long total8) {
  const int vec = C * p * p;
  const int per_row = vec / 8;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total8;
       g += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(g % per_row);
    const long long r = g / per_row;
    const int bl = (int)(r / keep), i = (int)(r % keep);
    const int tok = ids ? ids[(size_t)bl * ids_ld + i] : i;
    const int b = sample_idx ? sample_idx[bl] : bl;
    const int h = tok / gw, w = tok % gw;
    const int e0 = c8 * 8;
    const int c = e0 / (p * p), pp = (e0 / p) % p, q0 = e0 % p;
    const float* src = img + (((size_t)b * C + c) * H + (size_t)h * p + pp) * W + (size_t)w * p + q0;
    float v[8];
    if ((((uintptr_t)src) & 15) == 0) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldg(src + j);
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + (size_t)r * ld_out + e0) = o;
  }
}

// Generic patch geometry (ViT-H/14: patch 14 is not a multiple of 8 and does not divide 1024 x 128 — the stride-14
// conv of PatchEmbed simply never reads the last 2 rows / columns): one thread per 8 output columns
// of the PADDED row [0, ld_out), every element addressed on its own; columns >= the patch vector length are written
// as zeros, so the row can be the K-padded A operand of the patch-embed GEMM (TMA wants 16-byte row pitches:
// 196 -> 200, 588 -> 592 columns).  kind 0: audio [B, T, F], kind 1: video [B, C, H, W].
__global__ void patchify_generic_kernel(const float* __restrict__ in, const int* __restrict__ ids,
                                        const int* __restrict__ sample_idx, bf16* __restrict__ out, int kind, int C, int d0,
                                        int d1, int p, int gw, int keep, int ids_ld, int ld_out, long long total8) {
  const int vec = C * p * p;
  const int per_row = ld_out / 8;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total8;
       g += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(g % per_row);
    const long long r = g / per_row;
    const int bl = (int)(r / keep), i = (int)(r % keep);
    const int tok = ids ? ids[(size_t)bl * ids_ld + i] : i;
    const int b = sample_idx ? sample_idx[bl] : bl;
    const int tr = tok / gw, tc = tok % gw;   // audio: (f, t)   video: (h, w)
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = c8 * 8 + j;
      if (e >= vec) { v[j] = 0.f; continue; }
      if (kind == 0) {
        const int pf = e / p, pt = e % p;
        v[j] = __ldg(in + ((size_t)b * d0 + (size_t)tc * p + pt) * d1 + tr * p + pf);
      } else {
        const int c = e / (p * p), pp = (e / p) % p, q = e % p;
        v[j] = __ldg(in + (((size_t)b * C + c) * d0 + (size_t)tr * p + pp) * d1 + (size_t)tc * p + q);
      }
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + (size_t)r * ld_out + c8 * 8) = o;
  }
}

static int launch_patchify_generic(const float* in, const int32_t* ids, const int32_t* sample_idx, void* out, int kind, int B,
                                   int C, int d0, int d1, int patch, int gw, int keep, int ids_ld, int ld_out, void* stream) {
  const long long total8 = (long long)B * keep * (ld_out / 8);
  if (total8 == 0) return 0;
  const int blocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll(total8, 256));
  patchify_generic_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, ids, sample_idx, (bf16*)out, kind, C, d0, d1, patch, gw,
                                                                    keep, ids_ld, ld_out, total8);
  return avs_check_launch("patchify_generic_kernel");
}

extern "C" int avs_patchify_audio(const float* audio, const int32_t* ids, const int32_t* sample_idx, void* out, int B, int T, int F, int patch,
                                  int keep, int ids_ld, int ld_out, void* stream) {
  AVS_REQUIRE(audio && out, "avs_patchify_audio: null pointer");
  AVS_REQUIRE(patch > 0 && T >= patch && F >= patch, "avs_patchify_audio: bad patch size");
  AVS_REQUIRE(ld_out >= patch * patch && ld_out % 8 == 0 && ((uintptr_t)out & 15) == 0, "avs_patchify_audio: bad ld_out/alignment");
  const int ntok = (T / patch) * (F / patch);   // floor, as Conv2d(kernel = stride = patch) does
  AVS_REQUIRE(keep > 0 && keep <= ntok && (ids != nullptr || keep == ntok), "avs_patchify_audio: bad keep");
  if (patch % 8 != 0)
    return launch_patchify_generic(audio, ids, sample_idx, out, 0, B, 1, T, F, patch, T / patch, keep, ids_ld, ld_out, stream);
  const long long total8 = (long long)B * keep * (patch * patch / 8);
  if (total8 == 0) return 0;
  const int blocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll(total8, 256));
  patchify_audio_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(audio, ids, sample_idx, (bf16*)out, T, F, patch, T / patch, keep,
                                                                  ids_ld, ld_out, total8);
  return avs_check_launch("patchify_audio_kernel");
}

extern "C" int avs_patchify_video(const float* img, const int32_t* ids, const int32_t* sample_idx, void* out, int B, int C, int H, int W,
                                  int patch, int keep, int ids_ld, int ld_out, void* stream) {
  AVS_REQUIRE(img && out, "avs_patchify_video: null pointer");
  AVS_REQUIRE(patch > 0 && H >= patch && W >= patch, "avs_patchify_video: bad patch size");
  AVS_REQUIRE(ld_out >= C * patch * patch && ld_out % 8 == 0 && ((uintptr_t)out & 15) == 0, "avs_patchify_video: bad ld_out/alignment");
  const int ntok = (H / patch) * (W / patch);
  AVS_REQUIRE(keep > 0 && keep <= ntok && (ids != nullptr || keep == ntok), "avs_patchify_video: bad keep");
  if (patch % 8 != 0)
    return launch_patchify_generic(img, ids, sample_idx, out, 1, B, C, H, W, patch, W / patch, keep, ids_ld, ld_out, stream);
  const long long total8 = (long long)B * keep * (C * patch * patch / 8);
  if (total8 == 0) return 0;
  const int blocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll(total8, 256));
  patchify_video_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(img, ids, sample_idx, (bf16*)out, C, H, W, patch, W / patch, keep,
                                                                  ids_ld, ld_out, total8);
  return avs_check_launch("patchify_video_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// Decoder restore (cav_mae_base.py:604-626), one launch, no host sync:
//   out[b, j]      = (ira[b,j] < ka ? x[b, ira[b,j]]      : mask_token) + pos_a[j] + mod_a      j in [0,Ta)
//   out[b, Ta + j] = (irv[b,j] < kv ? x[b, ka + irv[b,j]] : mask_token) + pos_v[j] + mod_v      j in [0,Tv)
// x: bf16 [B, ka+kv, D] (decoder_embed output), out: bf16 [B, Ta+Tv, D]; fp32 parameters.
// ---------------------------------------------------------------------------------------------------------
// one warp per output row: the row's source / position decode happens once per warp in 32-bit arithmetic (the first
// version spent its time in 64-bit divisions per 16-byte chunk: 0.93 TB/s), lanes stride over the 16-byte chunks
__global__ void __launch_bounds__(256) decoder_restore_fwd_kernel(
    const bf16* __restrict__ x, const int* __restrict__ ira, const int* __restrict__ irv,
    const float* __restrict__ mask_token, const float* __restrict__ pos_a, const float* __restrict__ pos_v,
    const float* __restrict__ mod_a, const float* __restrict__ mod_v, bf16* __restrict__ out, int Ta, int Tv, int ka,
    int kv, int D, int rows) {
  const int S = Ta + Tv, K = ka + kv;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int b = r / S, j = r - b * S;
    const bool is_a = j < Ta;
    const int jj = is_a ? j : j - Ta;
    const int src = is_a ? ira[(size_t)b * Ta + jj] : irv[(size_t)b * Tv + jj];
    const bool kept = src < (is_a ? ka : kv);
    const bf16* xr = x + ((size_t)b * K + (is_a ? 0 : ka) + (kept ? src : 0)) * D;
    const float* pos = (is_a ? pos_a : pos_v) + (size_t)jj * D;
    const float* mod = is_a ? mod_a : mod_v;
    bf16* orow = out + (size_t)r * D;
    for (int c = lane * 8; c < D; c += 256) {
      float v[8];
      if (kept) {
        const uint4 u = *reinterpret_cast<const uint4*>(xr + c);
        float2 f;
        f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
      } else {
        const float4 m0 = *reinterpret_cast<const float4*>(mask_token + c);
        const float4 m1 = *reinterpret_cast<const float4*>(mask_token + c + 4);
        v[0] = m0.x; v[1] = m0.y; v[2] = m0.z; v[3] = m0.w; v[4] = m1.x; v[5] = m1.y; v[6] = m1.z; v[7] = m1.w;
      }
      const float4 p0 = *reinterpret_cast<const float4*>(pos + c), p1 = *reinterpret_cast<const float4*>(pos + c + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(mod + c), d1 = *reinterpret_cast<const float4*>(mod + c + 4);
      // (v + (pos + mod)) in the order of the first version: results stay bit-identical
      v[0] += p0.x + d0.x; v[1] += p0.y + d0.y; v[2] += p0.z + d0.z; v[3] += p0.w + d0.w;
      v[4] += p1.x + d1.x; v[5] += p1.y + d1.y; v[6] += p1.z + d1.z; v[7] += p1.w + d1.w;
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(orow + c) = o;
    }
  }
}

extern "C" int avs_decoder_restore_fwd(const void* x, const int32_t* ids_restore_a, const int32_t* ids_restore_v,
                                       const float* mask_token, const float* pos_a, const float* pos_v,
                                       const float* mod_a, const float* mod_v, void* out, int B, int Ta, int Tv,
                                       int keep_a, int keep_v, int D, void* stream) {
  AVS_REQUIRE(x && ids_restore_a && ids_restore_v && mask_token && pos_a && pos_v && mod_a && mod_v && out,
              "avs_decoder_restore_fwd: null pointer");
  AVS_REQUIRE(D % 8 == 0, "avs_decoder_restore_fwd: D must be a multiple of 8");
  AVS_REQUIRE((long long)B * (Ta + Tv) < (1ll << 31), "avs_decoder_restore_fwd: too many rows");
  AVS_REQUIRE(((uintptr_t)mask_token & 15) == 0 && ((uintptr_t)pos_a & 15) == 0 && ((uintptr_t)pos_v & 15) == 0 &&
                  ((uintptr_t)mod_a & 15) == 0 && ((uintptr_t)mod_v & 15) == 0,
              "avs_decoder_restore_fwd: fp32 parameters must be 16-byte aligned");
  const int rows = B * (Ta + Tv);
  if (rows == 0) return 0;
  const int blocks = min(avs_num_sms() * 8, ceil_div(rows, 8));
  decoder_restore_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, ids_restore_a, ids_restore_v, mask_token, pos_a, pos_v, mod_a, mod_v, (bf16*)out, Ta, Tv, keep_a,
      keep_v, D, rows);
  return avs_check_launch("decoder_restore_fwd_kernel");
}

// Backward of the restore. One CTA per position j (grid = Ta+Tv), looping over the batch:
//   dx[b, src]   = dout[b, j]                     where src = ids_restore[b,j] < keep     (pure scatter, unique)
//   dpos[j]     += sum_b dout[b,j]                (fp32 atomics, one per batch slice)
//   dmask_token += sum over masked (b,j)          (fp32 atomics, one per CTA per column)
//   dmod_{a,v}  += sum over (b,j) of the modality (fp32 atomics)
// grid (position j, batch slice): every thread owns 8 columns (16-byte loads / stores) and walks its slice of the batch
// with 4 independent loads in flight; the per-position sums of the slices are combined with fp32 atomics (the first
// version walked the whole batch serially with 4-byte accesses).
constexpr int RB_SLICES = 4;
__global__ void __launch_bounds__(64) decoder_restore_bwd_kernel(
    const bf16* __restrict__ dout, const int* __restrict__ ira, const int* __restrict__ irv, bf16* __restrict__ dx,
    float* __restrict__ dmask_token, float* __restrict__ dpos_a, float* __restrict__ dpos_v,
    float* __restrict__ dmod_a, float* __restrict__ dmod_v, int B, int Ta, int Tv, int ka, int kv, int D) {
  const int j = blockIdx.x;
  const int S = Ta + Tv, K = ka + kv;
  const bool is_a = j < Ta;
  const int jj = is_a ? j : j - Ta;
  const int keep = is_a ? ka : kv;
  const int* ir = is_a ? ira + jj : irv + jj;
  const int irs = is_a ? Ta : Tv;
  const int per = (B + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(B, b0 + per);
  for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
    float sp[8], sm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sp[i] = sm[i] = 0.f;
    for (int b = b0; b < b1; b += 4) {
      uint4 u[4];
      int src[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = b + k < b1;
        u[k] = ok ? *reinterpret_cast<const uint4*>(dout + ((size_t)(b + k) * S + j) * D + c) : make_uint4(0, 0, 0, 0);
        src[k] = ok ? ir[(size_t)(b + k) * irs] : keep;          // out-of-range rows count as "masked" zeros
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v[8];
        float2 f;
        f = unpack_bf16x2(u[k].x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(u[k].y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(u[k].z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(u[k].w); v[6] = f.x; v[7] = f.y;
#pragma unroll
        for (int i = 0; i < 8; ++i) sp[i] += v[i];
        if (src[k] < keep) {
          *reinterpret_cast<uint4*>(dx + ((size_t)(b + k) * K + (is_a ? 0 : ka) + src[k]) * D + c) = u[k];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) sm[i] += v[i];
        }
      }
    }
    float* dpos = (is_a ? dpos_a : dpos_v) + (size_t)jj * D + c;
    float* dmod = (is_a ? dmod_a : dmod_v) + c;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(dpos + i, sp[i]);
      atomicAdd(dmod + i, sp[i]);
      atomicAdd(dmask_token + c + i, sm[i]);
    }
  }
}

extern "C" int avs_decoder_restore_bwd(const void* dout, const int32_t* ids_restore_a, const int32_t* ids_restore_v,
                                       void* dx, float* dmask_token, float* dpos_a, float* dpos_v, float* dmod_a,
                                       float* dmod_v, int B, int Ta, int Tv, int keep_a, int keep_v, int D,
                                       void* stream) {
  AVS_REQUIRE(dout && ids_restore_a && ids_restore_v && dx && dmask_token && dpos_a && dpos_v && dmod_a && dmod_v,
              "avs_decoder_restore_bwd: null pointer");
  AVS_REQUIRE(D % 8 == 0, "avs_decoder_restore_bwd: D must be a multiple of 8");
  if (B == 0 || Ta + Tv == 0) return 0;
  decoder_restore_bwd_kernel<<<dim3(Ta + Tv, RB_SLICES), 64, 0, (cudaStream_t)stream>>>(
      (const bf16*)dout, ids_restore_a, ids_restore_v, (bf16*)dx, dmask_token, dpos_a, dpos_v, dmod_a, dmod_v, B, Ta,
      Tv, keep_a, keep_v, D);
  return avs_check_launch("decoder_restore_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// avs_scatter_add_rows: table[idx[m], :] += alpha * dy[m, :]   (fp32 red.global.add; pos-embed gradient of the
// fused patch-embed epilogue, cav_mae_base.py:449-450,454-455).   idx == NULL => m % table_rows.
// ---------------------------------------------------------------------------------------------------------
__global__ void scatter_add_rows_kernel(const bf16* __restrict__ dy, const int* __restrict__ idx,
                                        float* __restrict__ table, int D, int table_rows, float alpha,
                                        long long total4) {
  const int per_row = D / 4;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total4;
       g += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(g % per_row) * 4;
    const long long m = g / per_row;
    const int t = idx ? idx[m] : (int)(m % table_rows);
    const uint2 u = *reinterpret_cast<const uint2*>(dy + (size_t)m * D + c);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    float* dst = table + (size_t)t * D + c;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dst), "f"(a.x * alpha), "f"(a.y * alpha),
                 "f"(b.x * alpha), "f"(b.y * alpha)
                 : "memory");
  }
}

extern "C" int avs_scatter_add_rows(const void* dy, const int32_t* idx, float* table, int M, int D, int table_rows,
                                    float alpha, void* stream) {
  AVS_REQUIRE(dy && table, "avs_scatter_add_rows: null pointer");
  AVS_REQUIRE(D % 4 == 0 && ((uintptr_t)table & 15) == 0, "avs_scatter_add_rows: D %% 4 and 16-byte table alignment");
  const long long total4 = (long long)M * (D / 4);
  if (total4 == 0) return 0;
  const int blocks = (int)min((long long)avs_num_sms() * 8, ceil_div_ll(total4, 256));
  scatter_add_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, idx, table, D, table_rows, alpha,
                                                                    total4);
  return avs_check_launch("scatter_add_rows_kernel");
}
