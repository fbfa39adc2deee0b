// Input pipeline on the GPU, second half: what AudiosetDataset.__getitem__ does to a decoded
// sample after the fbank / frame decode (src/dataloader.py:148-155,452-459,491-516), batched:
//
//   avs_fbank_augment        SpecAugment frequency / time masks (torchaudio FrequencyMasking / TimeMasking: a band
//                            [start, end) filled with 0.0), normalisation (x - mean) / std, uniform noise * r / 10,
//                            torch.roll along time.  The random draws (band limits, noise scale, shift, optionally the
//                            noise field) are SUPPLIED by the host wrapper, so results are a pure function of them.
//   avs_frames_preprocess    frames uint8 [N, C, H, W] -> /255 -> Resize([oh, ow], BICUBIC, antialias=True) ->
//                            Normalize(mean, std) (-> optional mixup with a second frame set), fp32 [N, C, oh, ow].
//                            Separable antialiased bicubic (a = -0.5, PIL / ATen _upsample_bicubic2d_aa): horizontal
//                            then vertical pass fused in one kernel through a shared-memory tile; per-output tap
//                            windows and normalised weights are prepared by the host wrapper exactly as ATen computes
//                            them.
// Elementwise fp32 with IEEE division / unfused multiply-add where the reference's op order is observable, so the
// audio path is bit-exact against torch given the same draws.  HBM-bound.
#include <stdint.h>

#include "../../include/avsiam_b200.h"
#include "common.cuh"

namespace {

struct AugParams {  // per sample, int32 x 6 + float: [f0, f1, t0, t1, shift, _pad], scale
  int f0, f1, t0, t1, shift, pad;
};

// one warp per (b, t) row, lanes stride over float4 chunks (F = 128: one chunk per lane); 32-bit index math
__global__ void fbank_augment_kernel(const float* __restrict__ x, const int* __restrict__ prm,
                                     const float* __restrict__ scale, const float* __restrict__ noise,
                                     float* __restrict__ out, int rows, int T, int F, float mean, float stdv,
                                     int skip_norm) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const int b = row / T, t = row - b * T;
    const AugParams p = *reinterpret_cast<const AugParams*>(prm + b * 6);
    const bool tmask = t >= p.t0 && t < p.t1;
    int to = (t + p.shift) % T;                         // torch.roll: out[(t + shift) mod T] = in[t]
    if (to < 0) to += T;
    const float s = noise ? scale[b] : 0.f;
    const float* xr = x + (long long)row * F;
    const float* nr = noise ? noise + (long long)row * F : nullptr;
    float* orow = out + ((long long)b * T + to) * F;
    for (int c = lane * 4; c < F; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (tmask || (c + k >= p.f0 && c + k < p.f1)) e[k] = 0.f;
        if (!skip_norm) e[k] = __fdiv_rn(__fsub_rn(e[k], mean), stdv);
      }
      if (nr) {
        const float4 nz = *reinterpret_cast<const float4*>(nr + c);
        // `fbank + torch.rand(T, F) * np.random.rand() / 10` (:512) evaluates as fbank + ((noise * r) / 10)
        e[0] = __fadd_rn(e[0], __fdiv_rn(__fmul_rn(nz.x, s), 10.f));
        e[1] = __fadd_rn(e[1], __fdiv_rn(__fmul_rn(nz.y, s), 10.f));
        e[2] = __fadd_rn(e[2], __fdiv_rn(__fmul_rn(nz.z, s), 10.f));
        e[3] = __fadd_rn(e[3], __fdiv_rn(__fmul_rn(nz.w, s), 10.f));
      }
      *reinterpret_cast<float4*>(orow + c) = make_float4(e[0], e[1], e[2], e[3]);
    }
  }
}

// Fused separable resize: one CTA produces `ty` output rows x all `ow` columns of one (n, c) plane.
//   1. the input rows its vertical windows touch ([ymin[first], ymin[last] + ysize[last]), at most `span`) are staged
//      RB at a time into shared memory as floats through a 256-entry table of i / 255 (IEEE division, once per CTA),
//   2. horizontal pass out of shared memory (weights transposed into shared memory once, each weight reused for the
//      RB rows of a batch) into the fp32 tile tmp[span][ow], which never leaves the SM,
//   3. vertical pass + Normalize out of tmp, coalesced stores.
// HBM traffic = the uint8 frame (plus the window overlap between neighbouring tiles, served by L2) + the fp32 output.
constexpr int RS_RB = 4;
__global__ void __launch_bounds__(256) resize_fused_kernel(
    const uint8_t* __restrict__ in, const float* __restrict__ wx, const int* __restrict__ xmin,
    const int* __restrict__ xsize, int taps_x, const float* __restrict__ wy, const int* __restrict__ ymin,
    const int* __restrict__ ysize, int taps_y, int C, int H, int W, int oh, int ow, const float* __restrict__ mean,
    const float* __restrict__ stdv, float* __restrict__ out, int ty, int span) {
  extern __shared__ __align__(16) float rs_smem[];
  float* lut = rs_smem;                         // [256]
  float* wxs = lut + 256;                       // [taps_x][ow]
  float* tmp = wxs + taps_x * ow;               // [span][ow]
  float* rowf = tmp + span * ow;                // [RS_RB][W]
  const int plane = blockIdx.y, yo0 = blockIdx.x * ty, yo1 = min(yo0 + ty, oh);
  lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.f);
  for (int i = threadIdx.x; i < taps_x * ow; i += 256) {
    const int xo = i / taps_x, j = i - xo * taps_x;
    wxs[j * ow + xo] = wx[i];
  }
  const int yA = ymin[yo0], yB = ymin[yo1 - 1] + ysize[yo1 - 1];
  const uint8_t* src = in + ((long long)plane * H + yA) * W;
  const bool vec = (W & 3) == 0 && (((taps_x + span) * ow) & 3) == 0;   // 4-byte aligned rows, 16-byte aligned rowf
  for (int rb = 0; rb < yB - yA; rb += RS_RB) {
    const int nr = min(RS_RB, yB - yA - rb);
    __syncthreads();                            // lut / wxs ready; previous batch consumed
    if (vec) {
      const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + (long long)rb * W);
      for (int i = threadIdx.x; i < (nr * W) >> 2; i += 256) {
        const uint32_t v = s4[i];
        *reinterpret_cast<float4*>(rowf + 4 * i) =
            make_float4(lut[v & 255u], lut[(v >> 8) & 255u], lut[(v >> 16) & 255u], lut[v >> 24]);
      }
    } else {
      for (int i = threadIdx.x; i < nr * W; i += 256) rowf[i] = lut[src[(long long)rb * W + i]];
    }
    __syncthreads();
    for (int xo = threadIdx.x; xo < ow; xo += 256) {
      const float* r0 = rowf + xmin[xo];
      const int n = xsize[xo];
      float acc[RS_RB];
      {
        const float w = wxs[xo];
#pragma unroll
        for (int r = 0; r < RS_RB; ++r) acc[r] = __fmul_rn(r0[r * W], w);      // rows >= nr: stale smem, discarded
      }
      for (int j = 1; j < n; ++j) {
        const float w = wxs[j * ow + xo];
#pragma unroll
        for (int r = 0; r < RS_RB; ++r) acc[r] = fmaf(r0[r * W + j], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < RS_RB; ++r)
        if (r < nr) tmp[(rb + r) * ow + xo] = acc[r];
    }
  }
  __syncthreads();
  const int c = plane % C;
  const float m = mean[c], sd = stdv[c];
  for (int i = threadIdx.x; i < (yo1 - yo0) * ow; i += 256) {
    const int dy = i / ow, xo = i - dy * ow, yo = yo0 + dy;
    const float* w = wy + yo * taps_y;
    const float* t0 = tmp + (ymin[yo] - yA) * ow + xo;
    const int n = ysize[yo];
    float t = __fmul_rn(t0[0], w[0]);
    for (int j = 1; j < n; ++j) t = fmaf(t0[j * ow], w[j], t);
    out[((long long)plane * oh + yo) * ow + xo] = __fdiv_rn(__fsub_rn(t, m), sd);
  }
}

// image = weight * image + (1 - weight) * image2   (dataloader.py:419-420), weight per sample
__global__ void mix_frames_kernel(float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ weight,
                                  long long per_sample, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float w = weight[i / per_sample];
    a[i] = __fadd_rn(__fmul_rn(w, a[i]), __fmul_rn(__fsub_rn(1.f, w), b[i]));
  }
}

inline int grid_for(long long work, int threads) {
  return (int)min((long long)avs_num_sms() * 16, ceil_div_ll(work, threads));
}

}  // namespace

extern "C" int avs_fbank_augment(const float* fbank, const int* params, const float* noise_scale, const float* noise,
                                 float* out, int B, int T, int F, float norm_mean, float norm_std, int skip_norm,
                                 void* stream) {
  AVS_REQUIRE(fbank && params && out, "avs_fbank_augment: null pointer");
  AVS_REQUIRE(fbank != out, "avs_fbank_augment: the roll cannot run in place");
  AVS_REQUIRE(B >= 0 && T > 0 && F > 0 && F % 4 == 0, "avs_fbank_augment: bad shape (F must be a multiple of 4)");
  AVS_REQUIRE(!noise || noise_scale, "avs_fbank_augment: noise needs noise_scale");
  if (B == 0) return 0;
  AVS_REQUIRE((long long)B * T < (1ll << 31), "avs_fbank_augment: too many rows");
  const int rows = B * T;
  fbank_augment_kernel<<<min(avs_num_sms() * 16, ceil_div(rows, 8)), 256, 0, (cudaStream_t)stream>>>(
      fbank, params, noise_scale, noise, out, rows, T, F, norm_mean, norm_std, skip_norm);
  return avs_check_launch("fbank_augment_kernel");
}

extern "C" int avs_frames_preprocess(const unsigned char* frames, int N, int C, int H, int W, const float* wx,
                                     const int* xmin, const int* xsize, int taps_x, const float* wy, const int* ymin,
                                     const int* ysize, int taps_y, int out_h, int out_w, int tile_rows, int span_rows,
                                     const float* mean, const float* stdv, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AVS_REQUIRE(frames && wx && xmin && xsize && wy && ymin && ysize && mean && stdv && out,
              "avs_frames_preprocess: null pointer");
  AVS_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0 && taps_x > 0 && taps_y > 0 &&
                  tile_rows > 0 && span_rows > 0, "avs_frames_preprocess: bad shape");
  if (N == 0) return 0;
  AVS_REQUIRE((long long)N * C <= 65535, "avs_frames_preprocess: more than 65535 planes per call");
  const size_t smem = sizeof(float) * (256 + (size_t)taps_x * out_w + (size_t)span_rows * out_w + (size_t)RS_RB * W);
  AVS_REQUIRE(smem <= 200 * 1024, "avs_frames_preprocess: tile does not fit shared memory (smaller tile_rows)");
  static thread_local size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(resize_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    AVS_REQUIRE(e == cudaSuccess, "avs_frames_preprocess: cudaFuncSetAttribute failed");
    smem_set = smem;
  }
  resize_fused_kernel<<<dim3(ceil_div(out_h, tile_rows), N * C), 256, smem, stream>>>(
      frames, wx, xmin, xsize, taps_x, wy, ymin, ysize, taps_y, C, H, W, out_h, out_w, mean, stdv, out, tile_rows,
      span_rows);
  return avs_check_launch("resize_fused_kernel");
}

extern "C" int avs_mix_frames(float* image, const float* image2, const float* weight, int N, long long per_sample,
                              void* stream) {
  AVS_REQUIRE(image && image2 && weight, "avs_mix_frames: null pointer");
  AVS_REQUIRE(N >= 0 && per_sample > 0, "avs_mix_frames: bad shape");
  if (N == 0) return 0;
  const long long total = (long long)N * per_sample;
  mix_frames_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(image, image2, weight, per_sample, total);
  return avs_check_launch("mix_frames_kernel");
}
