// Audio front end on the GPU: kaldi-compatible log-mel filterbank of a batch of waveforms,
// fused with the data loader's mean removal, pad/crop to target_length and normalisation.
//
// Replaces (src/dataloader.py:287,323-339,506)
//   waveform = waveform - waveform.mean()
//   fbank = torchaudio.compliance.kaldi.fbank(waveform, htk_compat=True, sample_frequency=16000, use_energy=False,
//                                             window_type='hanning', num_mel_bins=128, dither=0.0, frame_shift=10)
//   zero-pad / crop to target_length;  fbank = (fbank - norm_mean) / norm_std
// which the reference runs on CPU loader workers, one clip at a time.
//
// One warp per frame: 400 samples (25 ms, shift 160) -> per-frame DC removal, pre-emphasis 0.97 (left sample
// replicated), Hann window (non-periodic), zero-pad to 512, radix-2 FFT in the warp's private shared-memory buffer
// (9 stages x 8 butterflies per lane, twiddles from a per-CTA table computed with sincospi in double), power
// spectrum, 128 triangular mel filters (dense fp32 weights [128, 257] + per-filter bin ranges prepared by the host
// wrapper exactly as kaldi's get_mel_banks defines them), log(max(., FLT_EPSILON)), normalise, coalesced store.
// fp32 throughout, like torchaudio.  HBM-bound in principle (0.64 KB in / 0.5 KB out per frame); ~25 kFLOP per frame.
#include <float.h>

#include "../../include/avsiam_b200.h"
#include "common.cuh"

namespace {

constexpr int FB_FRAME = 400, FB_SHIFT = 160, FB_NFFT = 512, FB_MEL = 128, FB_BINS = 257, FB_WARPS = 8;

// mean[b] = mean of wav[b, 0:L]   (one CTA per clip, double accumulation across the block)
__global__ void __launch_bounds__(256) fbank_mean_kernel(const float* __restrict__ wav, long long ld, int L,
                                                         float* __restrict__ mean) {
  const float* w = wav + (long long)blockIdx.x * ld;
  double acc = 0.0;
  for (int i = threadIdx.x; i < L; i += 256) acc += (double)w[i];
  __shared__ double s[256];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) mean[blockIdx.x] = (float)(s[0] / (double)L);
}

__global__ void __launch_bounds__(FB_WARPS * 32) fbank_kernel(
    const float* __restrict__ wav, long long ld, int L, const float* __restrict__ wav_mean,
    const float* __restrict__ melw, const short2* __restrict__ mel_range, float* __restrict__ out, int B,
    int target_len, int n_frames, float norm_mean, float inv_std) {
  __shared__ float2 tw[FB_NFFT / 2];
  __shared__ float re[FB_WARPS][FB_NFFT];
  __shared__ float im[FB_WARPS][FB_NFFT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {  // W_512^k = exp(-2 pi i k / 512), k = 0..255
    double sn, cs;
    sincospi(-(double)threadIdx.x / 256.0, &sn, &cs);
    tw[threadIdx.x] = make_float2((float)cs, (float)sn);
  }
  __syncthreads();
  const long long gframe = (long long)blockIdx.x * FB_WARPS + warp;   // frame index over [B, target_len]
  if (gframe >= (long long)B * target_len) return;
  const int b = (int)(gframe / target_len), f = (int)(gframe % target_len);
  float* o = out + gframe * FB_MEL;
  if (f >= n_frames) {  // zero padding happens before normalisation (dataloader.py:335-336, 506)
#pragma unroll
    for (int m = 0; m < 4; ++m) o[lane + 32 * m] = (0.0f - norm_mean) * inv_std;
    return;
  }
  float* xr = re[warp];
  float* xi = im[warp];
  const float* w = wav + (long long)b * ld + (long long)f * FB_SHIFT;
  const float wm = wav_mean ? wav_mean[b] : 0.0f;
  // ---- load, remove the frame's DC offset
  float v[13];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 13; ++k) {
    const int i = lane + 32 * k;
    v[k] = (i < FB_FRAME) ? (w[i] - wm) : 0.f;
    s += v[k];
  }
  const float fmean = warp_sum(s) * (1.0f / FB_FRAME);
#pragma unroll
  for (int k = 0; k < 13; ++k) {
    const int i = lane + 32 * k;
    if (i < FB_FRAME) xi[i] = v[k] - fmean;   // staging (imaginary buffer is free for now)
  }
  __syncwarp();
  // ---- pre-emphasis, window, bit-reversed placement
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int i = lane + 32 * k;
    float y = 0.f;
    if (i < FB_FRAME) {
      const float cur = xi[i], prev = xi[i > 0 ? i - 1 : 0];
      const float win = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)(FB_FRAME - 1));
      y = (cur - 0.97f * prev) * win;
    }
    xr[__brev((unsigned)i) >> 23] = y;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 16; ++k) xi[lane + 32 * k] = 0.f;
  __syncwarp();
  // ---- 512-point radix-2 decimation-in-time FFT
#pragma unroll 1
  for (int half = 1; half < FB_NFFT; half <<= 1) {
    const int tstep = (FB_NFFT / 2) / half;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int bf = lane + 32 * k;               // butterfly 0..255
      const int pos = bf & (half - 1);
      const int i = ((bf - pos) << 1) + pos, j = i + half;
      const float2 t = tw[pos * tstep];
      const float ar = xr[i], ai = xi[i], br = xr[j], bi = xi[j];
      const float pr = br * t.x - bi * t.y, pi = br * t.y + bi * t.x;
      xr[i] = ar + pr; xi[i] = ai + pi;
      xr[j] = ar - pr; xi[j] = ai - pi;
    }
    __syncwarp();
  }
  // ---- power spectrum (bins 0..256) in place
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int i = lane + 32 * k;
    if (i < FB_BINS) xr[i] = xr[i] * xr[i] + xi[i] * xi[i];
  }
  __syncwarp();
  // ---- mel filters, log, normalise
#pragma unroll
  for (int m4 = 0; m4 < 4; ++m4) {
    const int m = lane + 32 * m4;
    const short2 rg = mel_range[m];
    const float* wr = melw + m * FB_BINS;
    float acc = 0.f;
    for (int k = rg.x; k <= rg.y; ++k) acc = fmaf(wr[k], xr[k], acc);
    o[m] = (logf(fmaxf(acc, FLT_EPSILON)) - norm_mean) * inv_std;
  }
}

}  // namespace

extern "C" int avs_fbank(const float* wav, long long ld, int B, int L, int remove_mean, const float* melw,
                         const short* mel_range, float* mean_scratch, float* out, int target_len, float norm_mean,
                         float norm_std, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AVS_REQUIRE(wav && melw && mel_range && out, "avs_fbank: null pointer");
  AVS_REQUIRE(B >= 0 && L >= 0 && target_len > 0 && norm_std != 0.f, "avs_fbank: bad shape");
  AVS_REQUIRE(!remove_mean || mean_scratch != nullptr, "avs_fbank: remove_mean needs mean_scratch [B]");
  if (B == 0) return 0;
  int rc;
  if (remove_mean) {
    AVS_REQUIRE(L > 0, "avs_fbank: empty waveform");
    fbank_mean_kernel<<<B, 256, 0, stream>>>(wav, ld, L, mean_scratch);
    if ((rc = avs_check_launch("fbank_mean_kernel"))) return rc;
  }
  const int n_frames = L < FB_FRAME ? 0 : 1 + (L - FB_FRAME) / FB_SHIFT;   // snip_edges=True
  const long long frames = (long long)B * target_len;
  fbank_kernel<<<(unsigned)ceil_div_ll(frames, FB_WARPS), FB_WARPS * 32, 0, stream>>>(
      wav, ld, L, remove_mean ? mean_scratch : nullptr, melw, reinterpret_cast<const short2*>(mel_range), out, B,
      target_len, n_frames, norm_mean, 1.0f / norm_std);
  return avs_check_launch("fbank_kernel");
}
