// Evaluation statistics on the GPU: per-class average precision and ROC-AUC plus top-1
// accuracy, with scikit-learn's definitions (ties grouped by distinct score), replacing the per-class Python loop of
// src/utilities/stats.py:11-68 (metrics.average_precision_score / roc_auc_score / accuracy_score), which sorts every
// class column on the CPU.
//
// No sort is needed:  AP  = mean over the positives p of  #pos(s >= s_p) / #all(s >= s_p)
//                     AUC = ( #(s_pos > s_neg) + 0.5 #(s_pos == s_neg) ) / (n_pos n_neg)
// so a class costs n_pos x N comparisons, and multi-label targets are sparse (AudioSet: ~2.7 of 527 labels per clip).
// One CTA per class: the positives' scores are compacted into a scratch column, then each thread owns one positive
// and streams the class column through shared-memory tiles.  Counts are exact integers; the final ratios are fp64.
#include "../../include/avsiam_b200.h"
#include "common.cuh"

namespace {

constexpr int ST_THREADS = 256, ST_TILE = 1024;

__global__ void __launch_bounds__(ST_THREADS) eval_stats_kernel(const float* __restrict__ output,
                                                                 const float* __restrict__ target, int N, int C,
                                                                 float* __restrict__ pos_scratch,
                                                                 float* __restrict__ ap_out, float* __restrict__ auc_out) {
  __shared__ float ts[ST_TILE];
  __shared__ unsigned char ty[ST_TILE];
  __shared__ int s_cnt;
  __shared__ double red_ap[ST_THREADS / 32];
  __shared__ unsigned long long red_gt[ST_THREADS / 32], red_eq[ST_THREADS / 32];
  const int c = blockIdx.x, tid = threadIdx.x;
  float* pos = pos_scratch + (size_t)c * N;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int n = tid; n < N; n += ST_THREADS)
    if (target[(size_t)n * C + c] > 0.f) pos[atomicAdd(&s_cnt, 1)] = output[(size_t)n * C + c];
  __syncthreads();
  __threadfence_block();
  const int n_pos = s_cnt, n_neg = N - n_pos;
  if (n_pos == 0 || n_neg == 0) {   // stats.py:56-66: roc_auc_score fails -> auc = -1; AP is 0 (no positives) or 1
    if (tid == 0) {
      ap_out[c] = n_pos == 0 ? 0.f : 1.f;
      auc_out[c] = -1.f;
    }
    return;
  }
  double ap_acc = 0.0;
  unsigned long long gt_acc = 0, eq_acc = 0;
  for (int p0 = 0; p0 < n_pos; p0 += ST_THREADS) {
    const int p = p0 + tid;
    const bool have = p < n_pos;
    const float sp = have ? pos[p] : 0.f;
    unsigned ge_all = 0, ge_pos = 0, gt_neg = 0, eq_neg = 0;
    for (int n0 = 0; n0 < N; n0 += ST_TILE) {
      __syncthreads();
      for (int i = tid; i < ST_TILE; i += ST_THREADS) {
        const int n = n0 + i;
        ts[i] = n < N ? output[(size_t)n * C + c] : 0.f;
        ty[i] = n < N ? (target[(size_t)n * C + c] > 0.f ? 1 : 0) : 2;
      }
      __syncthreads();
      if (have) {
        const int lim = min(ST_TILE, N - n0);
        for (int i = 0; i < lim; ++i) {
          const float s = ts[i];
          const bool is_pos = ty[i] == 1;
          const bool ge = s >= sp;
          ge_all += ge;
          ge_pos += ge && is_pos;
          gt_neg += (!is_pos) && (sp > s);
          eq_neg += (!is_pos) && (sp == s);
        }
      }
    }
    if (have) {
      ap_acc += (double)ge_pos / (double)ge_all;
      gt_acc += gt_neg;
      eq_acc += eq_neg;
    }
  }
  // block reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ap_acc += __shfl_xor_sync(0xffffffffu, ap_acc, o);
    gt_acc += __shfl_xor_sync(0xffffffffu, gt_acc, o);
    eq_acc += __shfl_xor_sync(0xffffffffu, eq_acc, o);
  }
  if ((tid & 31) == 0) {
    red_ap[tid >> 5] = ap_acc;
    red_gt[tid >> 5] = gt_acc;
    red_eq[tid >> 5] = eq_acc;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0;
    unsigned long long g = 0, e = 0;
    for (int w = 0; w < ST_THREADS / 32; ++w) { a += red_ap[w]; g += red_gt[w]; e += red_eq[w]; }
    ap_out[c] = (float)(a / (double)n_pos);
    auc_out[c] = (float)(((double)g + 0.5 * (double)e) / ((double)n_pos * (double)n_neg));
  }
}

// hits += [argmax(target[n]) == argmax(output[n])]  (first maximum, like numpy.argmax); one warp per sample
__global__ void __launch_bounds__(256) eval_acc_kernel(const float* __restrict__ output, const float* __restrict__ target,
                                                       int N, int C, int* __restrict__ hits) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float bo = -INFINITY, bt = -INFINITY;
  int io = C, it = C;
  for (int c = lane; c < C; c += 32) {
    const float o = output[(size_t)n * C + c], t = target[(size_t)n * C + c];
    if (o > bo) { bo = o; io = c; }
    if (t > bt) { bt = t; it = c; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float o2 = __shfl_xor_sync(0xffffffffu, bo, off), t2 = __shfl_xor_sync(0xffffffffu, bt, off);
    const int i2 = __shfl_xor_sync(0xffffffffu, io, off), j2 = __shfl_xor_sync(0xffffffffu, it, off);
    if (o2 > bo || (o2 == bo && i2 < io)) { bo = o2; io = i2; }
    if (t2 > bt || (t2 == bt && j2 < it)) { bt = t2; it = j2; }
  }
  if (lane == 0 && io == it) atomicAdd(hits, 1);
}

}  // namespace

extern "C" int avs_eval_stats(const float* output, const float* target, int N, int C, float* pos_scratch, float* ap,
                              float* auc, int* hits, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AVS_REQUIRE(output && target && pos_scratch && ap && auc && hits, "avs_eval_stats: null pointer");
  AVS_REQUIRE(N > 0 && C > 0, "avs_eval_stats: empty problem");
  eval_stats_kernel<<<C, ST_THREADS, 0, stream>>>(output, target, N, C, pos_scratch, ap, auc);
  int rc = avs_check_launch("eval_stats_kernel");
  if (rc) return rc;
  eval_acc_kernel<<<ceil_div(N, 8), 256, 0, stream>>>(output, target, N, C, hits);
  return avs_check_launch("eval_acc_kernel");
}
