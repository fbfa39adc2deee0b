/*
 * avsiam_b200 — C-ABI of libavsiam_b200.so (sm_100a CUDA kernels for the AVSiam pretraining hot path).
 *
 * Drop-in boundary: the reference has no FFI of its own — every op below replaces a torch
 * library call made from /root/reference/src/models/cav_mae_base.py or src/traintest_cavmae_base.py; the
 * file:line each entry replaces is cited beside it.  The Python host (avsiam_b200/ops.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - All pointers are DEVICE pointers on the current device (cudaSetDevice by the caller). The caller owns
 *    every buffer; the library never allocates, frees or retains device memory.
 *  - `stream` is a cudaStream_t passed as void*. All work is enqueued asynchronously on it; no host sync;
 *    every entry is CUDA-graph capturable.
 *  - Return 0 on success; <0 argument/shape/alignment error (nothing launched); >0 a cudaError_t.
 *    avs_last_error() returns a thread-local message. No CPU fallback exists: unsupported input => error.
 *  - "bf16" = __nv_bfloat16; row-major; `ld*` = row pitch in ELEMENTS.
 */
#ifndef AVSIAM_B200_H
#define AVSIAM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* avs_last_error(void);
int avs_version(void);
/* number of kernels launched by this library in this process since load / last reset (bench.py gpu_launches) */
long long avs_launch_count(void);
void avs_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM (tcgen05 / TMEM / TMA).  D[M,N] = epi( sum_k A(m,k) * B(n,k) ), bf16 operands, fp32 accumulate.
 *   a_major / b_major: 0 = reduction dim contiguous (A is [M,K], B is [N,K]);
 *                      1 = M/N contiguous        (A is [K,M], B is [K,N]).
 * Replaces nn.Linear / Conv2d(k=s=16) forward+backward: cav_mae_base.py:51,55 (qkv, proj), :96 (patch embed),
 * :311 (decoder_embed), :334-335 (decoder_pred_*), timm Mlp fc1/fc2 (:138-143).
 * ---------------------------------------------------------------------------------------------- */
enum {
  AVS_EPI_GELU = 1,       /* v = gelu_erf(v); aux_out (optional) receives the pre-activation (bf16) */
  AVS_EPI_DGELU = 2,      /* v = v * gelu'(aux_in) */
  AVS_EPI_OUT_F32 = 4,    /* C is fp32 */
  AVS_EPI_OUT_ATOMIC = 8  /* C is fp32, accumulated in place (split-K, gradient accumulation) */
};
typedef struct {
  int flags;
  float alpha;            /* v *= alpha (applied after bias/rowadd/activation, before resid) */
  const float* bias;      /* [N] or NULL */
  const void* resid;      /* bf16 [M, ld_resid] or NULL : v += resid */
  long long ld_resid;
  const void* aux_in;     /* bf16 [M, ld_aux] */
  void* aux_out;          /* bf16 [M, ld_aux] */
  long long ld_aux;
  const float* rowadd;    /* fp32 [rowadd_rows, N] or NULL : v += rowadd[rowidx ? rowidx[m] : m % rowadd_rows] */
  const int32_t* rowidx;  /* [M] or NULL */
  int rowadd_rows;
} avs_gemm_epilogue_t;

int avs_gemm_bf16(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major, void* C,
                  long long ldc, int M, int N, int K, const avs_gemm_epilogue_t* epi, int split_k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVSIAM_B200_H */
