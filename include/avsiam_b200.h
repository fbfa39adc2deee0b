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
/* Drops the library's host-side caches (the TMA descriptors of the GEMM operands, keyed by pointer / shape / pitch).
 * Never required for correctness — entries are matched by value — it only bounds the cache of a long-lived process
 * that keeps changing shapes. */
void avs_reset(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM (tcgen05 / TMEM / TMA).  D[M,N] = epi( sum_k A(m,k) * B(n,k) ), bf16 operands, fp32 accumulate.
 *   a_major / b_major: 0 = reduction dim contiguous (A is [M,K], B is [N,K]);
 *                      1 = M/N contiguous        (A is [K,M], B is [K,N]).
 * Replaces nn.Linear / Conv2d(k=s=16) forward+backward: cav_mae_base.py:51,55 (qkv, proj), :96 (patch embed),
 * :311 (decoder_embed), :334-335 (decoder_pred_*), timm Mlp fc1/fc2 (:138-143).
 * Kernels behind the one entry point: products with a bf16 output, K-major A, M > 128 and N > 128 (every forward and
 * dgrad of the model) run on clusters of two CTAs (tcgen05.mma.cta_group::2, 256 x 256 tiles); fp32 / accumulating
 * outputs (wgrad, split-K) and narrow products on the one-CTA kernel. Both pass the same parity tests; AVS_GEMM_2CTA=0 in
 * the environment forces the one-CTA kernel everywhere.
 * ---------------------------------------------------------------------------------------------- */
enum {
  AVS_EPI_GELU = 1,       /* v = gelu_erf(v); aux_out (optional) receives the pre-activation (bf16) */
  AVS_EPI_DGELU = 2,      /* v = v * gelu'(aux_in) */
  AVS_EPI_OUT_F32 = 4,    /* C is fp32 */
  AVS_EPI_OUT_ATOMIC = 8, /* C is fp32, accumulated in place (split-K, gradient accumulation) */
  AVS_EPI_AUX_GRAD = 16,  /* with AVS_EPI_GELU: aux_out receives gelu'(pre) (bf16) instead of the pre-activation */
  AVS_EPI_MUL_AUX = 32    /* v = v * aux_in: the dGELU epilogue when aux_in already holds gelu'(pre) */
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
  float* colsum;          /* fp32 [N] or NULL : += column sums of the bf16-path output (bias gradient of the Linear
                             that produced this GEMM's dy), accumulated in the epilogue — bf16 outputs only */
} avs_gemm_epilogue_t;

int avs_gemm_bf16(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major, void* C,
                  long long ldc, int M, int N, int K, const avs_gemm_epilogue_t* epi, int split_k, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Random masking (bit-exact index work).  cav_mae_base.py:365-390 (unstructured), :392-439 (structured; the
 * host builds the noise, the device sorts it).  Stable ascending argsort (ties -> lower index):
 *   ids_shuffle[n,:] = argsort(noise[n,:]); ids_restore = argsort(ids_shuffle); mask[n,j] = ids_restore[n,j] >= len_keep
 * ---------------------------------------------------------------------------------------------- */
int avs_mask_argsort(const float* noise, int N, int L, int len_keep, int32_t* ids_shuffle, int32_t* ids_restore,
                     float* mask, void* stream);
/* Structured-mask pattern (cav_mae_base.py:404-423, modes 'time' / 'freq' / 'tf'): noise[n, :, cols[n,j]] = value and
 * noise[n, rows[n,j], :] = value on the [f, t] patch grid; cols int32 [N,kt], rows int32 [N,kf] are the host's
 * random.sample draws (NULL when the count is 0). Followed by avs_mask_argsort. */
int avs_mask_force_noise(float* noise, int N, int f, int t, const int32_t* cols, int kt, const int32_t* rows, int kf,
                         float value, void* stream);
/* out[n,i,:] = x[n, ids[n,i], :], i < keep; byte-exact (replaces torch.gather at cav_mae_base.py:382,431). */
int avs_gather_rows(const void* x, const int32_t* ids, void* out, int N, int L, int keep, int ids_ld, int row_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Patch extraction fused with the kept-token gather (A operand of the patch-embed GEMM, bf16).
 *   audio fp32 [B,T,F] -> [B*keep, ld_out], token f*(T/p)+t, vector (pf,pt)   cav_mae_base.py:444-448
 *   video fp32 [B,C,H,W] -> [B*keep, ld_out], token h*(W/p)+w, vector (c,p,q)  cav_mae_base.py:453
 * ids (int32 [B, ids_ld], first `keep` columns used) may be NULL => all tokens in order.
 * sample_idx (int32 [B], may be NULL): output sample b reads input sample sample_idx[b] (chunk permutation of
 * forward_encoder_mmixed, cav_mae_base.py:533-549).
 * ---------------------------------------------------------------------------------------------- */
int avs_patchify_audio(const float* audio, const int32_t* ids, const int32_t* sample_idx, void* out, int B, int T, int F, int patch, int keep,
                       int ids_ld, int ld_out, void* stream);
int avs_patchify_video(const float* img, const int32_t* ids, const int32_t* sample_idx, void* out, int B, int C, int H, int W, int patch,
                       int keep, int ids_ld, int ld_out, void* stream);
/* table[idx[m],:] += alpha*dy[m,:]  (pos-embed gradient; idx NULL => m % table_rows) */
int avs_scatter_add_rows(const void* dy, const int32_t* idx, float* table, int M, int D, int table_rows, float alpha,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder restore: mask-token fill + un-shuffle + positional + modality embeddings in one pass, and its backward.
 * cav_mae_base.py:604-626 (removes the two int(mask[0].sum()) host syncs).
 *   x bf16 [B, keep_a+keep_v, D]  ->  out bf16 [B, Ta+Tv, D]
 * ---------------------------------------------------------------------------------------------- */
int avs_decoder_restore_fwd(const void* x, const int32_t* ids_restore_a, const int32_t* ids_restore_v,
                            const float* mask_token, const float* pos_a, const float* pos_v, const float* mod_a,
                            const float* mod_v, void* out, int B, int Ta, int Tv, int keep_a, int keep_v, int D,
                            void* stream);
/* dx is fully written; the five fp32 parameter gradients are ACCUMULATED. */
int avs_decoder_restore_bwd(const void* dout, const int32_t* ids_restore_a, const int32_t* ids_restore_v, void* dx,
                            float* dmask_token, float* dpos_a, float* dpos_v, float* dmod_a, float* dmod_v, int B,
                            int Ta, int Tv, int keep_a, int keep_v, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (bf16 activations, fp32 affine + statistics).  Block norms eps 1e-5 (cav_mae_base.py:116,151-152),
 * final norm/norm_a eps 1e-6 (:492-495,:563-566), decoder_norm (:631).
 * Row maps (per-sample concatenation, torch.cat at :503 / slicing at :634-635), one for the x/dx/resid side and
 * one for the y/dy side; stride 0 => identity:
 *   row(r) = (seq_len && stride) ? (r/seq_len)*stride + off + r%seq_len : r
 * bwd: dx = [resid +] LN'(dy_eff), dy_eff = dy[y_row] (+ dpool[r/seq_len]*pool_scale — gradient of the token mean);
 *      dgamma/dbeta are ACCUMULATED (fp32 atomics).
 * ---------------------------------------------------------------------------------------------- */
int avs_layernorm_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                      float* rstd, int M, int D, int seq_len, int x_seq_stride, int x_off, int y_seq_stride, int y_off,
                      void* stream);
int avs_layernorm_bwd(const void* dy, const float* dpool, float pool_scale, const void* x, const float* mean,
                      const float* rstd, const float* gamma, const void* resid, void* dx, float* dgamma, float* dbeta,
                      float* dbias /* or NULL: += column sums of the produced dx (bias gradient of the Linear that
                                      writes this residual stream) */,
                      int M, int D, int seq_len, int x_seq_stride, int x_off, int y_seq_stride, int y_off,
                      void* stream);
/* Two affine sets in ONE launch: rows [0, split_row) use set 0, rows [split_row, M) set 1 — the audio | video token
 * ranges of the shared encoder, whose Block picks norm{1,2}_a / norm{1,2}_v per modality (cav_mae_base.py:151-152,
 * 169-170,190-191). Identity row maps. bwd2 accumulates each range's dgamma / dbeta into its own set; dbias (the
 * column sums of dx) is common to both ranges. */
int avs_layernorm_fwd2(const void* x, const float* gamma0, const float* beta0, int split_row, const float* gamma1,
                       const float* beta1, float eps, void* y, float* mean, float* rstd, int M, int D, void* stream);
int avs_layernorm_bwd2(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma0,
                       float* dgamma0, float* dbeta0, int split_row, const float* gamma1, float* dgamma1,
                       float* dbeta1, const void* resid /* or NULL */, void* dx, float* dbias /* or NULL */, int M,
                       int D, void* stream);
/* out fp32 [n_seq, D] = mean over the seq_len tokens of each sequence (.mean(dim=1), cav_mae_base.py:563-566,729) */
int avs_seq_mean_fwd(const void* y, float* out, int n_seq, int seq_len, int D, int y_seq_stride, int y_off,
                     void* stream);
/* backward of a token mean over the segment [off, off+seg_len) of every sequence (rows s*seq_stride + off + t):
 * dx bf16 (overwrite) = dpool[s, :] / seg_len.   av[:, :512].mean(1) | av[:, 512:].mean(1), cav_mae_base.py:1025-1028 */
int avs_seq_mean_bwd(const float* dpool, void* dx, int n_seq, int seg_len, int D, int seq_stride, int off,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Audio front end: batch kaldi log-mel filterbank + loader post-processing (src/dataloader.py:287,323-339,506:
 * waveform - mean, torchaudio.compliance.kaldi.fbank(htk_compat, hanning, 128 bins, 25 ms / 10 ms, dither 0),
 * zero-pad / crop to target_len, (x - norm_mean) / norm_std).  16 kHz mono.
 *   wav fp32 [B, L] (row pitch ld);  melw fp32 [128, 257] triangular weights and mel_range int16 [128, 2] (first /
 *   last non-zero FFT bin of each filter), both prepared by the host wrapper as kaldi's get_mel_banks defines them;
 *   mean_scratch fp32 [B] (only when remove_mean);  out fp32 [B, target_len, 128].
 * ------------------------------------------------------------------------------------------------ */
int avs_fbank(const float* wav, long long ld, int B, int L, int remove_mean, const float* melw, const short* mel_range,
              float* mean_scratch, float* out, int target_len, float norm_mean, float norm_std, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input pipeline, per-sample transforms of AudiosetDataset.__getitem__ batched on the device.
 * avs_fbank_augment (src/dataloader.py:491-516): SpecAugment bands filled with 0.0, (x - mean) / std (unless
 *   skip_norm), + (noise * noise_scale[b]) / 10 (noise NULL = off; noise_scale = the loader's np.random.rand()), torch.roll along time.  fbank / noise / out fp32 [B, T, F]
 *   (F % 4 == 0, out != fbank); params int32 [B, 6] = f0, f1, t0, t1 (bands [start, end)), shift, unused.
 * avs_frames_preprocess (:152-155,455-456): frames uint8 [N, C, H, W] -> / 255 -> antialiased bicubic resize to
 *   [out_h, out_w] -> (x - mean[c]) / std[c]; fp32 out [N, C, out_h, out_w]; N * C <= 65535 per call.
 *   wx fp32 [out_w, taps_x], xmin / xsize int32 [out_w]: tap window and normalised weights per output column as
 *   ATen's _upsample_bicubic2d_aa computes them (prepared by the host wrapper); wy / ymin / ysize likewise for rows.
 *   tile_rows = output rows per CTA; span_rows >= max over tiles of the input rows a tile's vertical windows cover
 *   (ymin[last] + ysize[last] - ymin[first]); 4 * (256 + taps_x*out_w + span_rows*out_w + 4*W) bytes of shared memory.
 * avs_mix_frames (:419-420): image[n] = weight[n] * image[n] + (1 - weight[n]) * image2[n], in place.
 * ------------------------------------------------------------------------------------------------ */
int avs_fbank_augment(const float* fbank, const int* params, const float* noise_scale, const float* noise, float* out,
                      int B, int T, int F, float norm_mean, float norm_std, int skip_norm, void* stream);
int avs_frames_preprocess(const unsigned char* frames, int N, int C, int H, int W, const float* wx, const int* xmin,
                          const int* xsize, int taps_x, const float* wy, const int* ymin, const int* ysize, int taps_y,
                          int out_h, int out_w, int tile_rows, int span_rows, const float* mean, const float* stdv,
                          float* out, void* stream);
int avs_mix_frames(float* image, const float* image2, const float* weight, int N, long long per_sample, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation statistics (src/utilities/stats.py:11-68): per-class average precision and ROC-AUC with scikit-learn's
 * tie-grouped definitions, and top-1 accuracy.  output / target fp32 [N, C] row-major (target > 0 = positive);
 * pos_scratch fp32 [C, N]; ap / auc fp32 [C] (classes without a positive or without a negative: auc = -1, ap = 0 / 1);
 * hits int32 [1], += number of samples whose argmax(output) == argmax(target) (caller zeroes it).
 * ------------------------------------------------------------------------------------------------ */
int avs_eval_stats(const float* output, const float* target, int N, int C, float* pos_scratch, float* ap, float* auc,
                   int* hits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Classification losses of the finetune loop (traintest_ft_base.py:106-109,148-158), mean reduction, fp32:
 *   avs_bce_with_logits    nn.BCEWithLogitsLoss()(logits, target), n = B*C elements
 *   avs_cross_entropy_prob nn.CrossEntropyLoss()(logits [B,C], target [B,C]) with probability targets (the loader's
 *                          label vectors, dataloader.py:497-503; they need not sum to 1)
 * *loss += the loss (caller zeroes it); dlogits (optional) = d loss / d logits.
 * ------------------------------------------------------------------------------------------------ */
int avs_bce_with_logits(const float* logits, const float* target, float* loss, float* dlogits, long long n,
                        void* stream);
int avs_cross_entropy_prob(const float* logits, const float* target, float* loss, float* dlogits, int B, int C,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Retrieval (retrieval.py:27-52).  avs_cosine_sim replaces get_sim_mat's Python double loop:
 *   sim[i, j] = (a_i . b_j) / (|a_i| |b_j|), a fp32 [n, d], b fp32 [m, d], sim fp32 [n, m]; norm_scratch fp32 [n + m].
 * avs_retrieval_ranks gives what compute_metrics reads off its row sort of a square sim [n, n]:
 *   greater[i] = #{j : sim[i,j] > sim[i,i]},  equal[i] = #{j : sim[i,j] == sim[i,i]} (>= 1, the diagonal itself);
 *   the sorted positions holding the diagonal's value are greater[i] .. greater[i] + equal[i] - 1.
 * ------------------------------------------------------------------------------------------------ */
int avs_cosine_sim(const float* a, const float* b, int n, int m, int d, float* norm_scratch, float* sim, void* stream);
int avs_retrieval_ranks(const float* sim, int n, int* greater, int* equal, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Finetune classification heads: logits = Linear(LayerNorm(x)), x fp32 [B, D] pooled features
 * (nn.Sequential(nn.LayerNorm(D), nn.Linear(D, C)): mlp_head / mlp_head_a / mlp_head_mm, cav_mae_base.py:813-815).
 * All fp32, parameters read from the fp32 master copy.  fwd saves xhat [B,D], rstd [B], y [B,D] for bwd.
 * bwd ACCUMULATES dW [C,D], dbias [C], dgamma [D], dbeta [D] and overwrites dx [B,D]; dy_scratch fp32 [B,D].
 * ------------------------------------------------------------------------------------------------ */
int avs_head_fwd(const float* x, const float* gamma, const float* beta, float eps, const float* W, const float* bias,
                 float* xhat, float* rstd, float* y, float* logits, int B, int C, int D, void* stream);
int avs_head_bwd(const float* dlogits, const float* xhat, const float* rstd, const float* y, const float* gamma,
                 const float* W, float* dW, float* dbias, float* dgamma, float* dbeta, float* dy_scratch, float* dx,
                 int B, int C, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused attention (flash-style), packed QKV in / O out.  Attention.forward, cav_mae_base.py:58-77.
 *   qkv bf16 [n_seq*S, ld_qkv] = [q | k | v], each H*head_dim wide; out bf16 [n_seq*S, ld_o];
 *   lse2 fp32 [n_seq, H, S] (log2-domain logsumexp, saved for backward); delta: scratch [n_seq, H, S].
 * ---------------------------------------------------------------------------------------------- */
int avs_attention_fwd(const void* qkv, long long ld_qkv, void* out, long long ld_o, float* lse2, int n_seq, int S,
                      int H, int head_dim, void* stream);
/* dbias: optional fp32 [3*H*head_dim], += column sums of dQKV (gradient of attn.qkv.bias, cav_mae_base.py:51) —
 * accumulated in the tcgen05 kernel's epilogues, otherwise one extra pass over dQKV */
int avs_attention_bwd(const void* qkv, long long ld_qkv, const void* out, const void* dout, long long ld_o,
                      const float* lse2, float* delta, void* dqkv, float* dbias, int n_seq, int S, int H, int head_dim,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Masked-MSE reconstruction loss, target read from the raw input (patchify + forward_mae_loss,
 * cav_mae_base.py:343-351,663-683).  kind 0 = audio [B,d0=T,d1=F], kind 1 = video [B,C,d0=H,d1=W].
 *   fwd: *loss_accum += sum_masked mean_e (pred-target)^2 / n_masked      (caller zeroes loss_accum)
 *   bwd: dpred = mask * 2 (pred-target) / (P*n_masked) * (*upstream)      (upstream NULL => 1)
 * ---------------------------------------------------------------------------------------------- */
int avs_mae_loss_fwd(const void* pred, const float* input, const float* mask, int kind, int B, int patch, int C,
                     int d0, int d1, float n_masked, float* loss_accum, void* stream);
int avs_mae_loss_bwd(const void* pred, const float* input, const float* mask, int kind, int B, int patch, int C,
                     int d0, int d1, float n_masked, const float* upstream, void* dpred, void* stream);

/* ------------------------------------------------------------------------------------------------
 * InfoNCE over L2-normalised embeddings (forward_contrastive, cav_mae_base.py:641-661), fp32.
 *   ea, ev fp32 [N, D] (global batch); temperature 0.05; bidirect => mean of both directions.
 *   bwd returns gradients for rows [row0,row0+rows) only (this rank's slice); `weight` multiplies the loss
 *   (contrast_loss_weight x world-size factor of GatherLayer.backward, gather_layer.py:34-37).
 * ---------------------------------------------------------------------------------------------- */
size_t avs_infonce_workspace_bytes(int N, int D);
int avs_infonce_fwd(const float* ea, const float* ev, int N, int D, float temperature, int bidirect, float* workspace,
                    float* loss_out, float* acc_out, void* stream);
int avs_infonce_bwd(int N, int D, float temperature, int bidirect, float weight, const float* upstream,
                    float* workspace, int row0, int rows, float* scratch /* 2*rows*D */, float* d_ea, float* d_ev,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Flat-arena optimizer plumbing.  torch.optim.Adam(lr, betas=(0.95,0.999), eps=1e-8, weight_decay=5e-7) +
 * GradScaler unscale/inf-skip (traintest_cavmae_base.py:64-66,138-140).  decoupled=1 => AdamW.
 * ---------------------------------------------------------------------------------------------- */
int avs_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16 /* or NULL */, long long n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, int decoupled,
                  const float* inv_scale /* or NULL */, const float* found_inf /* or NULL */,
                  const uint8_t* active_chunks /* or NULL: one byte per 64 elements, 0 = parameter had no gradient
                                                  (torch.optim.Adam skips grad-None parameters) */,
                  void* stream);
/* Same step with torch's param_groups (traintest_ft_base.py:78-83 builds three: base lr, lr*head_lr, lr*mm_lr):
 * group_chunks[c] = 0 -> chunk skipped, k > 0 -> chunk belongs to group k-1 and is stepped with lr[k-1] /
 * weight_decay[k-1] (host arrays of n_groups <= 8 entries). tick_state (optional, 16 device bytes
 * {int step; float bc1; float bc2_sqrt}, zero-initialised by the caller): the step counter is kept on the device and
 * advances only when found_inf is clear, so a GradScaler-skipped step leaves the bias corrections where torch's
 * per-parameter counters leave them; `step` is then ignored. */
int avs_adam_step_groups(float* p, const float* g, float* m, float* v, void* shadow_bf16 /* or NULL */, long long n,
                         const float* lr, const float* weight_decay, int n_groups, float beta1, float beta2, float eps,
                         int step, int decoupled, const float* inv_scale /* or NULL */,
                         const float* found_inf /* or NULL */, const uint8_t* group_chunks /* or NULL */,
                         void* tick_state /* or NULL */, void* stream);
int avs_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
int avs_colsum_bf16(const void* dy, long long ld, float* out_accum, int M, int N, float alpha, void* stream);
int avs_found_inf(const float* g, long long n, float* flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVSIAM_B200_H */
