"""Thin torch-tensor wrappers over the C-ABI (one function per entry of include/avsiam_b200.h).

PyTorch is used only for device memory and streams: every wrapper checks dtype / contiguity / device, passes raw
device pointers plus `torch.cuda.current_stream()` to the library and raises RuntimeError on a non-zero return.
Nothing here computes on the CPU and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import EPI_AUX_GRAD, EPI_DGELU, EPI_GELU, EPI_MUL_AUX, EPI_OUT_ATOMIC, EPI_OUT_F32, GemmEpilogue

MAJOR_K, MAJOR_MN = 0, 1
BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str, contiguous: bool = True):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (avsiam_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t


def _row_major_2d(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name}: expected a 2-D tensor with unit inner stride")
    return t.stride(0)


# --------------------------------------------------------------------------------------------- GEMM
def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, M: int, N: int, K: int, *, a_major: int = MAJOR_K,
         b_major: int = MAJOR_K, bias: Optional[torch.Tensor] = None, gelu: bool = False,
         aux_out: Optional[torch.Tensor] = None, dgelu_aux: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, rowadd: Optional[torch.Tensor] = None,
         rowidx: Optional[torch.Tensor] = None, alpha: float = 1.0, accumulate: bool = False, split_k: int = 1,
         colsum: Optional[torch.Tensor] = None, aux_grad: bool = False, mul_aux: Optional[torch.Tensor] = None):
    """out[M,N] = epi(sum_k A(m,k) B(n,k)); see avs_gemm_bf16. `a`/`b` are 2-D bf16 views (row pitch = stride(0)).
    gelu + aux_out: aux_out receives the pre-activation, or gelu'(pre) with aux_grad=True; mul_aux: out = acc * mul_aux
    (the stored derivative) — the pair replaces aux_out=pre / dgelu_aux=pre with one multiply in the backward epilogue."""
    _chk(a, BF16, "gemm.a", contiguous=False)
    _chk(b, BF16, "gemm.b", contiguous=False)
    lda, ldb = _row_major_2d(a, "gemm.a"), _row_major_2d(b, "gemm.b")
    ldc = _row_major_2d(out, "gemm.out")
    flags = 0
    if out.dtype == F32:
        flags |= EPI_OUT_ATOMIC if accumulate else EPI_OUT_F32
    elif out.dtype != BF16 or accumulate:
        raise RuntimeError("gemm.out: bf16 (overwrite) or fp32 (overwrite / accumulate) only")
    ep = GemmEpilogue()
    ep.alpha = alpha
    if bias is not None:
        ep.bias = _chk(bias, F32, "gemm.bias").data_ptr()
    if resid is not None:
        _chk(resid, BF16, "gemm.resid", contiguous=False)
        ep.resid, ep.ld_resid = resid.data_ptr(), _row_major_2d(resid, "gemm.resid")
    if gelu:
        flags |= EPI_GELU
        if aux_out is not None:
            _chk(aux_out, BF16, "gemm.aux_out", contiguous=False)
            ep.aux_out, ep.ld_aux = aux_out.data_ptr(), _row_major_2d(aux_out, "gemm.aux_out")
    if aux_grad:
        if not (gelu and aux_out is not None):
            raise RuntimeError("gemm: aux_grad needs gelu=True and aux_out")
        flags |= EPI_AUX_GRAD
    if mul_aux is not None:
        if dgelu_aux is not None:
            raise RuntimeError("gemm: mul_aux and dgelu_aux exclude each other")
        flags |= EPI_MUL_AUX
        _chk(mul_aux, BF16, "gemm.mul_aux", contiguous=False)
        ep.aux_in, ep.ld_aux = mul_aux.data_ptr(), _row_major_2d(mul_aux, "gemm.mul_aux")
    if dgelu_aux is not None:
        flags |= EPI_DGELU
        _chk(dgelu_aux, BF16, "gemm.dgelu_aux", contiguous=False)
        ep.aux_in, ep.ld_aux = dgelu_aux.data_ptr(), _row_major_2d(dgelu_aux, "gemm.dgelu_aux")
    if rowadd is not None:
        _chk(rowadd, F32, "gemm.rowadd")
        ep.rowadd, ep.rowadd_rows = rowadd.data_ptr(), rowadd.shape[0]
        if rowidx is not None:
            ep.rowidx = _chk(rowidx, I32, "gemm.rowidx").data_ptr()
    if colsum is not None:   # fp32 [N] += column sums of `out` (bias gradient of the Linear that consumes `out` as dy)
        ep.colsum = _chk(colsum, F32, "gemm.colsum").data_ptr()
    ep.flags = flags
    rc = _lib.lib().avs_gemm_bf16(a.data_ptr(), lda, a_major, b.data_ptr(), ldb, b_major, out.data_ptr(), ldc, M, N, K,
                                  ctypes.byref(ep), split_k, _stream())
    _lib.check(rc, "avs_gemm_bf16")
    return out


# --------------------------------------------------------------------------------------------- masking / index
def mask_argsort(noise: torch.Tensor, len_keep: int):
    """-> ids_shuffle int32 [N,L], ids_restore int32 [N,L], mask fp32 [N,L] (0 keep / 1 remove)."""
    _chk(noise, F32, "mask_argsort.noise")
    N, L = noise.shape
    ids_shuffle = torch.empty(N, L, dtype=I32, device=noise.device)
    ids_restore = torch.empty_like(ids_shuffle)
    mask = torch.empty(N, L, dtype=F32, device=noise.device)
    _lib.check(_lib.lib().avs_mask_argsort(noise.data_ptr(), N, L, len_keep, ids_shuffle.data_ptr(),
                                           ids_restore.data_ptr(), mask.data_ptr(), _stream()), "avs_mask_argsort")
    return ids_shuffle, ids_restore, mask


def mask_force_noise(noise: torch.Tensor, f: int, t: int, cols: Optional[torch.Tensor], rows: Optional[torch.Tensor],
                     value: float = 1.1) -> None:
    """In place: noise[n].view(f, t)[:, cols[n]] = value; noise[n].view(f, t)[rows[n], :] = value  (cav_mae_base.py:404-423)."""
    _chk(noise, F32, "mask_force_noise.noise")
    N, L = noise.shape
    if L != f * t:
        raise RuntimeError(f"mask_force_noise: L={L} != f*t={f}*{t}")
    kt = kf = 0
    if cols is not None and cols.numel():
        _chk(cols, I32, "mask_force_noise.cols"); kt = cols.shape[1]
    if rows is not None and rows.numel():
        _chk(rows, I32, "mask_force_noise.rows"); kf = rows.shape[1]
    _lib.check(_lib.lib().avs_mask_force_noise(noise.data_ptr(), N, f, t, _p(cols) if kt else None, kt,
                                               _p(rows) if kf else None, kf, value, _stream()), "avs_mask_force_noise")


def mask_from_ids(ids_shuffle: torch.Tensor, len_keep: int):
    """ids_restore / mask for SUPPLIED ids_shuffle (int32 [N,L]): sorting the ranks 0..L-1 placed at their
    source index reproduces ids_shuffle exactly, so the same kernel serves both entry points."""
    _chk(ids_shuffle, I32, "mask_from_ids.ids_shuffle")
    N, L = ids_shuffle.shape
    rank = torch.empty(N, L, dtype=F32, device=ids_shuffle.device)
    rank.scatter_(1, ids_shuffle.long(), torch.arange(L, device=ids_shuffle.device, dtype=F32).expand(N, L))
    return mask_argsort(rank, len_keep)


def gather_rows(x: torch.Tensor, ids: torch.Tensor, keep: int) -> torch.Tensor:
    """x [N,L,D] (any dtype, D*itemsize % 16 == 0), ids int32 [N,>=keep] -> [N,keep,D]; byte-exact."""
    if not x.is_cuda or not x.is_contiguous():
        raise RuntimeError("gather_rows.x: contiguous CUDA tensor required")
    _chk(ids, I32, "gather_rows.ids")
    N, L, D = x.shape
    out = torch.empty(N, keep, D, dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().avs_gather_rows(x.data_ptr(), ids.data_ptr(), out.data_ptr(), N, L, keep, ids.shape[1],
                                          D * x.element_size(), _stream()), "avs_gather_rows")
    return out


def patchify_audio(audio: torch.Tensor, ids: Optional[torch.Tensor], keep: int, patch: int, out: torch.Tensor,
                   sample_idx: Optional[torch.Tensor] = None):
    _chk(audio, F32, "patchify_audio.audio")
    _chk(out, BF16, "patchify_audio.out")
    B, T, F_ = audio.shape
    if sample_idx is not None:
        B = _chk(sample_idx, I32, "patchify_audio.sample_idx").numel()
    ids_ld = 0
    if ids is not None:
        _chk(ids, I32, "patchify_audio.ids")
        ids_ld = ids.shape[1]
    _lib.check(_lib.lib().avs_patchify_audio(audio.data_ptr(), _p(ids), _p(sample_idx), out.data_ptr(), B, T, F_, patch, keep, ids_ld,
                                             out.shape[-1], _stream()), "avs_patchify_audio")
    return out


def patchify_video(img: torch.Tensor, ids: Optional[torch.Tensor], keep: int, patch: int, out: torch.Tensor,
                   sample_idx: Optional[torch.Tensor] = None):
    _chk(img, F32, "patchify_video.img")
    _chk(out, BF16, "patchify_video.out")
    B, C, H, W = img.shape
    if sample_idx is not None:
        B = _chk(sample_idx, I32, "patchify_video.sample_idx").numel()
    ids_ld = 0
    if ids is not None:
        _chk(ids, I32, "patchify_video.ids")
        ids_ld = ids.shape[1]
    _lib.check(_lib.lib().avs_patchify_video(img.data_ptr(), _p(ids), _p(sample_idx), out.data_ptr(), B, C, H, W, patch, keep, ids_ld,
                                             out.shape[-1], _stream()), "avs_patchify_video")
    return out


def scatter_add_rows(dy: torch.Tensor, idx: Optional[torch.Tensor], table: torch.Tensor, alpha: float):
    _chk(dy, BF16, "scatter_add_rows.dy")
    _chk(table, F32, "scatter_add_rows.table")
    M, D = dy.shape
    if idx is not None:
        _chk(idx, I32, "scatter_add_rows.idx")
    _lib.check(_lib.lib().avs_scatter_add_rows(dy.data_ptr(), _p(idx), table.data_ptr(), M, D, table.shape[0], alpha,
                                               _stream()), "avs_scatter_add_rows")


def decoder_restore_fwd(x, ira, irv, mask_token, pos_a, pos_v, mod_a, mod_v, out, B, Ta, Tv, ka, kv, D):
    _chk(x, BF16, "restore.x"); _chk(out, BF16, "restore.out")
    _chk(ira, I32, "restore.ira"); _chk(irv, I32, "restore.irv")
    for t, n in ((mask_token, "mask_token"), (pos_a, "pos_a"), (pos_v, "pos_v"), (mod_a, "mod_a"), (mod_v, "mod_v")):
        _chk(t, F32, "restore." + n)
    _lib.check(_lib.lib().avs_decoder_restore_fwd(x.data_ptr(), ira.data_ptr(), irv.data_ptr(), mask_token.data_ptr(),
                                                  pos_a.data_ptr(), pos_v.data_ptr(), mod_a.data_ptr(),
                                                  mod_v.data_ptr(), out.data_ptr(), B, Ta, Tv, ka, kv, D, _stream()),
               "avs_decoder_restore_fwd")
    return out


def decoder_restore_bwd(dout, ira, irv, dx, dmask_token, dpos_a, dpos_v, dmod_a, dmod_v, B, Ta, Tv, ka, kv, D):
    _chk(dout, BF16, "restore_bwd.dout"); _chk(dx, BF16, "restore_bwd.dx")
    for t in (dmask_token, dpos_a, dpos_v, dmod_a, dmod_v):
        _chk(t, F32, "restore_bwd.grad")
    _lib.check(_lib.lib().avs_decoder_restore_bwd(dout.data_ptr(), ira.data_ptr(), irv.data_ptr(), dx.data_ptr(),
                                                  dmask_token.data_ptr(), dpos_a.data_ptr(), dpos_v.data_ptr(),
                                                  dmod_a.data_ptr(), dmod_v.data_ptr(), B, Ta, Tv, ka, kv, D,
                                                  _stream()), "avs_decoder_restore_bwd")


# --------------------------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x, gamma, beta, eps, y, mean, rstd, M, D, seq_len=0, y_seq_stride=0, y_off=0, x_seq_stride=0,
                  x_off=0):
    _chk(x, BF16, "ln.x"); _chk(y, BF16, "ln.y"); _chk(gamma, F32, "ln.gamma"); _chk(beta, F32, "ln.beta")
    _chk(mean, F32, "ln.mean"); _chk(rstd, F32, "ln.rstd")
    _lib.check(_lib.lib().avs_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, y.data_ptr(),
                                            mean.data_ptr(), rstd.data_ptr(), M, D, seq_len, x_seq_stride, x_off,
                                            y_seq_stride, y_off, _stream()), "avs_layernorm_fwd")


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, *, resid=None, dpool=None, pool_scale=0.0,
                  seq_len=0, y_seq_stride=0, y_off=0, x_seq_stride=0, x_off=0, dbias=None):
    if dy is not None:
        _chk(dy, BF16, "ln_bwd.dy")
    if dpool is not None:
        _chk(dpool, F32, "ln_bwd.dpool")
    if resid is not None:
        _chk(resid, BF16, "ln_bwd.resid")
    _chk(x, BF16, "ln_bwd.x"); _chk(dx, BF16, "ln_bwd.dx")
    _chk(dgamma, F32, "ln_bwd.dgamma"); _chk(dbeta, F32, "ln_bwd.dbeta")
    if dbias is not None:
        _chk(dbias, F32, "ln_bwd.dbias")
    _lib.check(_lib.lib().avs_layernorm_bwd(_p(dy), _p(dpool), pool_scale, x.data_ptr(), mean.data_ptr(),
                                            rstd.data_ptr(), gamma.data_ptr(), _p(resid), dx.data_ptr(),
                                            dgamma.data_ptr(), dbeta.data_ptr(), _p(dbias), M, D, seq_len, x_seq_stride, x_off,
                                            y_seq_stride, y_off, _stream()), "avs_layernorm_bwd")


def layernorm_fwd2(x, gamma0, beta0, split_row, gamma1, beta1, eps, y, mean, rstd, M, D):
    """Rows [0, split_row) normalised with (gamma0, beta0), rows [split_row, M) with (gamma1, beta1), one launch."""
    _chk(x, BF16, "ln.x"); _chk(y, BF16, "ln.y"); _chk(mean, F32, "ln.mean"); _chk(rstd, F32, "ln.rstd")
    for t in (gamma0, beta0, gamma1, beta1):
        _chk(t, F32, "ln.affine")
    _lib.check(_lib.lib().avs_layernorm_fwd2(x.data_ptr(), gamma0.data_ptr(), beta0.data_ptr(), split_row,
                                             gamma1.data_ptr(), beta1.data_ptr(), eps, y.data_ptr(), mean.data_ptr(),
                                             rstd.data_ptr(), M, D, _stream()), "avs_layernorm_fwd2")


def layernorm_bwd2(dy, x, mean, rstd, gamma0, dgamma0, dbeta0, split_row, gamma1, dgamma1, dbeta1, dx, M, D, *,
                   resid=None, dbias=None):
    _chk(dy, BF16, "ln_bwd.dy"); _chk(x, BF16, "ln_bwd.x"); _chk(dx, BF16, "ln_bwd.dx")
    for t in (gamma0, dgamma0, dbeta0, gamma1, dgamma1, dbeta1):
        _chk(t, F32, "ln_bwd.affine")
    if resid is not None:
        _chk(resid, BF16, "ln_bwd.resid")
    if dbias is not None:
        _chk(dbias, F32, "ln_bwd.dbias")
    _lib.check(_lib.lib().avs_layernorm_bwd2(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                             gamma0.data_ptr(), dgamma0.data_ptr(), dbeta0.data_ptr(), split_row,
                                             gamma1.data_ptr(), dgamma1.data_ptr(), dbeta1.data_ptr(), _p(resid),
                                             dx.data_ptr(), _p(dbias), M, D, _stream()), "avs_layernorm_bwd2")


def seq_mean_fwd(y, out, n_seq, seq_len, D, y_seq_stride=None, y_off=0):
    _chk(y, BF16, "seq_mean.y"); _chk(out, F32, "seq_mean.out")
    _lib.check(_lib.lib().avs_seq_mean_fwd(y.data_ptr(), out.data_ptr(), n_seq, seq_len, D,
                                           seq_len if y_seq_stride is None else y_seq_stride, y_off, _stream()),
               "avs_seq_mean_fwd")


def seq_mean_bwd(dpool, dx, n_seq, seg_len, D, seq_stride, off):
    _chk(dpool, F32, "seq_mean_bwd.dpool"); _chk(dx, BF16, "seq_mean_bwd.dx")
    _lib.check(_lib.lib().avs_seq_mean_bwd(dpool.data_ptr(), dx.data_ptr(), n_seq, seg_len, D, seq_stride, off,
                                           _stream()), "avs_seq_mean_bwd")


# --------------------------------------------------------------------------------------------- finetune heads
def head_fwd(x, gamma, beta, eps, W, bias):
    """logits = Linear(LayerNorm(x)); x fp32 [B, D]. Returns (logits fp32 [B, C], saved) for head_bwd."""
    _chk(x, F32, "head.x"); _chk(W, F32, "head.W")
    B, D = x.shape
    C = W.shape[0]
    xhat, y = torch.empty_like(x), torch.empty_like(x)
    rstd = torch.empty(B, dtype=F32, device=x.device)
    logits = torch.empty(B, C, dtype=F32, device=x.device)
    _lib.check(_lib.lib().avs_head_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, W.data_ptr(),
                                       bias.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), y.data_ptr(),
                                       logits.data_ptr(), B, C, D, _stream()), "avs_head_fwd")
    return logits, (xhat, rstd, y)


def head_bwd(dlogits, saved, gamma, W, dW, dbias, dgamma, dbeta):
    """Accumulates dW / dbias / dgamma / dbeta; returns dx fp32 [B, D]."""
    xhat, rstd, y = saved
    _chk(dlogits, F32, "head.dlogits")
    B, D = xhat.shape
    C = W.shape[0]
    scratch, dx = torch.empty_like(xhat), torch.empty_like(xhat)
    _lib.check(_lib.lib().avs_head_bwd(dlogits.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), y.data_ptr(),
                                       gamma.data_ptr(), W.data_ptr(), dW.data_ptr(), dbias.data_ptr(),
                                       dgamma.data_ptr(), dbeta.data_ptr(), scratch.data_ptr(), dx.data_ptr(), B, C, D,
                                       _stream()), "avs_head_bwd")
    return dx


# --------------------------------------------------------------------------------------------- attention
def attention_fwd(qkv, out, lse2, n_seq, S, H, head_dim):
    _chk(qkv, BF16, "attn.qkv", contiguous=False); _chk(out, BF16, "attn.out", contiguous=False)
    _chk(lse2, F32, "attn.lse2")
    _lib.check(_lib.lib().avs_attention_fwd(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0),
                                            lse2.data_ptr(), n_seq, S, H, head_dim, _stream()), "avs_attention_fwd")


def attention_bwd(qkv, out, dout, lse2, delta, dqkv, n_seq, S, H, head_dim, dbias=None):
    """dbias: optional fp32 [3*H*head_dim], += column sums of dQKV (the qkv-bias gradient)."""
    for t, n in ((qkv, "qkv"), (out, "out"), (dout, "dout"), (dqkv, "dqkv")):
        _chk(t, BF16, "attn_bwd." + n, contiguous=False)
    if dout.stride(0) != out.stride(0) or dqkv.stride(0) != qkv.stride(0):
        raise RuntimeError("attention_bwd: dout/out and dqkv/qkv must share row pitches")
    if dbias is not None:
        _chk(dbias, F32, "attn_bwd.dbias")
    _lib.check(_lib.lib().avs_attention_bwd(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), dout.data_ptr(),
                                            out.stride(0), lse2.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                                            _p(dbias), n_seq, S, H, head_dim, _stream()), "avs_attention_bwd")


# --------------------------------------------------------------------------------------------- losses
def mae_loss_fwd(pred, inp, mask, kind, B, patch, C, d0, d1, n_masked, loss_accum):
    _chk(pred, BF16, "mae.pred"); _chk(inp, F32, "mae.input"); _chk(mask, F32, "mae.mask")
    _chk(loss_accum, F32, "mae.loss")
    _lib.check(_lib.lib().avs_mae_loss_fwd(pred.data_ptr(), inp.data_ptr(), mask.data_ptr(), kind, B, patch, C, d0,
                                           d1, float(n_masked), loss_accum.data_ptr(), _stream()), "avs_mae_loss_fwd")


def mae_loss_bwd(pred, inp, mask, kind, B, patch, C, d0, d1, n_masked, upstream, dpred):
    _chk(pred, BF16, "mae_bwd.pred"); _chk(dpred, BF16, "mae_bwd.dpred")
    if upstream is not None:
        _chk(upstream, F32, "mae_bwd.upstream")
    _lib.check(_lib.lib().avs_mae_loss_bwd(pred.data_ptr(), inp.data_ptr(), mask.data_ptr(), kind, B, patch, C, d0,
                                           d1, float(n_masked), _p(upstream), dpred.data_ptr(), _stream()),
               "avs_mae_loss_bwd")


def infonce_workspace(N: int, D: int, device) -> torch.Tensor:
    nbytes = _lib.lib().avs_infonce_workspace_bytes(N, D)
    return torch.empty((nbytes + 3) // 4, dtype=F32, device=device)


def infonce_fwd(ea, ev, temperature, bidirect, workspace, loss_out, acc_out):
    _chk(ea, F32, "nce.ea"); _chk(ev, F32, "nce.ev"); _chk(workspace, F32, "nce.ws")
    N, D = ea.shape
    _lib.check(_lib.lib().avs_infonce_fwd(ea.data_ptr(), ev.data_ptr(), N, D, temperature, int(bidirect),
                                          workspace.data_ptr(), loss_out.data_ptr(), acc_out.data_ptr(), _stream()),
               "avs_infonce_fwd")


def infonce_bwd(N, D, temperature, bidirect, weight, upstream, workspace, row0, rows, scratch, d_ea, d_ev):
    _chk(d_ea, F32, "nce_bwd.d_ea"); _chk(d_ev, F32, "nce_bwd.d_ev"); _chk(scratch, F32, "nce_bwd.scratch")
    _lib.check(_lib.lib().avs_infonce_bwd(N, D, temperature, int(bidirect), weight, _p(upstream),
                                          workspace.data_ptr(), row0, rows, scratch.data_ptr(), d_ea.data_ptr(),
                                          d_ev.data_ptr(), _stream()), "avs_infonce_bwd")


# --------------------------------------------------------------------------------------------- optimizer plumbing
def adam_step(p, g, m, v, shadow, lr, beta1, beta2, eps, weight_decay, step, decoupled=False, inv_scale=None,
              found_inf=None, active=None, tick=None):
    """One fused Adam launch over a flat fp32 buffer. `lr` / `weight_decay`: floats, or equally long sequences — one
    entry per param_group, in which case `active` is the per-chunk group map (0 = skip, k = group k-1). `tick`: 16-byte
    device buffer holding the step counter and bias corrections (advances only when `found_inf` is clear)."""
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _chk(t, F32, "adam." + n)
    if shadow is not None:
        _chk(shadow, BF16, "adam.shadow")
    if active is not None:
        _chk(active, torch.uint8, "adam.active")
        if active.numel() * 64 < p.numel():
            raise RuntimeError("adam.active: one byte per 64 parameters required")
    lrs = [float(x) for x in lr] if isinstance(lr, (list, tuple)) else [float(lr)]
    wds = [float(x) for x in weight_decay] if isinstance(weight_decay, (list, tuple)) else [float(weight_decay)] * len(lrs)
    if len(wds) != len(lrs):
        raise RuntimeError("adam: lr and weight_decay need one entry per group")
    if tick is not None and (tick.numel() * tick.element_size() < 16 or not tick.is_cuda):
        raise RuntimeError("adam.tick: 16 device bytes required")
    arr = ctypes.c_float * len(lrs)
    _lib.check(_lib.lib().avs_adam_step_groups(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _p(shadow),
                                               p.numel(), arr(*lrs), arr(*wds), len(lrs), beta1, beta2, eps, int(step),
                                               int(decoupled), _p(inv_scale), _p(found_inf), _p(active), _p(tick),
                                               _stream()), "avs_adam_step_groups")


def cast_f32_to_bf16(src, dst):
    _chk(src, F32, "cast.src"); _chk(dst, BF16, "cast.dst")
    _lib.check(_lib.lib().avs_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()),
               "avs_cast_f32_to_bf16")


def colsum(dy, out, M, N, alpha: float = 1.0):
    _chk(dy, BF16, "colsum.dy", contiguous=False); _chk(out, F32, "colsum.out")
    _lib.check(_lib.lib().avs_colsum_bf16(dy.data_ptr(), dy.stride(0), out.data_ptr(), M, N, alpha, _stream()),
               "avs_colsum_bf16")


def found_inf(g, flag):
    _chk(g, F32, "found_inf.g"); _chk(flag, F32, "found_inf.flag")
    _lib.check(_lib.lib().avs_found_inf(g.data_ptr(), g.numel(), flag.data_ptr(), _stream()), "avs_found_inf")


def launch_count() -> int:
    return int(_lib.lib().avs_launch_count())


def reset_launch_count() -> None:
    _lib.lib().avs_reset_launch_count()


# --------------------------------------------------------------------------------------------- per-family timing
class KernelTimer:
    """CUDA-event timing of every C-ABI call, grouped by kernel family (bench.py's roofline leg).

    `with KernelTimer() as t: step()` brackets each wrapper call with two events recorded on the CURRENT stream
    (the stream the kernels are launched on); `t.summary()` synchronises and returns
    {family: {"ms": total, "calls": n, "work": algorithmic flops or bytes}}. Off by default: the wrappers
    pay one `is None` test when no timer is installed."""

    def __init__(self, detail: bool = False):
        self.records = []
        self.detail = detail     # True: GEMM / attention records are keyed by shape as well

    def __enter__(self):
        global _TIMER
        _TIMER = self
        return self

    def __exit__(self, *exc):
        global _TIMER
        _TIMER = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for fam, s, e, work in self.records:
            d = out.setdefault(fam, {"ms": 0.0, "calls": 0, "work": 0.0})
            d["ms"] += s.elapsed_time(e)
            d["calls"] += 1
            d["work"] += work
        return out


_TIMER: Optional[KernelTimer] = None


def _nbytes(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def _work_gemm(a, b, out, M, N, K, **kw):
    return 2.0 * M * N * K


def _work_attn_fwd(qkv, out, lse2, n_seq, S, H, hd):
    return 4.0 * n_seq * H * S * S * hd


def _work_attn_bwd(qkv, out, dout, lse2, delta, dqkv, n_seq, S, H, hd, dbias=None):
    return 10.0 * n_seq * H * S * S * hd       # 5 S x S x hd products (QK^T, dO V^T, P^T dO, dS^T Q, dS K)


def _detail_gemm(a, b, out, M, N, K, **kw):
    tag = "".join(c for c, f in (("b", kw.get("bias") is not None), ("g", kw.get("gelu")), ("r", kw.get("resid") is not None),
                                  ("d", kw.get("dgelu_aux") is not None or kw.get("mul_aux") is not None), ("p", kw.get("rowadd") is not None),
                                  ("A", kw.get("accumulate"))) if f)
    return f"gemm M={M} N={N} K={K} maj={kw.get('a_major', 0)}{kw.get('b_major', 0)} {tag}"


def _detail_attn(name):
    def f(*a, **k):
        n_seq, S, H, hd = a[6:10] if name == "attention_bwd" else a[-4:]
        return f"{name} n_seq={n_seq} S={S} H={H} hd={hd}"
    return f


def _instrument(family, fn, work_fn, detail_fn=None):
    def wrapped(*a, **k):
        t = _TIMER
        if t is None:
            return fn(*a, **k)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = fn(*a, **k)
        e.record()
        key = detail_fn(*a, **k) if (t.detail and detail_fn is not None) else family
        t.records.append((key, s, e, work_fn(*a, **k)))
        return r
    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


gemm = _instrument("gemm", gemm, _work_gemm, _detail_gemm)
attention_fwd = _instrument("attention_fwd", attention_fwd, _work_attn_fwd, _detail_attn("attention_fwd"))
attention_bwd = _instrument("attention_bwd", attention_bwd, _work_attn_bwd, _detail_attn("attention_bwd"))
layernorm_fwd = _instrument("layernorm_fwd", layernorm_fwd,
                            lambda x, gamma, beta, eps, y, mean, rstd, M, D, **k: 4.0 * M * D)
layernorm_fwd2 = _instrument("layernorm_fwd", layernorm_fwd2,
                             lambda x, g0, b0, split, g1, b1, eps, y, mean, rstd, M, D: 4.0 * M * D)
layernorm_bwd2 = _instrument("layernorm_bwd", layernorm_bwd2,
                             lambda dy, x, mean, rstd, g0, dg0, db0, split, g1, dg1, db1, dx, M, D, **k:
                             (8.0 if k.get("resid") is not None else 6.0) * M * D)
layernorm_bwd = _instrument("layernorm_bwd", layernorm_bwd,
                            lambda dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, **k:
                            (8.0 if k.get("resid") is not None else 6.0) * M * D)
colsum = _instrument("colsum", colsum, lambda dy, out, M, N, alpha=1.0: 2.0 * M * N)
_ACTIVE_ELEMS = {}   # (data_ptr, numel) of an activity bitmap -> parameters it marks active (timing aid only)


def _work_adam(p, g, m, v, shadow, *a, **k):
    """Algorithmic bytes of one Adam launch: 28 B (+2 B bf16 shadow) per ACTIVE parameter — chunks whose bitmap byte is
    0 are skipped without touching p / g / m / v (the inactive `ast_base` copy in the single-pass arrangement)."""
    active = k.get("active", a[9] if len(a) > 9 else None)
    n = p.numel()
    if active is not None:
        key = (active.data_ptr(), active.numel())
        if key not in _ACTIVE_ELEMS:
            _ACTIVE_ELEMS[key] = int((active != 0).sum()) * 64
        n = min(n, _ACTIVE_ELEMS[key])
    return (28.0 + (2.0 if shadow is not None else 0.0)) * n


adam_step = _instrument("adam", adam_step, _work_adam)
cast_f32_to_bf16 = _instrument("cast", cast_f32_to_bf16, lambda src, dst: 6.0 * src.numel())
patchify_audio = _instrument("patchify", patchify_audio, lambda audio, ids, keep, patch, out, sample_idx=None: 3.0 * out.numel())
patchify_video = _instrument("patchify", patchify_video, lambda img, ids, keep, patch, out, sample_idx=None: 3.0 * out.numel())
decoder_restore_fwd = _instrument("decoder_restore", decoder_restore_fwd,
                                  lambda x, ira, irv, mt, pa, pv, ma, mv, out, *a: _nbytes(x, out))
decoder_restore_bwd = _instrument("decoder_restore", decoder_restore_bwd,
                                  lambda dout, ira, irv, dx, *a: _nbytes(dout, dx))
mae_loss_fwd = _instrument("mae_loss", mae_loss_fwd, lambda pred, inp, mask, *a: _nbytes(pred, inp))
mae_loss_bwd = _instrument("mae_loss", mae_loss_bwd, lambda pred, inp, mask, *a: 2.0 * _nbytes(pred) + _nbytes(inp))
infonce_fwd = _instrument("infonce", infonce_fwd, lambda ea, ev, *a: 2.0 * ea.shape[0] * ea.shape[0] * ea.shape[1])
infonce_bwd = _instrument("infonce", infonce_bwd, lambda N, D, *a: 4.0 * N * N * D)
mask_argsort = _instrument("mask_argsort", mask_argsort, lambda noise, keep: 16.0 * noise.numel())
gather_rows = _instrument("gather_rows", gather_rows, lambda x, ids, keep: 2.0 * x.shape[0] * keep * x.shape[2] * x.element_size())
scatter_add_rows = _instrument("scatter_add", scatter_add_rows, lambda dy, idx, table, alpha: _nbytes(dy) * 3.0)
seq_mean_fwd = _instrument("seq_mean", seq_mean_fwd, lambda y, out, n_seq, seq_len, D, **k: 2.0 * n_seq * seq_len * D)
