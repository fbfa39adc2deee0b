"""Evaluation path: the reference's validation loops and retrieval metrics with the per-sample
arithmetic on the GPU.

  distributed_concat   traintest_cavmae_base.py:21-26   all-gather of per-rank predictions, truncated to the dataset
  validate             traintest_cavmae_base.py:381-424 pretraining losses over the validation loader
  validate_mlp         traintest_cavmae_base.py:426-492 / traintest_ft_base.py (finetune): logits for every sample,
                       sigmoid, mean over frames, mAP / mAUC / acc
  get_sim_mat          retrieval.py:31-37               cosine-similarity matrix (CUDA kernel instead of a B^2 Python loop)
  compute_metrics      retrieval.py:39-52               R@1 / R@5 / R@10 / median rank
  get_retrieval_result retrieval.py:59-95               features -> mean pool -> normalise -> sim -> metrics

Models are the drop-ins of this package (`CAVMAE_BASE`, `CAVMAEFT_BASE`); everything stays on the device until the
final scalars.  Unlike the reference loops nothing here wraps the forward in fp16 autocast: the engine computes in the
model's own mode (bf16 activations / fp32 heads and losses).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .losses import bce_with_logits, cross_entropy
from .stats import calculate_stats

F32 = torch.float32


def distributed_concat(tensor: torch.Tensor, num_total_examples: int) -> torch.Tensor:
    """All ranks' `tensor` concatenated along dim 0 in rank order, cut to `num_total_examples` (the padding that
    SequentialDistributedSampler adds is dropped).  One `all_gather_into_tensor` instead of W clones + a list gather.
    Without an initialised process group this is the single-rank identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor[:num_total_examples]
    W = dist.get_world_size()
    tensor = tensor.contiguous()
    out = tensor.new_empty((W * tensor.shape[0],) + tuple(tensor.shape[1:]))
    dist.all_gather_into_tensor(out, tensor)
    return out[:num_total_examples]


def validate(audio_model, val_loader, val_sampler=None, args=None):
    """-> (loss, loss_mae, loss_mae_a, loss_mae_v, loss_c, c_acc), each the mean over validation batches."""
    device = torch.device("cuda")
    audio_model.eval()
    sums = torch.zeros(6, dtype=F32, device=device)
    n = 0
    with torch.no_grad():
        for a_input, v_input, _ in val_loader:
            a_input = a_input.to(device, non_blocking=True)
            v_input = v_input.to(device, non_blocking=True)
            loss, loss_mae, loss_mae_a, loss_mae_v, loss_c, _, _, c_acc = audio_model(
                a_input, v_input, args.masking_ratio, args.masking_ratio, mae_loss_weight=args.mae_loss_weight,
                contrast_loss_weight=args.contrast_loss_weight, mask_mode=args.mask_mode)
            sums += torch.stack([loss.sum(), loss_mae.sum(), loss_mae_a.sum(), loss_mae_v.sum(), loss_c.sum(),
                                 c_acc.mean()]).float()
            n += 1
    return tuple((sums / max(n, 1)).tolist())                 # the only device -> host read of the loop


def validate_mlp(audio_model, val_loader, val_sampler, mode, args=None, output_pred=False):
    """Finetune validation.  The model is called with `is_eval=True` and returns [B, T, C] logits (T frames; T = 1 for
    'audioonly').  -> (stats, loss) or (stats, probabilities [N, T, C], target [N, C]) with `output_pred`.
    `stats` is avsiam_b200.calculate_stats' dict of device tensors (AP [C], auc [C], acc)."""
    device = torch.device("cuda")
    audio_model.eval()
    # traintest_ft_base.py:328 evaluates `args.loss_fn` (BCEWithLogits for AudioSet, CrossEntropy for VGGSound, set at
    # :106-110); traintest_cavmae_base.py:436 hard-codes BCE. args.loss_fn may be the torch module, a callable, or the
    # run script's --loss string ('BCE' / 'CE'); absent -> BCE.
    lf = getattr(args, "loss_fn", None) if args is not None else None
    if lf is None and args is not None and isinstance(getattr(args, "loss", None), str):
        lf = getattr(args, "loss")
    if isinstance(lf, str):
        loss_of = bce_with_logits if lf.upper().startswith("BCE") else cross_entropy
    elif isinstance(lf, torch.nn.CrossEntropyLoss):
        loss_of = cross_entropy
    elif lf is None or isinstance(lf, torch.nn.BCEWithLogitsLoss):
        loss_of = bce_with_logits
    else:
        loss_of = lf
    preds, labels_all = [], []
    loss_sum = torch.zeros((), dtype=F32, device=device)
    n = 0
    with torch.no_grad():
        for a_input, v_input, labels in val_loader:
            a_input = a_input.to(device, non_blocking=True)
            v_input = v_input.to(device, non_blocking=True)
            labels = labels.to(device, non_blocking=True)
            out = audio_model(a_input, v_input, mode, is_eval=True)
            preds.append(out)
            labels_all.append(labels)
            loss_sum += loss_of(out.float().mean(dim=1), labels.float())
            n += 1
        total = len(val_sampler.dataset) if val_sampler is not None else sum(p.shape[0] for p in preds)
        audio_output = distributed_concat(torch.cat(preds, dim=0), total)
        target = distributed_concat(torch.cat(labels_all, dim=0), total)
        audio_output = torch.sigmoid(audio_output.float())
        stats = calculate_stats(audio_output.mean(dim=1), target.float())
    loss = float(loss_sum / max(n, 1))
    if not output_pred:
        return stats, loss
    return stats, audio_output, target


# ------------------------------------------------------------------------------------------------ retrieval
def get_sim_mat(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """sim[i, j] = cos(a_i, b_j); a [n, d], b [m, d] on the GPU -> fp32 [n, m]."""
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("avsiam_b200.evaluate.get_sim_mat runs on CUDA only — there is no CPU path")
    a, b = a.contiguous().to(F32), b.contiguous().to(F32)
    n, d = a.shape
    m = b.shape[0]
    assert b.shape[1] == d
    sim = torch.empty(n, m, dtype=F32, device=a.device)
    scratch = torch.empty(n + m, dtype=F32, device=a.device)
    _lib.check(_lib.lib().avs_cosine_sim(a.data_ptr(), b.data_ptr(), n, m, d, scratch.data_ptr(), sim.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "avs_cosine_sim")
    return sim


def compute_metrics(x: torch.Tensor) -> Dict[str, float]:
    """R@1 / R@5 / R@10 / median rank of the matching (diagonal) candidate in each row of a square similarity matrix.
    Ties follow the reference: every sorted position holding the diagonal's value is listed."""
    if not x.is_cuda:
        raise RuntimeError("avsiam_b200.evaluate.compute_metrics runs on CUDA only — there is no CPU path")
    x = x.contiguous().to(F32)
    n = x.shape[0]
    assert x.shape == (n, n)
    greater = torch.empty(n, dtype=torch.int32, device=x.device)
    equal = torch.empty(n, dtype=torch.int32, device=x.device)
    _lib.check(_lib.lib().avs_retrieval_ranks(x.data_ptr(), n, greater.data_ptr(), equal.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "avs_retrieval_ranks")
    g, e = greater.cpu().numpy().astype(np.int64), equal.cpu().numpy().astype(np.int64)
    ind = np.repeat(g, e) + (np.arange(int(e.sum())) - np.repeat(np.cumsum(e) - e, e))   # g_i .. g_i + e_i - 1 per row
    return {"R1": float(np.sum(ind == 0)) / len(ind), "R5": float(np.sum(ind < 5)) / len(ind),
            "R10": float(np.sum(ind < 10)) / len(ind), "MR": float(np.median(ind) + 1)}


def get_retrieval_result(audio_model, val_loader, direction: str = "audio") -> Tuple[float, float, float, float]:
    """direction 'audio' = audio -> visual retrieval, 'video' = visual -> audio.  The model is called in mode
    'retrieval' (cav_mae_base.py:885-917) and returns per-token features (audio [B, 512, D], video frame [B, 196, D])."""
    if direction not in ("audio", "video"):
        raise ValueError("direction must be 'audio' or 'video'")
    device = torch.device("cuda")
    audio_model.eval()
    fa, fv = [], []
    with torch.no_grad():
        for a_input, v_input, _ in val_loader:
            a_out, v_out = audio_model(a_input.to(device), v_input.to(device), "retrieval")
            fa.append(torch.nn.functional.normalize(a_out.float().mean(dim=1), dim=-1))
            fv.append(torch.nn.functional.normalize(v_out.float().mean(dim=1), dim=-1))
    fa, fv = torch.cat(fa), torch.cat(fv)
    sim = get_sim_mat(fa, fv) if direction == "audio" else get_sim_mat(fv, fa)
    r = compute_metrics(sim)
    return r["R1"], r["R5"], r["R10"], r["MR"]
