"""GPU evaluation statistics (reference: src/utilities/stats.py — calculate_stats / d_prime; consumed by
traintest_ft_base.py:196-220 as mAP = mean AP, mAUC = mean AUC, d' = d_prime(mAUC), acc).

    stats = avsiam_b200.calculate_stats(output, target)      # CUDA tensors [N, C]
    mAP = float(stats["AP"].mean())

Returns tensors instead of the reference's list of per-class dicts (its precision/recall/fpr curve samples are only
kept for plotting and are not reproduced); `as_reference_list()` converts to the `[{'AP':…, 'auc':…, 'acc':…}]` shape.
"""
from __future__ import annotations

import math
from statistics import NormalDist
from typing import Dict

import torch

from . import _lib


def calculate_stats(output: torch.Tensor, target: torch.Tensor) -> Dict[str, torch.Tensor]:
    if not (output.is_cuda and target.is_cuda):
        raise RuntimeError("avsiam_b200.calculate_stats runs on CUDA only — there is no CPU path")
    output = output.contiguous().float()
    target = target.contiguous().float()
    N, C = output.shape
    assert target.shape == (N, C)
    scratch = torch.empty(C, N, dtype=torch.float32, device=output.device)
    ap = torch.empty(C, dtype=torch.float32, device=output.device)
    auc = torch.empty(C, dtype=torch.float32, device=output.device)
    hits = torch.zeros(1, dtype=torch.int32, device=output.device)
    rc = _lib.lib().avs_eval_stats(output.data_ptr(), target.data_ptr(), N, C, scratch.data_ptr(), ap.data_ptr(),
                                   auc.data_ptr(), hits.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "avs_eval_stats")
    return {"AP": ap, "auc": auc, "acc": hits.float() / N}


def as_reference_list(stats: Dict[str, torch.Tensor]):
    ap, auc, acc = stats["AP"].cpu().tolist(), stats["auc"].cpu().tolist(), float(stats["acc"])
    return [{"AP": a, "auc": u, "acc": acc} for a, u in zip(ap, auc)]


def d_prime(auc: float) -> float:
    """stats.py:6-9."""
    return NormalDist().inv_cdf(auc) * math.sqrt(2.0)
