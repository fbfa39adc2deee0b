"""Classification losses of the finetune loop as single fused CUDA passes (forward value and gradient together).

Reference: traintest_ft_base.py:106-109 builds `nn.BCEWithLogitsLoss()` (AudioSet, multi-label) or
`nn.CrossEntropyLoss()` (VGGSound; the loader hands it float label VECTORS, dataloader.py:497-503, so it is the
probability-target form) and applies it to the [B, C] logits at :148-158.  Both use mean reduction.

    loss = avsiam_b200.losses.bce_with_logits(logits, labels)     # autograd-connected 0-dim fp32 tensor
    loss = avsiam_b200.losses.cross_entropy(logits, labels)
    loss_fn = avsiam_b200.losses.loss_fn('BCE' | 'CE')            # args.loss of run_cavmae_ft_base.py
"""
from __future__ import annotations

import torch

from . import _lib

F32 = torch.float32


def _prep(logits: torch.Tensor, target: torch.Tensor, name: str):
    if not (logits.is_cuda and target.is_cuda):
        raise RuntimeError(f"avsiam_b200.losses.{name} runs on CUDA only — there is no CPU path")
    if logits.shape != target.shape:
        raise ValueError(f"{name}: logits {tuple(logits.shape)} and target {tuple(target.shape)} differ in shape")
    return logits.detach().contiguous().to(F32), target.detach().contiguous().to(F32)


class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        x, y = _prep(logits, target, "bce_with_logits")
        loss = torch.zeros((), dtype=F32, device=x.device)
        dx = torch.empty_like(x) if logits.requires_grad else None
        _lib.check(_lib.lib().avs_bce_with_logits(x.data_ptr(), y.data_ptr(), loss.data_ptr(),
                                                  dx.data_ptr() if dx is not None else None, x.numel(),
                                                  torch.cuda.current_stream().cuda_stream), "avs_bce_with_logits")
        ctx.dx, ctx.dtype = dx, logits.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.dx * g).to(ctx.dtype), None


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        x, y = _prep(logits, target, "cross_entropy")
        if x.dim() != 2:
            raise ValueError("cross_entropy: logits must be [B, C]")
        loss = torch.zeros((), dtype=F32, device=x.device)
        dx = torch.empty_like(x) if logits.requires_grad else None
        _lib.check(_lib.lib().avs_cross_entropy_prob(x.data_ptr(), y.data_ptr(), loss.data_ptr(),
                                                     dx.data_ptr() if dx is not None else None, x.shape[0], x.shape[1],
                                                     torch.cuda.current_stream().cuda_stream), "avs_cross_entropy_prob")
        ctx.dx, ctx.dtype = dx, logits.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.dx * g).to(ctx.dtype), None


def bce_with_logits(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.BCEWithLogitsLoss()(logits, target)."""
    return _BCEFn.apply(logits, target)


def cross_entropy(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.CrossEntropyLoss()(logits [B, C], target [B, C] float label vectors)."""
    return _CEFn.apply(logits, target)


def loss_fn(name: str):
    """`args.loss` of run_cavmae_ft_base.py ('BCE' | 'CE') -> callable(logits, labels), traintest_ft_base.py:106-109."""
    if name == "BCE":
        return bce_with_logits
    if name == "CE":
        return cross_entropy
    raise ValueError(f"unknown loss {name!r} (BCE | CE)")
