"""Differentiable all-gather of the pooled embeddings (reference: src/models/gather_layer.py:21-37).

`GatherLayer.apply(x) -> tuple of W tensors` keeps the reference's API (call sites cav_mae_base.py:724-725): forward
= all_gather; backward = all_reduce of the stacked output grads, own slice returned. The engine itself uses
`all_gather_embeddings`, which moves both modalities in ONE collective ([B, 2D] packed) and, in backward, needs no
collective at all: every rank computes the identical global loss, so the all-reduced slice equals W x the local
slice — the InfoNCE backward kernel applies that factor directly.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GatherLayer(torch.autograd.Function):
    """Gather tensors from all processes, supporting backward propagation (gather_layer.py:21-37)."""

    @staticmethod
    def forward(ctx, x):
        world = dist.get_world_size()
        out = torch.empty((world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out.view(-1), x.contiguous().view(-1))
        return tuple(out.unbind(0))

    @staticmethod
    def backward(ctx, *grads):
        all_gradients = torch.stack(grads)
        dist.all_reduce(all_gradients)
        return all_gradients[dist.get_rank()]


def all_gather_embeddings(ea: torch.Tensor, ev: torch.Tensor, group=None):
    """ea, ev fp32 [B, D] -> global [W*B, D] each (rank-major, like torch.cat(GatherLayer.apply(x), dim=0))."""
    B, D = ea.shape
    world = dist.get_world_size(group)
    packed = torch.cat([ea, ev], dim=1).contiguous()                  # [B, 2D]: one message for both modalities
    out = torch.empty(world * B, 2 * D, dtype=ea.dtype, device=ea.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out[:, :D].contiguous(), out[:, D:].contiguous()
