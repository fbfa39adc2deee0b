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


class PendingGather:
    """An all-gather of the pooled embeddings in flight: issued as soon as the encoder's final norm has produced them
    (before the fusion blocks and the 8-block decoder run), joined just before the InfoNCE kernels need the global
    batch — the exchange and, more importantly, the wait for the slowest rank's encoder hide under the MAE branch."""

    def __init__(self, ea: torch.Tensor, ev: torch.Tensor, group=None):
        B, D = ea.shape
        self.D = D
        world = dist.get_world_size(group)
        self.packed = torch.cat([ea, ev], dim=1).contiguous()              # [B, 2D]: one message for both modalities
        self.out = torch.empty(world * B, 2 * D, dtype=ea.dtype, device=ea.device)
        self.work = dist.all_gather_into_tensor(self.out, self.packed, group=group, async_op=True)

    def wait(self):
        if self.work is not None:
            self.work.wait()          # stream-level join on NCCL (no host sync)
            self.work = None
        D = self.D
        return self.out[:, :D].contiguous(), self.out[:, D:].contiguous()


def all_gather_embeddings(ea: torch.Tensor, ev: torch.Tensor, group=None):
    """ea, ev fp32 [B, D] -> global [W*B, D] each (rank-major, like torch.cat(GatherLayer.apply(x), dim=0))."""
    B, D = ea.shape
    world = dist.get_world_size(group)
    packed = torch.cat([ea, ev], dim=1).contiguous()                  # [B, 2D]: one message for both modalities
    out = torch.empty(world * B, 2 * D, dtype=ea.dtype, device=ea.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out[:, :D].contiguous(), out[:, D:].contiguous()
