"""CUDA-graph capture of one whole training step.

The step is ~500 kernel launches issued from Python (forward, hand-written reverse pass, fused Adam). `GraphedTrainStep`
records them ONCE — forward, `loss.backward()` and `optimizer.step()` — into a `torch.cuda.CUDAGraph` for a fixed batch
shape and replays the graph afterwards: no Python, no ctypes, no launch gaps, no allocator traffic in the steady state.
Everything the reference loop does per iteration stays inside the graph: the mask draws (`torch.rand` under capture uses
the graph-safe Philox offsets, so every replay draws fresh noise), the gradient zeroing, the bucketed all-reduce (N > 1:
NCCL collectives are capturable) and the Adam update, whose step counter and bias corrections live on the device
(`FusedAdam.capturable`). Learning-rate / weight-decay values are launch parameters baked into the graph: when a
scheduler changes them the step is re-captured (once per change).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(net, optimizer); loss = step(audio, imgs)  (loss: 0-d device tensor, valid until the next
    call). `net` is avsiam_b200.CAVMAE_BASE or B200DDP around it with `direct_grads = True`; `optimizer` a FusedAdam."""

    def __init__(self, net, optimizer, mask_ratio_a: float = 0.75, mask_ratio_v: float = 0.75,
                 mae_loss_weight: float = 1.0, contrast_loss_weight: float = 0.01, mask_mode: str = "unstructured",
                 warmup: int = 2):
        self.net, self.opt = net, optimizer
        self.model = net.module if hasattr(net, "module") else net
        if not getattr(self.model, "direct_grads", False):
            raise ValueError("GraphedTrainStep needs model.direct_grads = True (gradients stay in the arena)")
        if mask_mode != "unstructured":
            raise ValueError("GraphedTrainStep: structured masks draw from Python's `random` on the host every step and "
                             "cannot be replayed from a graph")
        self.kw = dict(mae_loss_weight=mae_loss_weight, contrast_loss_weight=contrast_loss_weight, mask_mode=mask_mode)
        self.ratios = (mask_ratio_a, mask_ratio_v)
        self.warmup = warmup
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_in: Optional[Sequence[torch.Tensor]] = None
        self.static_out = None
        self._sig = None
        self.launches_per_step = 0

    def _eager_step(self, a, v):
        out = self.net(a, v, self.ratios[0], self.ratios[1], **self.kw)
        self.opt.zero_grad(set_to_none=True)
        out[0].backward()
        self.opt.step()
        return out

    def _hyper(self):
        return tuple((g["lr"], g["weight_decay"], tuple(g["betas"]), g["eps"]) for g in self.opt.param_groups)

    def _capture(self, audio, imgs):
        from . import ops
        self.opt.capturable = True
        self.static_in = (torch.empty_like(audio), torch.empty_like(imgs))
        self.static_in[0].copy_(audio)
        self.static_in[1].copy_(imgs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the default stream, as capture requires
            for _ in range(self.warmup):
                self._eager_step(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()                           # the warm-up's activations go back to the driver
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_out = self._eager_step(*self.static_in)
        self.launches_per_step = ops.launch_count() - n0
        self._sig = (tuple(audio.shape), tuple(imgs.shape), self._hyper())

    def __call__(self, audio: torch.Tensor, imgs: torch.Tensor):
        sig = (tuple(audio.shape), tuple(imgs.shape), self._hyper())
        if self.graph is None or sig != self._sig:
            self.graph = None
            self.static_out = None
            self._capture(audio, imgs)
        self.static_in[0].copy_(audio, non_blocking=True)
        self.static_in[1].copy_(imgs, non_blocking=True)
        self.graph.replay()
        return self.static_out[0]

    @property
    def outputs(self):
        """The 8-tuple of the last replay (static tensors: loss, loss_mae, loss_mae_a, loss_mae_v, loss_c, masks, c_acc)."""
        return self.static_out
