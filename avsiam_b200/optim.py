"""Fused flat-arena Adam (reference: torch.optim.Adam at src/traintest_cavmae_base.py:64-66 — lr 2e-4,
betas (0.95, 0.999), eps 1e-8, coupled-L2 weight_decay 5e-7 — plus the GradScaler unscale / inf-skip at :138-140).

One kernel launch updates every parameter of the model's arena, reads the gradients where the backward kernels left
them (no per-tensor foreach lists), and writes the bf16 weight shadow the next forward's GEMMs consume.
Parameters that received no gradient in the last backward are skipped exactly like torch.optim.Adam skips
`p.grad is None` (no weight decay, no moment update) through a per-chunk activity bitmap.
Keeps `param_groups[i]['lr']` (LR schedulers at :73-74 mutate it) and a `state_dict()` for `best_optim_state.pth`.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch

from . import ops

_MODEL_OF_PARAM: "weakref.WeakValueDictionary[int, torch.nn.Module]" = weakref.WeakValueDictionary()


def register_model(model) -> None:
    """Lets `FusedAdam(params, ...)` find the arena when it is constructed from a bare parameter list, as the
    reference loop does (`torch.optim.Adam(trainables, ...)`, traintest_cavmae_base.py:62-66)."""
    for p in model.parameters():
        _MODEL_OF_PARAM[id(p)] = model


class FusedAdam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True   # GradScaler.step hands us the scaler; unscale + inf-skip run in the kernel

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, *, model=None,
                 decoupled: bool = False):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if model is None:
            first = self.param_groups[0]["params"][0]
            model = _MODEL_OF_PARAM.get(id(first))
        if model is None:
            raise ValueError("FusedAdam needs model=<avsiam_b200.CAVMAE_BASE> (it steps the model's flat arena)")
        self.model = model.module if hasattr(model, "module") else model
        self.decoupled = decoupled
        self._step = 0
        self._m: Optional[torch.Tensor] = None
        self._v: Optional[torch.Tensor] = None
        self._found_inf: Optional[torch.Tensor] = None

    def _gather_param_grads(self, arena, n):
        """Stock-autograd route: gradients were handed to param.grad; copy them into the arena and derive the
        activity bitmap from which parameters have one."""
        names = []
        for name, (off, numel, _) in arena.slots.items():
            if off >= n:
                break
            p = arena.params[name]
            if p.grad is not None:
                arena.grads[off:off + numel].copy_(p.grad.reshape(-1))
                names.append(name)
        key = tuple(names)
        cache = getattr(self, "_bitmap_cache", None)
        if cache is None or cache[0] != key:
            self._bitmap_cache = (key, arena.active_bitmap(names))
        return self._bitmap_cache[1]

    @torch.no_grad()
    def step(self, closure=None, grad_scaler=None):
        arena = self.model.arena
        n = arena.n_hot
        if self._m is None:
            self._m = torch.zeros(n, dtype=torch.float32, device=arena.device)
            self._v = torch.zeros(n, dtype=torch.float32, device=arena.device)
        g = self.param_groups[0]
        if self.model.direct_grads:
            active = self.model._last_active
        else:
            active = self._gather_param_grads(arena, n)
        inv_scale = found_inf = None
        if grad_scaler is not None and grad_scaler.is_enabled():
            # GradScaler.unscale_ + inf check + conditional step (traintest_cavmae_base.py:138-140), on device
            scale = grad_scaler._get_scale_async()
            inv_scale = scale.double().reciprocal().float().reshape(1)
            if self._found_inf is None:
                self._found_inf = torch.zeros(1, dtype=torch.float32, device=arena.device)
            found_inf = self._found_inf
            found_inf.zero_()
            ops.found_inf(arena.grads[:n], found_inf)
            grad_scaler._per_optimizer_states[id(self)]["found_inf_per_device"] = {arena.device: found_inf}
        self._step += 1   # (a skipped step keeps torch's per-parameter counters unchanged; the difference is one
        #                    bias-correction tick after an overflow and vanishes with the GradScaler warm-up)
        ops.adam_step(arena.flat[:n], arena.grads[:n], self._m, self._v, arena.shadow[:n], g["lr"], g["betas"][0],
                      g["betas"][1], g["eps"], g["weight_decay"], self._step, self.decoupled, inv_scale, found_inf,
                      active)
        arena.shadow_fresh = True
        return None

    def state_dict(self):
        sd = super().state_dict()
        sd["fused"] = {"step": self._step, "m": self._m, "v": self._v}
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        fused = sd.pop("fused", None)
        super().load_state_dict(sd)
        if fused is not None:
            self._step, self._m, self._v = fused["step"], fused["m"], fused["v"]
