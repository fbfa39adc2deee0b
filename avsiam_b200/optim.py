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
    """torch.optim.Adam over the model's flat arena. param_groups are honoured: each group's lr / weight_decay apply to
    exactly the parameters it lists (the finetune loop's three groups, traintest_ft_base.py:78-83); parameters handed
    to no group (frozen backbones, `freeze_base`) are neither stepped nor decayed, like torch."""
    _step_supports_amp_scaling = True   # GradScaler.step hands us the scaler; unscale + inf-skip run in the kernel
    MAX_GROUPS = 8

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
        if len(self.param_groups) > self.MAX_GROUPS:
            raise ValueError(f"FusedAdam supports at most {self.MAX_GROUPS} param_groups")
        self.decoupled = decoupled
        self.capturable = False    # True: the step counter / bias corrections always live on the device (CUDA graphs)
        self._step = 0
        self._m: Optional[torch.Tensor] = None
        self._v: Optional[torch.Tensor] = None
        self._found_inf: Optional[torch.Tensor] = None
        self._tick: Optional[torch.Tensor] = None      # device {int step; float bc1; float bc2_sqrt} (GradScaler route)
        self._group_map = None                         # (signature, uint8 per chunk: 0 = not ours, k = group k-1)
        self._chunk_cache = {}

    # ------------------------------------------------------------------------------------------ group / activity maps
    def _group_chunks(self, arena) -> torch.Tensor:
        """uint8 per 64-element chunk of the hot prefix: 0 = parameter not handed to this optimizer, k = group k-1."""
        sig = tuple(tuple(id(p) for p in g["params"]) for g in self.param_groups)
        if self._group_map is not None and self._group_map[0] == sig:
            return self._group_map[1]
        name_of = {id(p): n for n, p in arena.params.items()}
        gm = torch.zeros(arena.n_hot // arena.ALIGN, dtype=torch.uint8)
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                n = name_of.get(id(p))
                if n is None:
                    raise ValueError("FusedAdam: a parameter of group %d does not belong to the model's arena" % gi)
                off, numel, _ = arena.slots[n]
                if off >= arena.n_hot:
                    continue                           # cold tail (heads / unused copies): never receives a gradient
                gm[off // arena.ALIGN:(off + numel + arena.ALIGN - 1) // arena.ALIGN] = gi + 1
        self._group_map = (sig, gm.to(arena.device))
        self._chunk_cache = {}
        return self._group_map[1]

    def _step_chunks(self, arena, active: Optional[torch.Tensor]) -> torch.Tensor:
        """Group map restricted to the chunks that received a gradient in the last backward."""
        gm = self._group_chunks(arena)
        if active is None:
            return gm
        key = (active.data_ptr(), active.numel())
        hit = self._chunk_cache.get(key)
        if hit is None or hit[0] is not active:
            hit = (active, torch.where(active != 0, gm, torch.zeros_like(gm)))
            self._chunk_cache[key] = hit
        return hit[1]

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

    def _ensure_state(self, arena):
        n = arena.n_hot
        if self._m is None:
            self._m = torch.zeros(n, dtype=torch.float32, device=arena.device)
            self._v = torch.zeros(n, dtype=torch.float32, device=arena.device)
        return n

    @torch.no_grad()
    def step(self, closure=None, grad_scaler=None):
        arena = self.model.arena
        n = self._ensure_state(arena)
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:
            if tuple(g["betas"]) != tuple(g0["betas"]) or g["eps"] != g0["eps"]:
                raise ValueError("FusedAdam: betas / eps must be the same in every param_group (lr and weight_decay may differ)")
        if self.model.direct_grads:
            active = self.model._last_active
        else:
            active = self._gather_param_grads(arena, n)
        chunks = self._step_chunks(arena, active)
        inv_scale = found_inf = tick = None
        if grad_scaler is not None and grad_scaler.is_enabled():
            # GradScaler.unscale_ + inf check + conditional step (traintest_cavmae_base.py:138-140), on device; the
            # step counter advances on the device only when the step is taken (torch's per-parameter `step` does too)
            scale = grad_scaler._get_scale_async()
            inv_scale = scale.double().reciprocal().float().reshape(1)
            if self._found_inf is None:
                self._found_inf = torch.zeros(1, dtype=torch.float32, device=arena.device)
            found_inf = self._found_inf
            found_inf.zero_()
            ops.found_inf(arena.grads[:n], found_inf)
            grad_scaler._per_optimizer_states[id(self)]["found_inf_per_device"] = {arena.device: found_inf}
            if self._tick is None:
                self._tick = torch.zeros(4, dtype=torch.int32, device=arena.device)
                self._tick[0] = self._step
            tick = self._tick
        elif self.capturable:
            if self._tick is None:
                self._tick = torch.zeros(4, dtype=torch.int32, device=arena.device)
                self._tick[0] = self._step
            tick = self._tick
        elif self._tick is not None:                     # scaler was switched off: fold the device counter back
            self._step = int(self._tick[0])
            self._tick = None
        self._step += 1   # host-side count of step() calls (exact step count when no GradScaler skips occurred)
        ops.adam_step(arena.flat[:n], arena.grads[:n], self._m, self._v, arena.shadow[:n],
                      [g["lr"] for g in self.param_groups], g0["betas"][0], g0["betas"][1], g0["eps"],
                      [g["weight_decay"] for g in self.param_groups], self._step, self.decoupled, inv_scale, found_inf,
                      chunks, tick)
        arena.shadow_fresh = True
        return None

    # ------------------------------------------------------------------------------------------ checkpoint interchange
    def steps_taken(self) -> int:
        """Optimizer steps actually applied (GradScaler-skipped ones excluded; reads the device counter)."""
        return int(self._tick[0]) if self._tick is not None else self._step

    def state_dict(self):
        """torch.optim.Adam's layout (`best_optim_state.pth`, traintest_cavmae_base.py:230): per-parameter `step`,
        `exp_avg`, `exp_avg_sq` for every parameter that has been stepped — a torch.optim.Adam over the same parameter
        list loads it unchanged, and load_state_dict accepts torch's."""
        sd = super().state_dict()
        if self._m is not None:
            arena = self.model.arena
            name_of = {id(p): n for n, p in arena.params.items()}
            steps = self.steps_taken()
            chunks = self._group_chunks(arena).cpu()
            state, idx = {}, 0
            for g in self.param_groups:
                for p in g["params"]:
                    n = name_of.get(id(p))
                    if n is not None:
                        off, numel, shape = arena.slots[n]
                        if off < arena.n_hot and chunks[off // arena.ALIGN] != 0 and bool(self._v[off:off + numel].any()):
                            state[idx] = {"step": torch.tensor(float(steps)),
                                          "exp_avg": self._m[off:off + numel].view(shape).clone(),
                                          "exp_avg_sq": self._v[off:off + numel].view(shape).clone()}
                    idx += 1
            sd["state"] = state
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        sd.pop("fused", None)                              # round-1 private layout: superseded by torch's
        state = sd.get("state", {})
        super().load_state_dict({"state": {}, "param_groups": sd["param_groups"]})
        arena = self.model.arena
        self._ensure_state(arena)
        self._m.zero_(); self._v.zero_()
        name_of = {id(p): n for n, p in arena.params.items()}
        flat_params = [p for g in self.param_groups for p in g["params"]]
        steps = 0
        for idx, st in state.items():
            p = flat_params[int(idx)]
            n = name_of[id(p)]
            off, numel, shape = arena.slots[n]
            if tuple(st["exp_avg"].shape) != tuple(shape):
                raise ValueError(f"FusedAdam.load_state_dict: {n}: moment shape {tuple(st['exp_avg'].shape)} != {tuple(shape)}")
            if off >= arena.n_hot:
                continue
            self._m[off:off + numel].copy_(st["exp_avg"].reshape(-1).to(arena.device, torch.float32))
            self._v[off:off + numel].copy_(st["exp_avg_sq"].reshape(-1).to(arena.device, torch.float32))
            steps = max(steps, int(float(st["step"])))
        self._step = steps
        self._tick = None
