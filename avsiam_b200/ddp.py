"""Data-parallel gradient exchange for the flat parameter arena (reference: torch DDP wrap at
src/traintest_cavmae_base.py:58-59 — `DDP(audio_model, device_ids=[gpu], find_unused_parameters=True)`).

The reference all-reduces all 248 M parameters (unused ones included) in 25 MB buckets discovered by autograd
hooks. Here the gradients already sit in ONE contiguous fp32 buffer written by the hand-written reverse pass, so
the exchange is planned statically per step shape:

  * `plan_buckets` cuts the used part of the arena into contiguous buckets and computes, for each, the index of
    the last reverse-pass closure that writes into it;
  * `GradSync.after_closure(i)` enqueues the NCCL all-reduce (average) of every bucket that became final at
    closure i — torch.distributed runs it on the process group's own stream, ordered after the compute stream by
    an event, so it overlaps the remaining backward kernels over NVLink 5 / NVSwitch;
  * `finish()` joins the outstanding work before the optimizer reads the gradients.

Parameters that received no gradient are neither communicated nor stepped (their gradient is identically zero
on every rank), which is what `find_unused_parameters=True` achieves in the reference at 2x the traffic.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

DEFAULT_BUCKET_ELEMS = 16 * 1024 * 1024   # 64 MB of fp32 gradients per all-reduce


def plan_buckets(slots: "Dict[str, Tuple[int, int, torch.Size]]", used: Sequence[str],
                 touches: Sequence[Sequence[str]], bucket_elems: int = DEFAULT_BUCKET_ELEMS,
                 align: int = 64, gap_max: int = 65536) -> List[Tuple[int, int, int]]:
    """-> [(start, end, ready_at)] over arena element offsets.

    slots: arena layout name -> (offset, numel, shape), in arena order. used: names that receive gradient this
    step. touches[i]: name prefixes written by the i-th closure of the reverse pass IN EXECUTION ORDER.
    A bucket is final after closure `ready_at` (= the latest closure touching any of its parameters).
    Buckets never span a LARGE unused parameter (e.g. the whole `ast_base` copy in the single-pass arrangement);
    unused tensors of <= gap_max elements (the spare LayerNorm sets inside a block) are bridged so a block stays
    one message."""
    used_set = set(used)
    last: Dict[str, int] = {}
    for i, prefixes in enumerate(touches):
        for n in used_set:
            for pf in prefixes:
                if n.startswith(pf):
                    last[n] = i
                    break
    buckets: List[Tuple[int, int, int]] = []
    cur: Optional[List[int]] = None   # [start, end, ready]
    for n, (off, numel, _) in slots.items():
        if n not in used_set:
            if cur is not None:
                if numel <= gap_max and cur[1] == off:   # bridge a small unused tensor (its gradient is zero)
                    cur[1] = (off + numel + align - 1) // align * align
                else:
                    buckets.append(tuple(cur)); cur = None
            continue
        end = (off + numel + align - 1) // align * align
        ready = last.get(n, len(touches) - 1)   # unknown writer: conservatively final only at the very end
        if cur is not None and (cur[1] != off or (cur[1] - cur[0]) + (end - off) > bucket_elems):
            buckets.append(tuple(cur)); cur = None
        if cur is None:
            cur = [off, end, ready]
        else:
            cur[1] = end
            cur[2] = max(cur[2], ready)
    if cur is not None:
        buckets.append(tuple(cur))
    return buckets


class GradSync:
    """Bucketed, backward-overlapped gradient averaging over the arena's flat gradient buffer."""

    def __init__(self, process_group=None, bucket_elems: int = DEFAULT_BUCKET_ELEMS):
        self.group = process_group
        self.bucket_elems = bucket_elems
        self._plans: Dict[tuple, List[Tuple[int, int, int]]] = {}
        self._by_ready: Dict[int, List[Tuple[int, int]]] = {}
        self._work: list = []
        self._grads: Optional[torch.Tensor] = None
        self.bytes_last_step = 0

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def begin(self, key: tuple, arena, used: Sequence[str], touches: Sequence[Sequence[str]]) -> None:
        plan = self._plans.get(key)
        if plan is None:
            plan = plan_buckets(arena.slots, used, touches, self.bucket_elems, arena.ALIGN)
            self._plans[key] = plan
        self._by_ready = {}
        for (s, e, r) in plan:
            self._by_ready.setdefault(r, []).append((s, e))
        self._grads = arena.grads
        self._work = []
        self.bytes_last_step = 0

    def after_closure(self, i: int) -> None:
        for (s, e) in self._by_ready.pop(i, ()):  # buckets whose last writer just ran
            self._launch(s, e)

    def _launch(self, s: int, e: int) -> None:
        if self.world == 1:
            return
        view = self._grads[s:e]
        if dist.get_backend(self.group) == "nccl":
            w = dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._work.append((w, None))
        else:                                       # gloo (CPU tests) has no AVG
            w = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._work.append((w, view))
        self.bytes_last_step += (e - s) * 4

    def finish(self) -> None:
        for r in sorted(self._by_ready):           # anything not released by a closure index (defensive)
            for (s, e) in self._by_ready[r]:
                self._launch(s, e)
        self._by_ready = {}
        for w, view in self._work:
            w.wait()                               # stream-level join: no host sync with NCCL
            if view is not None:
                view.div_(self.world)
        self._work = []


class B200DDP(nn.Module):
    """Stand-in for torch.nn.parallel.DistributedDataParallel around avsiam_b200.CAVMAE_BASE: same `.module`
    attribute and `module.`-prefixed state_dict (checkpoints written at traintest_cavmae_base.py:229,234 stay
    interchangeable), same call signature as used at :59. Gradient averaging is done by GradSync inside the
    model's reverse pass; weights are broadcast from rank 0 at construction like DDP does."""

    def __init__(self, module: nn.Module, device_ids=None, output_device=None, find_unused_parameters: bool = True,
                 process_group=None, bucket_cap_mb: float = 64.0, broadcast_buffers: bool = True, **_ignored):
        super().__init__()
        self.module = module
        if not hasattr(module, "arena"):
            raise TypeError("B200DDP wraps avsiam_b200.CAVMAE_BASE (it synchronises the model's gradient arena)")
        module.process_group = process_group
        module.grad_sync = GradSync(self._gradient_group(process_group), int(bucket_cap_mb * 1024 * 1024 // 4))
        arena = module.arena                        # builds + binds the arena on the module's device
        if dist.is_initialized() and dist.get_world_size(process_group) > 1:
            dist.broadcast(arena.flat, src=dist.get_global_rank(process_group, 0) if process_group else 0,
                           group=process_group)
            arena.shadow_fresh = False

    @staticmethod
    def _gradient_group(process_group):
        """A communicator of its own for the gradient buckets, limited to a few CTAs per all-reduce: the buckets overlap
        the persistent 148-CTA GEMMs of the reverse pass, every SM an NCCL kernel holds is taken from them, and over
        NVLink 5 / NVSwitch a handful of channels already carries the bandwidth the overlap needs
        (AVS_DDP_MAX_CTAS = n > 0 enables it; default 0 = share the caller's group as it is). Measured on 8 x B200
        (profiles/r02_bench_8gpu*.log): 8 CTAs 95.3 ms / step against 94.4 ms with NCCL's own choice — the limit costs more
        all-reduce time at the tail of the reverse pass than it returns to the GEMMs, so it is off by default."""
        import os
        max_ctas = int(os.environ.get("AVS_DDP_MAX_CTAS", "0"))
        if (not dist.is_initialized() or dist.get_world_size(process_group) == 1 or max_ctas <= 0
                or dist.get_backend(process_group) != "nccl"):
            return process_group
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = max_ctas
            opts.config.min_ctas = min(max_ctas, 4)
            ranks = dist.get_process_group_ranks(process_group) if process_group is not None else None
            return dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
        except (AttributeError, RuntimeError):     # older NCCL / torch without communicator config: keep the group
            return process_group

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
