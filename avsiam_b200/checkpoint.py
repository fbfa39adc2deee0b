"""Checkpoint interchange with the reference — host-side logic only.

  * save_model / save_optimizer: what the training loops write (traintest_cavmae_base.py:227-234): the state dict of
    the DDP / DataParallel wrapper, i.e. every key prefixed with `module.`, plus `best_optim_state.pth`.
  * load_pretrained: the PT -> FT transfer of run_cavmae_ft_base.py:245-258 (`strict=False` load of a `module.`-prefixed
    pretraining checkpoint into CAVMAEFT_BASE; returns the same (missing, unexpected) key lists, prefixed like the
    reference prints them), optionally followed by `__create_fusion__()`.
  * wa_model: checkpoint weight averaging of run_cavmae_ft_base.py:169-180 (plain mean over epochs, not an ensemble).
The `my_blocks.*` keys are aliases of `vit_base.blocks.*` (cav_mae_base.py:278): they are saved twice like the
reference does and tolerated in any order on load (both names bind the same tensor).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

PREFIX = "module."


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def prefixed_state_dict(model) -> "OrderedDict[str, torch.Tensor]":
    """state_dict() as torch.nn.parallel.DistributedDataParallel / DataParallel would return it."""
    sd = _unwrap(model).state_dict()
    return OrderedDict((PREFIX + k, v) for k, v in sd.items())


def strip_prefix(sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k[len(PREFIX):] if k.startswith(PREFIX) else k, v) for k, v in sd.items())


def save_model(model, path: str) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(OrderedDict((k, v.detach().cpu()) for k, v in prefixed_state_dict(model).items()), path)


def save_optimizer(optimizer, path: str) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(optimizer.state_dict(), path)


def load_model(model, path_or_sd, strict: bool = True) -> Tuple[List[str], List[str]]:
    """Loads a checkpoint written by the reference or by save_model (with or without the `module.` prefix)."""
    sd = torch.load(path_or_sd, map_location="cpu") if isinstance(path_or_sd, (str, os.PathLike)) else path_or_sd
    res = _unwrap(model).load_state_dict(strip_prefix(sd), strict=strict)
    return list(res.missing_keys), list(res.unexpected_keys)


def load_pretrained(ft_model, path_or_sd, create_fusion: bool = False) -> Tuple[List[str], List[str]]:
    """run_cavmae_ft_base.py:245-258: `audio_model = DataParallel(audio_model); load_state_dict(mdl_weight,
    strict=False)`.  Returns (missing, unexpected) with the `module.` prefix the reference prints."""
    missing, unexpected = load_model(ft_model, path_or_sd, strict=False)
    if create_fusion:
        _unwrap(ft_model).__create_fusion__()
    return [PREFIX + k for k in missing], [PREFIX + k for k in unexpected]


def wa_model(exp_dir: str, start_epoch: int, end_epoch: int) -> "OrderedDict[str, torch.Tensor]":
    """Average of exp_dir/models/audio_model.{start..end}.pth (run_cavmae_ft_base.py:169-180)."""
    sd_a = torch.load(os.path.join(exp_dir, "models", f"audio_model.{start_epoch}.pth"), map_location="cpu")
    acc = OrderedDict((k, v.clone().to(torch.float64) if v.is_floating_point() else v.clone()) for k, v in sd_a.items())
    count = 1
    for epoch in range(start_epoch + 1, end_epoch + 1):
        sd_b = torch.load(os.path.join(exp_dir, "models", f"audio_model.{epoch}.pth"), map_location="cpu")
        for k in acc:
            acc[k] = acc[k] + sd_b[k]
        count += 1
    out = OrderedDict()
    for k, v in acc.items():
        out[k] = (v / float(count)).to(sd_a[k].dtype) if sd_a[k].is_floating_point() else v // count
    return out
