"""avsiam_b200 — B200-native (sm_100a) implementation of the AVSiam / CAV-MAE pretraining hot path.

Public surface mirrors the reference (GenjiB/AVSiam, src/models): `CAVMAE_BASE`, `GatherLayer`; plus the opt-in fast
path `FusedAdam`, `B200DDP`, `patch()`.  Importing the package never touches the GPU; the first kernel call loads
libavsiam_b200.so and raises if it is missing (there is no CPU / PyTorch fallback).
"""
from .cav_mae_base import CAVMAE_BASE, _Dims as Dims, VIT_L_DIMS, VIT_H_DIMS  # noqa: F401
from .cav_mae_ft import CAVMAEFT_BASE  # noqa: F401
from . import augment, checkpoint, evaluate, losses  # noqa: F401
from .fbank import wav2fbank  # noqa: F401
from .stats import calculate_stats, d_prime  # noqa: F401
from .ddp import B200DDP, GradSync  # noqa: F401
from .gather_layer import GatherLayer  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401

__all__ = ["CAVMAE_BASE", "CAVMAEFT_BASE", "Dims", "VIT_L_DIMS", "VIT_H_DIMS", "GatherLayer", "wav2fbank", "calculate_stats", "d_prime", "FusedAdam", "B200DDP", "GradSync", "GraphedTrainStep", "patch"]
__version__ = "0.1.0"


def patch(traintest_module=None, models_module=None):
    """Opt-in rebinding for the UNEDITED reference loop (SURVEY.md §8b "Optimizer / DDP injection").

    `traintest_module` is the imported `traintest_cavmae_base` module: its global `DDP`
    (traintest_cavmae_base.py:59) becomes `B200DDP`, and the `torch.optim.Adam` it constructs at :64-66 becomes
    `FusedAdam` (same positional signature; it finds the model's arena through the parameters it is given).
    `models_module` is the reference's `models` package: its `CAVMAE_BASE` attribute is replaced by this one
    (run_cavmae_pretrain_base.py:175). Returns an `undo()` callable."""
    import torch

    undo = []
    if models_module is not None:
        old = getattr(models_module, "CAVMAE_BASE", None)
        models_module.CAVMAE_BASE = CAVMAE_BASE
        undo.append(lambda: setattr(models_module, "CAVMAE_BASE", old))
        old_ft = getattr(models_module, "CAVMAEFT_BASE", None)
        models_module.CAVMAEFT_BASE = CAVMAEFT_BASE                       # run_cavmae_ft_base.py
        undo.append(lambda: setattr(models_module, "CAVMAEFT_BASE", old_ft))
    if traintest_module is not None:
        old_ddp = getattr(traintest_module, "DDP", None)
        traintest_module.DDP = B200DDP
        undo.append(lambda: setattr(traintest_module, "DDP", old_ddp))

        class _Optim:  # stands in for the `torch.optim` attribute looked up as `torch.optim.Adam`
            def __getattr__(self, name):
                return FusedAdam if name == "Adam" else getattr(torch.optim, name)

        class _Torch:  # module-level `torch` seen by the loop; everything but optim.Adam passes through
            optim = _Optim()

            def __getattr__(self, name):
                return getattr(torch, name)

        old_torch = getattr(traintest_module, "torch", None)
        traintest_module.torch = _Torch()
        undo.append(lambda: setattr(traintest_module, "torch", old_torch))

    def _undo():
        for fn in reversed(undo):
            fn()
    return _undo
