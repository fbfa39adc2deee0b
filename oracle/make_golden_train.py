"""TEST INFRASTRUCTURE ONLY — golden trace of the reference's OWN training loop.

    python -m oracle.make_golden_train        (only works where /root/reference exists)

Runs the UNMODIFIED `train()` of src/traintest_cavmae_base.py (:29-262) for one epoch of two iterations on the CPU with
the unmodified reference model (oracle/ref_shim.py): DDP wrap (:58-59), the two Adam optimizers over the same parameter
list (:62-66), GradScaler + autocast (:83-84,131-152), the contrastive pass stepping `optimizer` and the MAE pass
stepping `optimizer2`, then validate() (:381-424). What cannot exist in this container is replaced from the OUTSIDE
(no reference line is edited): the module-level names `DDP` (a pass-through wrapper: single process) and `torch` (a proxy
whose `device()` answers "cpu"; the loop hard-codes torch.device("cuda", local_rank) at :57), `linear_val` (the MLP probe
that follows an epoch: a different model, not on this path), and the model's random draws (MaskPlan per forward call,
as in oracle/make_golden.py). On the CPU GradScaler / autocast disable themselves, i.e. the trace is the fp32
unscaled sequence the scaled GPU loop must reproduce (loss scales are powers of two).

tests/golden/train_loop.pt records every forward's 8 outputs, validate()'s return values and the final values of a few
small parameters; tests/test_train_loop_gpu.py replays the same body through avsiam_b200.patch() on the B200.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import avsiam_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import _Injector  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PLAN_SEED0 = 7000            # forward call k (0-based, training and validation alike) uses MaskPlan seed PLAN_SEED0 + k
BATCH = 2
TRACK = ["vit_base.blocks.0.norm1_a.weight", "vit_base.blocks.11.attn.proj.bias", "vit_base.norm.weight",
         "vit_base.patch_embed_a.proj.bias", "mm_layer_1.norm1_a.weight", "decoder_embed.bias", "mask_token",
         "decoder_blocks.7.attn.qkv.bias", "decoder_norm.bias", "decoder_pred_a.bias", "vit_base.blocks.5.mlp.fc2.bias"]


def loop_args(exp_dir):
    """The fields train() / validate() read (run_cavmae_pretrain_base.py argparse defaults where they matter)."""
    return types.SimpleNamespace(
        exp_dir=exp_dir, local_rank=0, rank=1, lr=2e-4, lr_adapt=False, lr_patience=2, lrscheduler_start=10,
        lrscheduler_step=5, lrscheduler_decay=0.5, dataset="audioset", n_epochs=1, data_train="synthetic",
        label_csv=None, batch_size=BATCH, num_workers=0, masking_ratio_a=0.75, masking_ratio=0.75,
        mask_mode="unstructured", n_print_steps=1, mae_loss_weight=1.0, contrast_loss_weight=0.01, wandb=False,
        save_model=False, n_class=527)


class PlanFeeder(nn.Module):
    """Wraps the reference model: every forward consumes the MaskPlan of its call index and is recorded."""

    def __init__(self, model, d):
        super().__init__()
        self.model, self.d, self.calls, self.trace = model, d, 0, []

    def forward(self, audio, imgs, ra, rv, mae_loss_weight=1.0, contrast_loss_weight=0.01, mask_mode="unstructured"):
        B = audio.shape[0]
        plan = O.make_mask_plan(B, self.d, PLAN_SEED0 + self.calls, two_pass=True)
        sorts, perms = [], []
        if mae_loss_weight != 0:
            perms += [None, None]
            sorts += [plan.ids_shuffle_a, plan.ids_shuffle_v]
        if contrast_loss_weight != 0:
            perms += [plan.perm_a, plan.perm_v]
            for ia, iv in zip(plan.chunk_ids_a, plan.chunk_ids_v):
                sorts += [ia, iv]
        with _Injector(sorts, perms):
            out = self.model(audio, imgs, ra, rv, mae_loss_weight=mae_loss_weight,
                             contrast_loss_weight=contrast_loss_weight, mask_mode=mask_mode)
        self.trace.append({"call": self.calls, "B": B, "mae_w": float(mae_loss_weight), "c_w": float(contrast_loss_weight),
                           "training": self.model.training,
                           "out": [float(o) if o is not None else None for o in (out[0], out[1], out[2], out[3], out[4], out[7])]})
        self.calls += 1
        return out


class PassDDP(nn.Module):
    """Single-process stand-in for DistributedDataParallel: `.module`, same call signature."""

    def __init__(self, module, device_ids=None, output_device=None, find_unused_parameters=False, **_):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


class CpuTorch:
    """The `torch` the loop sees: everything passes through, but torch.device(...) always answers the CPU."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref_shim.ensure_process_group()
    tt = ref_shim.load_traintest()
    d = O.VIT_B
    model = ref_shim.build_reference_model()
    sd0 = O.with_aliases(O.init_state(d, seed=0))
    model.load_state_dict(sd0, strict=True)
    feeder = PlanFeeder(model, d)
    tt.DDP, tt.torch, tt.linear_val = PassDDP, CpuTorch(), (lambda *a, **k: None)
    ds = ref_shim.SyntheticAVDataset()
    val_loader = torch.utils.data.DataLoader(torch.utils.data.Subset(ds, [0, 1]), batch_size=BATCH, shuffle=False)
    with tempfile.TemporaryDirectory() as exp_dir:
        os.makedirs(os.path.join(exp_dir, "models"))
        args = loop_args(exp_dir)
        tt.train(feeder, None, (val_loader, None), (None, None), None, args, {})
    named = {k.replace("my_blocks.", "vit_base.blocks."): p for k, p in model.named_parameters()}
    rec = {
        "plan_seed0": PLAN_SEED0, "batch": BATCH, "dataset_seed": ds.seed, "n_samples": ds.n_samples, "weights_seed": 0,
        "adam": {"lr": 2e-4, "weight_decay": 5e-7, "betas": (0.95, 0.999)},
        "trace": feeder.trace,
        "final": {k: named[k].detach().clone() for k in TRACK},
        "initial": {k: sd0[k].clone() for k in TRACK},
    }
    for t in feeder.trace:
        print("[golden]", t)
    torch.save(rec, os.path.join(GOLDEN_DIR, "train_loop.pt"))
    print("[golden] wrote train_loop.pt:", len(feeder.trace), "forward calls")


if __name__ == "__main__":
    main()
