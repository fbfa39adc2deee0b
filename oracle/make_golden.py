"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.pt by executing the UNMODIFIED reference model file.

    python -m oracle.make_golden            (only works where /root/reference exists)

The reference's random draws are replaced, from the outside, by the MaskPlan the oracle and the CUDA path also
consume: `torch.argsort` on a floating tensor (the noise sort at cav_mae_base.py:377,426) returns the supplied
ids_shuffle, `torch.randperm` (:465,469,533,537) returns the supplied permutations. No reference source line is
edited or copied.  Weights and inputs are regenerated from seeds, so the fixtures stay small: they hold the
outputs, the masks, per-parameter gradient norms / random projections and a few full gradients.
"""
from __future__ import annotations

import json
import os
import sys
import zlib

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import avsiam_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

FULL_GRAD_KEYS = [
    "vit_base.blocks.0.norm1_a.weight", "vit_base.blocks.0.norm1_v.bias", "vit_base.blocks.11.attn.proj.bias",
    "vit_base.blocks.5.mlp.fc2.bias", "vit_base.norm.weight", "vit_base.norm_a.bias",
    "vit_base.patch_embed_a.proj.bias", "ast_base.blocks.3.norm2.weight", "ast_base.norm_a.weight",
    "mm_layer_1.norm1_a.weight", "mm_layer_2.mlp.fc2.bias", "decoder_embed.bias", "mask_token",
    "decoder_modality_a", "decoder_blocks.7.attn.qkv.bias", "decoder_norm.bias", "decoder_pred_a.bias",
]


def synth_inputs(B: int, d: O.Dims, seed: int):
    """Synthetic AudioSet-shaped batch: randn fbank [B,1024,128] + randn frame [B,3,224,224]."""
    g = torch.Generator().manual_seed(seed)
    audio = torch.randn(B, d.audio_len, d.mel, generator=g)
    imgs = torch.randn(B, d.in_chans, d.img, d.img, generator=g)
    return audio, imgs


def proj_vector(key: str, numel: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(zlib.crc32(("proj:" + key).encode()) % (2**31))
    return torch.randn(numel, generator=g)


class _Injector:
    """Context manager feeding supplied indices to the reference's RNG-dependent calls."""

    def __init__(self, sorts, perms):
        self.sorts, self.perms = list(sorts), list(perms)

    def __enter__(self):
        self._argsort, self._randperm = torch.argsort, torch.randperm
        inj = self

        def argsort(x, *a, **k):
            if x.is_floating_point():
                ids = inj.sorts.pop(0)
                assert ids.shape == x.shape, (ids.shape, x.shape)
                return ids.clone()
            return inj._argsort(x, *a, **k)

        def randperm(n, *a, **k):
            p = inj.perms.pop(0)
            return torch.arange(n) if p is None else p.clone()

        torch.argsort, torch.randperm = argsort, randperm
        return self

    def __exit__(self, *exc):
        torch.argsort, torch.randperm = self._argsort, self._randperm
        assert not self.sorts and not self.perms, "unused injected draws"


def run_reference(model, audio, imgs, plan: O.MaskPlan, mae_w, c_w):
    sorts, perms = [], []
    if mae_w != 0:
        perms += [None, None]                                  # the two unused randperm at :465,:469
        sorts += [plan.ids_shuffle_a, plan.ids_shuffle_v]
    if c_w != 0:
        perms += [plan.perm_a, plan.perm_v]
        for ia, iv in zip(plan.chunk_ids_a, plan.chunk_ids_v):
            sorts += [ia, iv]
    model.zero_grad(set_to_none=True)
    with _Injector(sorts, perms):
        out = model(audio, imgs, 0.75, 0.75, mae_loss_weight=mae_w, contrast_loss_weight=c_w,
                    mask_mode="unstructured")
    return out


def grad_summary(named_grads):
    norms, projs, full = {}, {}, {}
    for k, g in named_grads.items():
        if g is None:
            continue
        g = g.detach().double().flatten()
        norms[k] = float(g.norm())
        projs[k] = float(g @ proj_vector(k, g.numel()).double())
    for k in FULL_GRAD_KEYS:
        if named_grads.get(k) is not None:
            full[k] = named_grads[k].detach().clone()
    return norms, projs, full


def make_case(model, d, name, B, mae_w, c_w, seed_in, seed_mask):
    audio, imgs = synth_inputs(B, d, seed_in)
    plan = O.make_mask_plan(B, d, seed_mask, two_pass=True)
    out = run_reference(model, audio, imgs, plan, mae_w, c_w)
    loss = out[0]
    loss.backward()
    # unique parameters under their canonical (non-alias) names
    named = {}
    for k, p in model.named_parameters():  # named_parameters() de-duplicates aliases; first name wins
        named[k.replace("my_blocks.", "vit_base.blocks.")] = p.grad
    norms, projs, full = grad_summary(named)
    rec = {
        "name": name, "B": B, "mae_w": mae_w, "c_w": c_w, "seed_in": seed_in, "seed_mask": seed_mask,
        "weights_seed": 0,
        "loss": float(out[0]), "loss_mae": float(out[1]), "loss_mae_a": float(out[2]), "loss_mae_v": float(out[3]),
        "loss_c": float(out[4]), "c_acc": float(out[7]),
        "mask_a": None if out[5] is None else out[5].to(torch.uint8),
        "mask_v": None if out[6] is None else out[6].to(torch.uint8),
        "grad_norm": norms, "grad_proj": projs, "grad_full": full,
        "input_checksum": float(audio.double().sum() + imgs.double().sum()),
    }
    print(f"[golden] {name}: loss={rec['loss']:.6f} mae_a={rec['loss_mae_a']:.6f} mae_v={rec['loss_mae_v']:.6f} "
          f"c={rec['loss_c']:.6f} acc={rec['c_acc']:.3f} n_grads={len(norms)}")
    return rec


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref_shim.ensure_process_group()
    d = O.VIT_B
    model = ref_shim.build_reference_model()
    ref_sd = model.state_dict()
    # 1. checkpoint layout
    layout = {k: list(v.shape) for k, v in ref_sd.items()}
    with open(os.path.join(GOLDEN_DIR, "state_dict_layout.json"), "w") as f:
        json.dump(layout, f, indent=0, sort_keys=True)
    assert set(layout) == set(O.state_dict_keys(d)), set(layout) ^ set(O.state_dict_keys(d))
    for k, s in O.param_shapes(d).items():
        assert tuple(layout[k]) == tuple(s), (k, layout[k], s)
    print(f"[golden] layout: {len(layout)} keys, "
          f"{sum(p.numel() for p in model.parameters()) / 1e6:.2f} M unique params")
    # 2. load the seeded weights the oracle / CUDA path will regenerate
    sd = O.with_aliases(O.init_state(d, seed=0))
    missing, unexpected = model.load_state_dict(sd, strict=True)
    model.train()
    # 3. masking primitive, bit-exact fixture (random_masking_unstructured run by the reference itself)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 40, 16, generator=g)
    ids = O.stable_argsort(torch.rand(3, 40, generator=g))
    with _Injector([ids], []):
        xm, mask, ids_restore = model.random_masking_unstructured(x, 0.75)
    torch.save({"x": x, "ids_shuffle": ids, "x_masked": xm, "mask": mask, "ids_restore": ids_restore,
                "mask_ratio": 0.75}, os.path.join(GOLDEN_DIR, "masking_unstructured.pt"))
    # 4. whole-forward cases
    cases = [
        make_case(model, d, "pass2_mae_B2", 2, 1, 0, seed_in=87, seed_mask=1234),
        make_case(model, d, "pass1_contrastive_B5", 5, 0, 1, seed_in=88, seed_mask=1235),
        make_case(model, d, "both_B2", 2, 1.0, 0.01, seed_in=89, seed_mask=1236),
    ]
    torch.save(cases, os.path.join(GOLDEN_DIR, "cavmae_base_forward.pt"))
    print("[golden] wrote", GOLDEN_DIR)


if __name__ == "__main__":
    main()
