"""TEST INFRASTRUCTURE ONLY — golden vectors for the finetune model, produced by executing the UNMODIFIED reference
class CAVMAEFT_BASE (src/models/cav_mae_base.py:745-1036) through oracle/ref_shim.py.

    python -m oracle.make_golden_ft            (only works where /root/reference exists)

Stores, for seeded weights (O.init_ft_state) and seeded inputs: the state_dict layout, the logits of the three
training modes, and for a BCE-with-logits objective (traintest_ft_base.py: loss_fn = nn.BCEWithLogitsLoss on the
three heads of 'mm_grad') every parameter's gradient norm + random projection and a few full gradients.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import avsiam_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import GOLDEN_DIR, grad_summary  # noqa: E402

LABEL_DIM = 527   # AudioSet (BASELINE.json config 3)
FT_FULL_GRAD_KEYS = [
    "vit_base.blocks.0.norm1_a.weight", "vit_base.blocks.0.norm1_v.bias", "vit_base.blocks.11.attn.proj.bias",
    "vit_base.blocks.5.mlp.fc2.bias", "vit_base.norm.weight", "vit_base.norm_a.bias",
    "vit_base.patch_embed_a.proj.bias", "mm_layer_1.norm1_a.weight", "mm_layer_2.mlp.fc2.bias",
    "mlp_head.0.weight", "mlp_head_a.1.bias", "mlp_head_mm.0.bias", "mlp_head_mm.1.bias",
]


def synth_ft_inputs(B: int, T: int, d: O.Dims, seed: int, label_dim: int = LABEL_DIM):
    g = torch.Generator().manual_seed(seed)
    audio = torch.randn(B, d.audio_len, d.mel, generator=g)
    video = torch.randn(B, T, d.in_chans, d.img, d.img, generator=g)
    labels = (torch.rand(B, label_dim, generator=g) < 0.02).float()   # sparse multi-label targets
    return audio, video, labels


def ft_loss(outs, labels):
    """BCE-with-logits summed over the heads the mode returns (traintest_ft_base.py:150-165 for 'mm_grad')."""
    if not isinstance(outs, (tuple, list)):
        outs = (outs,)
    return sum(torch.nn.functional.binary_cross_entropy_with_logits(o, labels) for o in outs)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    d = O.VIT_B
    ref = ref_shim.load_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref.CAVMAEFT_BASE(label_dim=LABEL_DIM, audio_length=1024, modality_specific_depth=23, tr_pos=False)
    layout = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(GOLDEN_DIR, "ft_state_dict_layout.json"), "w") as f:
        json.dump(layout, f, indent=0, sort_keys=True)
    shapes = O.ft_param_shapes(d, LABEL_DIM)
    want = set(shapes) | {k.replace("vit_base.blocks.", "my_blocks.") for k in shapes if k.startswith("vit_base.blocks.")}
    assert set(layout) == want, sorted(set(layout) ^ want)[:10]
    for k, s in shapes.items():
        assert tuple(layout[k]) == tuple(s), (k, layout[k], s)
    sd = O.with_aliases(O.init_ft_state(d, LABEL_DIM, seed=0))
    model.load_state_dict(sd, strict=True)
    model.train()
    cases = []
    for name, mode, B, T, seed in (("mm_grad_B2", "mm_grad", 2, 1, 301), ("audioonly_B3", "audioonly", 3, 1, 302),
                                    ("videoonly_B2_T2", "videoonly", 2, 2, 303)):
        audio, video, labels = synth_ft_inputs(B, T, d, seed)
        model.zero_grad(set_to_none=True)
        outs = model(audio, video, mode)
        lab = labels if mode != "videoonly" or T == 1 else labels.unsqueeze(1).expand(-1, T, -1)
        loss = ft_loss(outs, lab)
        loss.backward()
        named = {k.replace("my_blocks.", "vit_base.blocks."): p.grad for k, p in model.named_parameters()}
        norms, projs, _ = grad_summary(named)
        full = {k: named[k].detach().clone() for k in FT_FULL_GRAD_KEYS if named.get(k) is not None}
        outs_t = outs if isinstance(outs, (tuple, list)) else (outs,)
        cases.append({"name": name, "mode": mode, "B": B, "T": T, "seed": seed, "label_dim": LABEL_DIM,
                      "logits": [o.detach().clone() for o in outs_t], "loss": float(loss),
                      "grad_norm": norms, "grad_proj": projs, "grad_full": full})
        print(f"[golden-ft] {name}: loss={float(loss):.6f} n_grads={len(norms)} "
              f"logit0 mean={float(outs_t[0].mean()):.5f}")
    torch.save(cases, os.path.join(GOLDEN_DIR, "cavmaeft_base_forward.pt"))
    print("[golden-ft] wrote", GOLDEN_DIR)


if __name__ == "__main__":
    main()
