"""Developer aid: print the measured relative loss errors / worst gradient cosines of the CUDA path against the
reference golden fixtures (ViT-B/16) — the numbers quoted in DESIGN.md §5."""
import os, sys, dataclasses, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avsiam_b200 import CAVMAE_BASE, Dims
from oracle import avsiam_oracle as O
from oracle.make_golden import synth_inputs
d = O.VIT_B
model = CAVMAE_BASE(dims=Dims(**dataclasses.asdict(d)), arrangement="two_pass")
model.load_state_dict(O.with_aliases(O.init_state(d, seed=0)), strict=True)
model = model.cuda()
cases = torch.load(os.path.join(ROOT, "tests/golden/cavmae_base_forward.pt"), weights_only=False)
names = ("loss", "loss_mae", "loss_mae_a", "loss_mae_v", "loss_c")
for idx, c in enumerate(cases):
    audio, imgs = synth_inputs(c["B"], d, c["seed_in"])
    model.mask_plan = O.make_mask_plan(c["B"], d, c["seed_mask"], two_pass=True)
    model.zero_grad(set_to_none=True)
    out = model(audio.cuda(), imgs.cuda(), 0.75, 0.75, mae_loss_weight=c["mae_w"], contrast_loss_weight=c["c_w"])
    rel = {n: abs(float(out[i]) - c[n]) / max(abs(c[n]), 1e-12) for i, n in enumerate(names) if c[n] != 0}
    out[0].backward()
    named = dict(model.named_parameters())
    cs = [float(torch.nn.functional.cosine_similarity(named[k].grad.flatten().double().cpu(), g.flatten().double(), dim=0))
          for k, g in c["grad_full"].items()]
    print(f"case {idx} (B={c['B']}, mae_w={c['mae_w']}, c_w={c['c_w']}): loss rel err " +
          ", ".join(f"{n}={v:.2e}" for n, v in rel.items()) + f"; worst of {len(cs)} full-gradient cosines {min(cs):.6f}")
