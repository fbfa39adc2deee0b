"""Finetune path parity (B200): the drop-in CAVMAEFT_BASE on libavsiam_b200.so against
  (1) golden fixtures recorded from the UNMODIFIED reference class CAVMAEFT_BASE (tests/golden, ViT-B/16), and
  (2) the CPU oracle executed live (TINY geometry, every parameter's gradient).
Tolerances (bf16 mode): logits within 2e-2 of their scale, BCE loss 2e-2 relative, gradient cosine >= 0.999."""
import dataclasses
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from avsiam_b200 import CAVMAEFT_BASE, Dims, FusedAdam, ops  # noqa: E402
from oracle import avsiam_oracle as O  # noqa: E402
from oracle.make_golden_ft import ft_loss, synth_ft_inputs  # noqa: E402

DEV = "cuda"


def cos(a, b):
    return float(F.cosine_similarity(a.flatten().double().cpu(), b.flatten().double().cpu(), dim=0))


def make_model(d, label_dim, seed=0):
    m = CAVMAEFT_BASE(label_dim=label_dim, dims=Dims(**dataclasses.asdict(d)))
    m.load_state_dict(O.with_aliases(O.init_ft_state(d, label_dim, seed=seed)), strict=True)
    return m.to(DEV)


def labels_for(c_mode, T, labels):
    return labels if c_mode != "videoonly" or T == 1 else labels.unsqueeze(1).expand(-1, T, -1)


def test_head_kernels_match_torch():
    B, D, C = 5, 1536, 527
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, D, generator=g).to(DEV)
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV), (0.1 * torch.randn(D, generator=g)).to(DEV)
    W, b = (torch.randn(C, D, generator=g) * D ** -0.5).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    logits, saved = ops.head_fwd(x, gamma, beta, 1e-5, W, b)
    leaves = [t.clone().requires_grad_(True) for t in (x, gamma, beta, W, b)]
    ref = F.linear(F.layer_norm(leaves[0], (D,), leaves[1], leaves[2], 1e-5), leaves[3], leaves[4])
    assert torch.allclose(logits, ref, rtol=1e-4, atol=1e-4)
    dl = torch.randn(B, C, generator=g).to(DEV)
    ref.backward(dl)
    dW, db = torch.ones_like(W), torch.ones_like(b)
    dg, dbeta = torch.zeros_like(gamma), torch.zeros_like(beta)
    dx = ops.head_bwd(dl, saved, gamma, W, dW, db, dg, dbeta)
    assert torch.allclose(dx, leaves[0].grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(dW, 1 + leaves[3].grad, rtol=1e-4, atol=1e-4)       # accumulated
    assert torch.allclose(db, 1 + leaves[4].grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(dg, leaves[1].grad, rtol=1e-3, atol=1e-3)
    assert torch.allclose(dbeta, leaves[2].grad, rtol=1e-3, atol=1e-3)


def test_seq_mean_bwd_matches_torch():
    n, Ta, Tv, D = 3, 20, 12, 64
    dpool = torch.randn(n, D, device=DEV)
    dx = torch.full((n * (Ta + Tv), D), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.seq_mean_bwd(dpool, dx, n, Ta, D, Ta + Tv, 0)
    ops.seq_mean_bwd(2 * dpool, dx, n, Tv, D, Ta + Tv, Ta)
    ref = torch.cat([(dpool / Ta).unsqueeze(1).expand(n, Ta, D), (2 * dpool / Tv).unsqueeze(1).expand(n, Tv, D)], 1)
    assert torch.equal(dx.view(n, Ta + Tv, D), ref.to(torch.bfloat16))


@pytest.fixture(scope="module")
def ft_cases(golden_dir):
    return torch.load(os.path.join(golden_dir, "cavmaeft_base_forward.pt"), weights_only=False)


@pytest.fixture(scope="module")
def vitb_ft():
    return make_model(O.VIT_B, 527)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_vitb_ft_against_reference_golden(vitb_ft, ft_cases, idx):
    c = ft_cases[idx]
    d = O.VIT_B
    audio, video, labels = synth_ft_inputs(c["B"], c["T"], d, c["seed"])
    vitb_ft.zero_grad(set_to_none=True)
    outs = vitb_ft(audio.to(DEV), video.to(DEV), c["mode"])
    outs_t = outs if isinstance(outs, tuple) else (outs,)
    for o, ref in zip(outs_t, c["logits"]):
        assert o.shape == ref.shape and o.dtype == torch.float32
        assert float((o.cpu() - ref).abs().max()) <= 2e-2 * float(ref.abs().max()) + 2e-3
    loss = ft_loss(outs, labels_for(c["mode"], c["T"], labels).to(DEV))
    assert float(loss) == pytest.approx(c["loss"], rel=2e-2)
    loss.backward()
    named = dict(vitb_ft.named_parameters())
    got = {k for k, p in named.items() if p.grad is not None}
    want = {k for k in c["grad_norm"] if ".head." not in k}
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    for k, gref in c["grad_full"].items():
        assert cos(named[k].grad, gref) >= 0.999, (k, cos(named[k].grad, gref))
    bad = [(k, float(named[k].grad.double().norm()), c["grad_norm"][k]) for k in want
           if abs(float(named[k].grad.double().norm()) - c["grad_norm"][k]) > 0.06 * c["grad_norm"][k] + 1e-7]
    assert not bad, bad[:8]


@pytest.mark.parametrize("mode,B,T", [("mm_grad", 3, 1), ("audioonly", 2, 1), ("videoonly", 2, 3)])
def test_tiny_ft_against_oracle_all_grads(mode, B, T):
    d = O.TINY
    label_dim = 37
    model = make_model(d, label_dim)
    sd = O.init_ft_state(d, label_dim, seed=0, skip_heads=True)
    state = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    audio, video, labels = synth_ft_inputs(B, T, d, 400 + B, label_dim)
    lab = labels_for(mode, T, labels)
    ref = O.forward_ft(audio, video, state, d, mode)
    ft_loss(ref, lab).backward()
    outs = model(audio.to(DEV), video.to(DEV), mode)
    loss = ft_loss(outs, lab.to(DEV))
    assert float(loss) == pytest.approx(float(ft_loss(ref, lab)), rel=2e-2)
    loss.backward()
    named = dict(model.named_parameters())
    got = {k for k, p in named.items() if p.grad is not None}
    want = {k for k, v in state.items() if v.grad is not None}
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    low = [(k, cos(named[k].grad, state[k].grad)) for k in want if float(state[k].grad.norm()) > 1e-9]
    low = [(k, c) for k, c in low if c < 0.999]
    assert not low, low[:8]


def test_ft_eval_paths_and_fused_adam():
    d = O.TINY
    model = make_model(d, 11)
    audio, video, labels = synth_ft_inputs(2, 4, d, 77, 11)
    with torch.no_grad():
        out = model(audio.to(DEV), video.to(DEV), "mm_grad", is_eval=True)      # fuse audio with every frame (:936-980)
        assert out.shape == (2, 4, 11)
        sd = O.init_ft_state(d, 11, seed=0, skip_heads=True)
        a = O.ft_encode_audio(audio, sd, d)
        v = O.ft_encode_video(video, sd, d).reshape(2, 4, d.Tv, -1)
        for t in range(4):
            av = torch.cat((a, v[:, t]), 1)
            av = O.block(O.block(av, sd, "mm_layer_1.", d.heads, "a"), sd, "mm_layer_2.", d.heads, "a")
            ref = O.ft_head(torch.cat((av[:, :d.Ta].mean(1), av[:, d.Ta:].mean(1)), -1), sd, "mlp_head_mm")
            assert float((out[:, t].cpu() - ref).abs().max()) <= 2e-2 * float(ref.abs().max()) + 2e-3
        assert model(audio.to(DEV), None, "audioonly", is_eval=True).shape == (2, 1, 11)
        # retrieval features (:883-917): normalised audio tokens and the tokens of frame 5
        audio6, video6, _ = synth_ft_inputs(2, 6, d, 78, 11)
        fa, fv = model(audio6.to(DEV), video6.to(DEV), "retrieval")
        ra = O.ft_encode_audio(audio6, sd, d)
        rv = O.ft_encode_video(video6, sd, d).reshape(2, 6, d.Tv, -1)[:, 5]
        assert fa.shape == ra.shape and fv.shape == rv.shape
        assert float((fa.cpu() - ra).norm() / ra.norm()) < 1e-2 and float((fv.cpu() - rv).norm() / rv.norm()) < 1e-2
    # a few optimisation steps through the fused arena optimizer reduce the loss
    model.direct_grads = True
    opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=5e-7, betas=(0.95, 0.999), model=model)
    first = last = None
    for _ in range(5):
        outs = model(audio.to(DEV), video[:, :1].to(DEV), "mm_grad")
        loss = ft_loss(outs, labels.to(DEV))
        loss.backward()
        opt.step()
        first = float(loss) if first is None else first
        last = float(loss)
    assert last < first
