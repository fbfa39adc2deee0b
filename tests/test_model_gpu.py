"""Whole-path parity (B200): the drop-in CAVMAE_BASE running on libavsiam_b200.so against
  (1) the golden fixtures recorded from the UNMODIFIED reference model file (tests/golden, ViT-B/16), and
  (2) the CPU oracle executed live on identical weights, inputs and supplied mask indices.
Tolerances: masks bit-exact; losses 1e-3 relative — BASELINE.json allows 2e-2 in bf16 mode and asks 1e-3 of an fp32
mode; the bf16-operand / fp32-accumulate path measures 4e-5 ... 2.2e-4 (profiles/r01_loss_parity.log), so the fp32-mode
bound is what is asserted; per-parameter gradient cosine >= 0.999.
The benchmarked configuration itself (ViT-B/16, B = 256 per GPU, single_pass; two_pass at B = 64) is checked against
the same oracle executed in fp32 ON THE GPU (it is plain torch; blocks are recomputed in backward to bound memory)."""
import dataclasses
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import avsiam_b200  # noqa: E402
from avsiam_b200 import CAVMAE_BASE, Dims, FusedAdam  # noqa: E402
from oracle import avsiam_oracle as O  # noqa: E402
from oracle.make_golden import synth_inputs  # noqa: E402

DEV = "cuda"
LOSS_RTOL_BF16 = 1e-3      # BASELINE.json north_star: 2e-2 in bf16 mode, 1e-3 in fp32 mode — the tighter one is asserted
#                            at the reference's geometry (ViT-B/16: golden fixtures, B = 256 / 64 benchmark configs)
LOSS_RTOL_TINY = 2e-2      # the stated bf16 bound for the TINY test geometry: its pooled embeddings average only 16 + 4
#                            tokens of width 128, so bf16 rounding reaches the InfoNCE logits (x 1/0.05) un-averaged
GRAD_COS_MIN = 0.999       # "per-parameter gradients match at cosine >= 0.999"


def make_model(d: O.Dims, arrangement="two_pass", seed=0, **kw):
    model = CAVMAE_BASE(dims=Dims(**dataclasses.asdict(d)), arrangement=arrangement, **kw)
    sd = O.with_aliases(O.init_state(d, seed=seed))
    model.load_state_dict(sd, strict=True)
    return model.to(DEV), sd


def cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().double().cpu(), b.flatten().double().cpu(), dim=0))


def check_losses(out, ref, names=("loss", "loss_mae", "loss_mae_a", "loss_mae_v", "loss_c"), rtol=LOSS_RTOL_BF16):
    for i, n in enumerate(names):
        r = float(ref[i]) if not isinstance(ref, dict) else ref[n]
        assert float(out[i]) == pytest.approx(r, rel=rtol, abs=1e-4), (n, float(out[i]), r)


# ------------------------------------------------------------------------------------------------ golden (ViT-B/16)
@pytest.fixture(scope="module")
def vitb():
    model, sd = make_model(O.VIT_B)
    return model


@pytest.fixture(scope="module")
def golden_cases(golden_dir):
    return torch.load(os.path.join(golden_dir, "cavmae_base_forward.pt"), weights_only=False)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_vitb_against_reference_golden(vitb, golden_cases, idx):
    c = golden_cases[idx]
    d = O.VIT_B
    audio, imgs = synth_inputs(c["B"], d, c["seed_in"])
    vitb.mask_plan = O.make_mask_plan(c["B"], d, c["seed_mask"], two_pass=True)
    vitb.zero_grad(set_to_none=True)
    out = vitb(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=c["mae_w"], contrast_loss_weight=c["c_w"])
    check_losses(out, c)
    assert float(out[7]) == pytest.approx(c["c_acc"], abs=0.11)   # argmax hits can flip on near-ties at B<=5
    if c["mask_a"] is not None:
        assert torch.equal(out[5].cpu().to(torch.uint8), c["mask_a"])      # bit-exact masks
        assert torch.equal(out[6].cpu().to(torch.uint8), c["mask_v"])
    out[0].backward()
    named = dict(vitb.named_parameters())
    got = {k for k, p in named.items() if p.grad is not None}
    want = {k for k in c["grad_norm"] if ".head." not in k}
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    for k, gref in c["grad_full"].items():
        assert cos(named[k].grad, gref) >= GRAD_COS_MIN, (k, cos(named[k].grad, gref))
    bad = []
    for k in want:
        n = float(named[k].grad.double().norm())
        if abs(n - c["grad_norm"][k]) > 0.03 * c["grad_norm"][k] + 1e-6:
            bad.append((k, n, c["grad_norm"][k]))
    assert not bad, bad[:8]


# ------------------------------------------------------------------------------------------------ live oracle (TINY)
def run_oracle(fn, audio, imgs, sd, d, plan, **kw):
    state = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.startswith("my_blocks.")}
    out = fn(audio, imgs, state, d, plan, **kw)
    out[0].backward()
    return out, state


@pytest.mark.parametrize("arrangement,B,mae_w,c_w", [("two_pass", 4, 1.0, 0.0), ("two_pass", 7, 0.0, 1.0),
                                                      ("two_pass", 5, 1.0, 0.01), ("single_pass", 6, 1.0, 0.01),
                                                      ("single_pass", 3, 0.0, 1.0), ("single_pass", 1, 1.0, 0.0)])
def test_tiny_against_oracle_all_grads(arrangement, B, mae_w, c_w):
    d = O.TINY
    model, sd = make_model(d, arrangement)
    audio, imgs = synth_inputs(B, d, 100 + B)
    plan = O.make_mask_plan(B, d, 200 + B, two_pass=True)
    fn = O.forward if arrangement == "two_pass" else O.forward_single_pass
    ref, state = run_oracle(fn, audio, imgs, sd, d, plan, mae_loss_weight=mae_w, contrast_loss_weight=c_w)
    model.mask_plan = plan
    out = model(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=mae_w, contrast_loss_weight=c_w)
    check_losses(out, ref, rtol=LOSS_RTOL_TINY)
    if mae_w != 0:
        assert torch.equal(out[5].cpu(), ref[5]) and torch.equal(out[6].cpu(), ref[6])
    out[0].backward()
    named = dict(model.named_parameters())
    got = {k for k, p in named.items() if p.grad is not None}
    want = {k for k, v in state.items() if v.grad is not None}
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    low = [(k, cos(named[k].grad, state[k].grad)) for k in want if float(state[k].grad.norm()) > 1e-9]
    low = [(k, c) for k, c in low if c < GRAD_COS_MIN]
    assert not low, low[:8]


def test_vitb_single_pass_against_oracle(vitb):
    """The north-star arrangement at full ViT-B/16 geometry, every parameter's gradient checked."""
    d = O.VIT_B
    B = 2
    sd = O.init_state(d, seed=0, skip_heads=True)
    audio, imgs = synth_inputs(B, d, 31)
    plan = O.make_mask_plan(B, d, 32, two_pass=False)
    ref, state = run_oracle(O.forward_single_pass, audio, imgs, sd, d, plan, mae_loss_weight=1.0,
                            contrast_loss_weight=0.01)
    vitb.arrangement = "single_pass"
    try:
        vitb.mask_plan = plan
        vitb.zero_grad(set_to_none=True)
        out = vitb(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01)
        check_losses(out, ref)
        assert torch.equal(out[5].cpu(), ref[5]) and torch.equal(out[6].cpu(), ref[6])
        out[0].backward()
    finally:
        vitb.arrangement = "two_pass"
    named = dict(vitb.named_parameters())
    want = {k for k, v in state.items() if v.grad is not None}
    got = {k for k, p in named.items() if p.grad is not None}
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    low = [(k, cos(named[k].grad, state[k].grad)) for k in want]
    low = [(k, c) for k, c in low if c < GRAD_COS_MIN]
    assert not low, low[:8]


# ------------------------------------------------------------------------------------------------ API behaviour
def test_internal_rng_masks_and_eval_mode():
    d = O.TINY
    model, _ = make_model(d, "single_pass")
    audio, imgs = synth_inputs(4, d, 5)
    with torch.no_grad():
        out = model(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01)
    assert len(out) == 8 and out[0].dim() == 0 and not out[0].requires_grad
    ma, mv = out[5], out[6]
    assert ma.shape == (4, d.Ta) and mv.shape == (4, d.Tv) and ma.dtype == torch.float32
    assert torch.all(ma.sum(1) == d.Ta - O.len_keep_of(d.Ta, 0.75))
    assert torch.all(mv.sum(1) == d.Tv - O.len_keep_of(d.Tv, 0.75))
    out2 = model(audio.to(DEV), imgs.to(DEV), 0.5, 0.5, mae_loss_weight=1.0, contrast_loss_weight=0.0,
                 mask_mode="tf")   # structured audio mask (cav_mae_base.py:392-439)
    assert torch.all(out2[5].sum(1) == d.Ta - O.len_keep_of(d.Ta, 0.5))
    assert float(out2[4]) == 0.0 and float(out2[7]) == 0.0


def test_cpu_input_is_rejected():
    model, _ = make_model(O.TINY)
    a, v = synth_inputs(2, O.TINY, 1)
    with pytest.raises(RuntimeError):
        model(a, v)


def test_fused_adam_step_matches_oracle_adam():
    """Literal train-step body (traintest_cavmae_base.py:131-152 minus autocast/GradScaler) on the fast route:
    gradients stay in the arena, one fused Adam launch; compared with torch.optim.Adam on the oracle's gradients."""
    d = O.TINY
    B = 4
    model, sd = make_model(d, "single_pass")
    model.direct_grads = True
    opt = FusedAdam(model.parameters(), lr=2e-4, weight_decay=5e-7, betas=(0.95, 0.999), model=model)
    audio, imgs = synth_inputs(B, d, 9)
    plan = O.make_mask_plan(B, d, 10, two_pass=False)
    ref, state = run_oracle(O.forward_single_pass, audio, imgs, sd, d, plan, mae_loss_weight=1.0,
                            contrast_loss_weight=0.01)
    ref_params = [state[k] for k in state if state[k].grad is not None]
    ref_opt = torch.optim.Adam(ref_params, lr=2e-4, weight_decay=5e-7, betas=(0.95, 0.999))
    ref_opt.step()
    model.mask_plan = plan
    out = model(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01)
    opt.zero_grad()
    out[0].backward()
    opt.step()
    named = dict(model.named_parameters())
    sd = {k: v for k, v in sd.items()}
    # (1) exact: the fused launch == torch.optim.Adam applied to the SAME gradients (the arena's), parameter by parameter
    arena = model.arena
    used = model.last_used_names()
    for k in used:
        p0 = sd[k].clone()
        g = arena.grad(k).detach().cpu().clone()
        m, v = torch.zeros_like(p0), torch.zeros_like(p0)
        O.adam_step(p0, g, m, v, 1)
        assert torch.allclose(named[k].detach().cpu(), p0, rtol=0, atol=2e-7 + 1e-6 * float(p0.abs().max())), k
    for k, p in named.items():      # parameters without gradient are untouched (no weight decay, no moments)
        if k not in used:
            assert torch.equal(p.detach().cpu(), sd[k]), k
    # (2) against the oracle's own Adam step: the update direction on the well-conditioned entries
    worst = 1.0
    for k, v in state.items():
        if v.grad is None or float(v.grad.abs().max()) < 1e-7:
            continue
        upd = named[k].detach().cpu() - sd[k]
        ref_upd = v.detach() - sd[k]
        big = v.grad.abs() > 10 * v.grad.abs().mean() * 0.1   # ignore the sign-ambiguous near-zero gradients
        if big.sum() == 0:
            continue
        worst = min(worst, cos(upd[big], ref_upd[big]))
    assert worst > 0.98, worst
    # the bf16 shadow the next forward consumes was refreshed by the same launch
    k = "vit_base.blocks.0.attn.qkv.weight"
    assert torch.equal(model.arena.bf16(k), named[k].detach().to(torch.bfloat16))


def test_state_dict_round_trip_keeps_arena_binding():
    d = O.TINY
    model, sd = make_model(d)
    audio, imgs = synth_inputs(2, d, 3)
    model.mask_plan = O.make_mask_plan(2, d, 4)
    with torch.no_grad():
        l0 = float(model(audio.to(DEV), imgs.to(DEV), mae_loss_weight=1.0, contrast_loss_weight=0.0)[0])
        saved = {k: v.clone() for k, v in model.state_dict().items()}
        assert set(saved) == set(O.state_dict_keys(d))
        model.load_state_dict(O.with_aliases(O.init_state(d, seed=1)))
        l1 = float(model(audio.to(DEV), imgs.to(DEV), mae_loss_weight=1.0, contrast_loss_weight=0.0)[0])
        model.load_state_dict(saved)
        l2 = float(model(audio.to(DEV), imgs.to(DEV), mae_loss_weight=1.0, contrast_loss_weight=0.0)[0])
    # the loss reductions use fp32 atomics (summation order varies run to run): equal up to fp32 rounding
    assert abs(l0 - l1) > 1e-4 * abs(l0) and l2 == pytest.approx(l0, rel=1e-5)


def test_wide_geometry_against_oracle():
    """A two-block model with ViT-L's widths (embed 1024, 16 heads of 64, decoder 512 / 16 heads — BASELINE config 5
    geometry) exercises the D = 1024 LayerNorm / GEMM / attention instantiations end to end."""
    d = dataclasses.replace(O.TINY, embed_dim=1024, heads=16, dec_dim=512, dec_heads=16)
    model, sd = make_model(d, "single_pass")
    B = 2
    audio, imgs = synth_inputs(B, d, 55)
    plan = O.make_mask_plan(B, d, 56, two_pass=False)
    ref, state = run_oracle(O.forward_single_pass, audio, imgs, sd, d, plan, mae_loss_weight=1.0,
                            contrast_loss_weight=0.01)
    model.mask_plan = plan
    out = model(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01)
    check_losses(out, ref, rtol=LOSS_RTOL_TINY)
    out[0].backward()
    named = dict(model.named_parameters())
    want = {k for k, v in state.items() if v.grad is not None}
    low = [(k, cos(named[k].grad, state[k].grad)) for k in want if float(state[k].grad.norm()) > 1e-9]
    low = [(k, c) for k, c in low if c < GRAD_COS_MIN]
    assert not low, low[:8]


# ------------------------------------------------------------------------------------------------ benchmarked config
def run_oracle_gpu(fn, audio, imgs, sd, d, plan, **kw):
    """The fp32 oracle on the GPU (true fp32: TF32 off), every block recomputed in backward."""
    assert not torch.backends.cuda.matmul.allow_tf32
    state = {k: v.to(DEV).requires_grad_(True) for k, v in sd.items() if not k.startswith("my_blocks.")}
    O.CHECKPOINT_BLOCKS = True
    try:
        out = fn(audio.to(DEV), imgs.to(DEV), state, d, plan.to(DEV), **kw)
        out[0].backward()
    finally:
        O.CHECKPOINT_BLOCKS = False
    return out, state


@pytest.mark.parametrize("arrangement,B", [("single_pass", 256), ("two_pass", 64)])
def test_vitb_benchmark_config_against_fp32_oracle_on_gpu(arrangement, B):
    """bench.py's workload (BASELINE config 2: ViT-B/16, 256 samples per GPU, 75 % mask, MAE + InfoNCE) — the size at
    which every GEMM CTA walks many tiles (persistent loop, TMEM double-buffer phases, wave-aware split-K at
    K = 181 248) — against the fp32 oracle: 5 losses 1e-3, masks bit-exact, EVERY parameter's gradient cosine >= 0.999."""
    d = O.VIT_B
    torch.cuda.empty_cache()
    model, _ = make_model(d, arrangement)
    sd = O.init_state(d, seed=0, skip_heads=True)
    audio, imgs = synth_inputs(B, d, 71)
    plan = O.make_mask_plan(B, d, 72, two_pass=(arrangement == "two_pass"))
    fn = O.forward if arrangement == "two_pass" else O.forward_single_pass
    ref, state = run_oracle_gpu(fn, audio, imgs, sd, d, plan, mae_loss_weight=1.0, contrast_loss_weight=0.01)
    ref_vals = [float(r) for r in ref[:5]]
    ref_masks = (ref[5].cpu(), ref[6].cpu())
    ref_acc = float(ref[7])
    ref_grads = {k: v.grad.detach().cpu() for k, v in state.items() if v.grad is not None}
    del ref, state
    torch.cuda.empty_cache()
    model.mask_plan = plan
    out = model(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01)
    check_losses(out, ref_vals)
    assert float(out[7]) == pytest.approx(ref_acc, abs=2.0 / B)     # an argmax can flip only on a near-tie
    assert torch.equal(out[5].cpu(), ref_masks[0]) and torch.equal(out[6].cpu(), ref_masks[1])
    out[0].backward()
    named = dict(model.named_parameters())
    got = {k for k, p in named.items() if p.grad is not None}
    assert got == set(ref_grads), (sorted(got - set(ref_grads))[:5], sorted(set(ref_grads) - got)[:5])
    cosines = {k: cos(named[k].grad, g) for k, g in ref_grads.items() if float(g.norm()) > 1e-12}
    low = sorted((c, k) for k, c in cosines.items() if c < GRAD_COS_MIN)
    print(f"[parity B={B} {arrangement}] losses {[float(o) for o in out[:5]]} vs {ref_vals}; "
          f"worst gradient cosine {min(cosines.values()):.6f} over {len(cosines)} parameters")
    assert not low, low[:8]
    del model, out
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ train-step body
def _tiny_step_setup(seed=9, B=4, direct=True):
    d = O.TINY
    model, sd = make_model(d, "single_pass")
    model.direct_grads = direct
    audio, imgs = synth_inputs(B, d, seed)
    model.mask_plan = O.make_mask_plan(B, d, seed + 1, two_pass=False)
    return model, sd, audio.to(DEV), imgs.to(DEV)


def test_grad_scaler_protocol_matches_unscaled_step_and_skips_on_overflow():
    """The reference's mixed-precision protocol (traintest_cavmae_base.py:83-84,137-140): scaler.scale(loss).backward()
    -> scaler.step(optimizer) -> scaler.update(). The loss scale is a power of two, so the scaled route must land on the
    unscaled route's weights; an overflow must skip the step (weights, moments AND the bias-correction step count
    untouched) and halve the scale."""
    adam_kw = dict(lr=2e-4, weight_decay=5e-7, betas=(0.95, 0.999))
    ref_model, _, audio, imgs = _tiny_step_setup()
    ref_opt = FusedAdam(ref_model.parameters(), **adam_kw)
    for _ in range(2):
        out = ref_model(audio, imgs, 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01)
        ref_opt.zero_grad()
        out[0].backward()
        ref_opt.step()
    model, sd, audio, imgs = _tiny_step_setup()
    opt = FusedAdam(model.parameters(), **adam_kw)
    scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 12)
    losses = []
    for it in range(3):
        with torch.autocast("cuda", dtype=torch.float16):       # `with autocast():` at :131 — a no-op for this module
            out = model(audio, imgs, 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01)
        opt.zero_grad()
        scaler.scale(out[0]).backward()
        if it == 1:                                             # overflow in the second iteration
            before = model.arena.flat.clone()
            m_before = opt._m.clone()
            model.arena.grads[5] = float("inf")
        scaler.step(opt)
        scale_before = scaler.get_scale()
        scaler.update()
        if it == 1:
            assert torch.equal(model.arena.flat, before) and torch.equal(opt._m, m_before)
            assert scaler.get_scale() == scale_before * 0.5
            assert opt.steps_taken() == 1
        losses.append(float(out[0]))
    assert opt.steps_taken() == 2
    a, b = model.arena.flat[:model.arena.n_hot], ref_model.arena.flat[:ref_model.arena.n_hot]
    # identical up to the rounding of bf16 activations' gradients under a power-of-two scale (exact unless a value
    # crosses the bf16 subnormal range) and the fp32 1/scale multiply
    assert float((a - b).abs().max()) <= 2e-6 + 1e-5 * float(b.abs().max()), float((a - b).abs().max())


def test_two_optimizer_two_pass_body():
    """traintest_cavmae_base.py:131-152 literally: contrastive pass -> optimizer, MAE pass -> optimizer2, each Adam with its
    own moments over the same parameter list; compared with the oracle + two torch.optim.Adam on the same draws."""
    d = O.TINY
    B = 5
    model, sd = make_model(d, "two_pass")
    model.direct_grads = True
    adam_kw = dict(lr=2e-4, weight_decay=5e-7, betas=(0.95, 0.999))
    trainables = [p for p in model.parameters() if p.requires_grad]
    opt1, opt2 = FusedAdam(trainables, **adam_kw), FusedAdam(trainables, **adam_kw)
    state = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.startswith("my_blocks.")}
    ref1 = torch.optim.Adam(list(state.values()), **adam_kw)
    ref2 = torch.optim.Adam(list(state.values()), **adam_kw)
    audio, imgs = synth_inputs(B, d, 77)
    for it in range(2):
        plan = O.make_mask_plan(B, d, 400 + it, two_pass=True)
        model.mask_plan = plan
        for (mw, cw, opt, ropt) in ((0, 1, opt1, ref1), (1, 0, opt2, ref2)):
            out = model(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=mw, contrast_loss_weight=cw)
            opt.zero_grad()
            out[0].backward()
            opt.step()
            ref = O.forward(audio, imgs, state, d, plan, mae_loss_weight=mw, contrast_loss_weight=cw)
            ropt.zero_grad(set_to_none=True)
            ref[0].backward()
            ropt.step()
            # the losses of the second iteration are taken on weights both optimizers have already moved
            assert float(out[0]) == pytest.approx(float(ref[0]), rel=LOSS_RTOL_TINY, abs=1e-4), (it, mw, cw)
    named = dict(model.named_parameters())
    moved = 0
    for k, v in state.items():
        delta_ref = v.detach() - sd[k]
        delta = named[k].detach().cpu() - sd[k]
        if float(delta_ref.abs().max()) == 0:
            assert float(delta.abs().max()) == 0, k
            continue
        moved += 1
        if k.endswith("attn.qkv.bias"):     # the key-bias gradient is identically zero in exact arithmetic (softmax is
            n3 = delta.numel() // 3        # shift-invariant): Adam turns its rounding noise into full-size steps
            delta, delta_ref = torch.cat([delta[:n3], delta[2 * n3:]]), torch.cat([delta_ref[:n3], delta_ref[2 * n3:]])
        # Adam steps are ~lr * sign-like: after 2 x 2 steps compare where the reference moved decisively
        big = delta_ref.abs() > 0.5 * delta_ref.abs().max()
        assert cos(delta[big], delta_ref[big]) > 0.95, (k, cos(delta[big], delta_ref[big]))
    assert moved > 50


def test_fused_adam_param_groups_and_excluded_parameters():
    """torch semantics of param_groups (traintest_ft_base.py:78-83: three groups with their own lr) and of parameters
    left out of the optimizer (freeze_base): per-group lr / weight_decay, no update and no decay outside the groups."""
    model, sd, audio, imgs = _tiny_step_setup(direct=False)
    named = dict(model.named_parameters())
    heads = [p for k, p in named.items() if k.startswith("decoder_")]
    fusion = [p for k, p in named.items() if k.startswith("mm_layer_")]
    base = [p for k, p in named.items() if k.startswith("vit_base.blocks.")]
    groups = [{"params": base, "lr": 1e-4}, {"params": heads, "lr": 1e-2, "weight_decay": 1e-3},
              {"params": fusion, "lr": 3e-3}]
    opt = FusedAdam(groups, lr=5e-5, weight_decay=5e-7, betas=(0.95, 0.999), model=model)
    out = model(audio, imgs, 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01)
    opt.zero_grad()
    out[0].backward()
    grads = {k: p.grad.detach().clone() for k, p in named.items() if p.grad is not None}
    before = {k: p.detach().clone() for k, p in named.items()}
    ref_params = {k: before[k].clone().requires_grad_(True) for k in named}
    for k, g in grads.items():
        ref_params[k].grad = g.clone()
    pick = lambda ps: [ref_params[k] for k, p in named.items() if any(p is q for q in ps)]
    ref = torch.optim.Adam([{"params": pick(base), "lr": 1e-4}, {"params": pick(heads), "lr": 1e-2, "weight_decay": 1e-3},
                            {"params": pick(fusion), "lr": 3e-3}], lr=5e-5, weight_decay=5e-7, betas=(0.95, 0.999))
    ref.step()
    opt.step()
    for k, p in named.items():
        tol = 2e-7 + 2e-6 * float(before[k].abs().max())
        assert torch.allclose(p.detach(), ref_params[k].detach(), rtol=0, atol=tol), k
    untouched = [k for k in named if k.startswith("vit_base.patch_embed") or k.startswith("vit_base.norm")]
    assert untouched and all(torch.equal(named[k].detach(), before[k]) for k in untouched)   # had grads, not in a group
    # torch.optim.Adam loads our state_dict and vice versa (best_optim_state.pth interchange)
    sd_f = opt.state_dict()
    ref.load_state_dict(sd_f)
    sd_t = ref.state_dict()
    opt2 = FusedAdam(groups, lr=5e-5, weight_decay=5e-7, betas=(0.95, 0.999), model=model)
    opt2.load_state_dict(torch.load(_roundtrip(sd_t), map_location="cpu"))
    assert torch.equal(opt2._m, opt._m) and torch.equal(opt2._v, opt._v) and opt2.steps_taken() == 1
    assert opt2._m.is_cuda


def _roundtrip(obj):
    import io
    buf = io.BytesIO()
    torch.save(obj, buf)
    buf.seek(0)
    return buf


def test_vit_large_single_pass_against_fp32_oracle_on_gpu():
    """BASELINE config 5, ViT-L/16 geometry (D 1024, depth 24, 16 heads of 64, decoder 512 x 6 x 16 heads,
    512 / 196 tokens): the full-depth model in the single-pass arrangement against the fp32 oracle on the GPU — losses
    1e-3, masks bit-exact, every parameter's gradient cosine >= 0.999."""
    d = dataclasses.replace(O.VIT_B, embed_dim=1024, depth=24, heads=16, dec_depth=6, head_classes=64)
    B = 8
    torch.cuda.empty_cache()
    model, _ = make_model(d, "single_pass")
    sd = O.init_state(d, seed=0, skip_heads=True)
    audio, imgs = synth_inputs(B, d, 91)
    plan = O.make_mask_plan(B, d, 92, two_pass=False)
    ref, state = run_oracle_gpu(O.forward_single_pass, audio, imgs, sd, d, plan, mae_loss_weight=1.0,
                                contrast_loss_weight=0.01)
    ref_vals = [float(r) for r in ref[:5]]
    ref_masks = (ref[5].cpu(), ref[6].cpu())
    ref_grads = {k: v.grad.detach().cpu() for k, v in state.items() if v.grad is not None}
    del ref, state
    torch.cuda.empty_cache()
    model.mask_plan = plan
    out = model(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01)
    check_losses(out, ref_vals)
    assert torch.equal(out[5].cpu(), ref_masks[0]) and torch.equal(out[6].cpu(), ref_masks[1])
    out[0].backward()
    named = dict(model.named_parameters())
    got = {k for k, p in named.items() if p.grad is not None}
    assert got == set(ref_grads), (sorted(got - set(ref_grads))[:5], sorted(set(ref_grads) - got)[:5])
    cosines = {k: cos(named[k].grad, g) for k, g in ref_grads.items() if float(g.norm()) > 1e-12}
    low = sorted((c, k) for k, c in cosines.items() if c < GRAD_COS_MIN)
    print(f"[parity ViT-L B={B}] losses {[float(o) for o in out[:5]]} vs {ref_vals}; worst gradient cosine "
          f"{min(cosines.values()):.6f} over {len(cosines)} parameters")
    assert not low, low[:8]
    del model, out
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ ViT-H/14 (config 5)
def _huge_case(d, B, seed, ratio=0.75):
    """CAVMAE_HUGE as shipped (cav_mae_huge.cpython-39.pyc): single pass over the shared blocks, no
    decoder, single-direction InfoNCE."""
    model, sd = make_model(d, "single_pass", bidirect_contrast=False)
    audio, imgs = synth_inputs(B, d, seed)
    plan = O.make_mask_plan(B, d, seed + 1, two_pass=False)
    kw = dict(mae_loss_weight=0.0, contrast_loss_weight=1.0, ratio_a=ratio, ratio_v=ratio, bidirect=False)
    return model, sd, audio, imgs, plan, kw


def _check_contrastive_only(model, out, ref, ref_grads, rtol, tag):
    for i in (0, 4):        # loss, loss_c (the MAE losses are zero on both sides)
        assert float(out[i]) == pytest.approx(float(ref[i]), rel=rtol, abs=1e-4), (i, float(out[i]), float(ref[i]))
    assert float(out[1]) == 0.0 and float(ref[1]) == 0.0
    assert float(out[7]) == pytest.approx(float(ref[7]), abs=1e-6)      # c_acc
    out[0].backward()
    named = dict(model.named_parameters())
    got = {k for k, p in named.items() if p.grad is not None}
    assert got == set(ref_grads), (sorted(got - set(ref_grads))[:5], sorted(set(ref_grads) - got)[:5])
    cosines = {k: cos(named[k].grad, g) for k, g in ref_grads.items() if float(g.norm()) > 1e-12}
    low = sorted((c, k) for k, c in cosines.items() if c < GRAD_COS_MIN)
    print(f"[parity {tag}] loss {float(out[0]):.6f} vs {float(ref[0]):.6f}; worst gradient cosine "
          f"{min(cosines.values()):.6f} over {len(cosines)} parameters")
    assert not low, low[:8]


@pytest.mark.parametrize("B,ratio", [(5, 0.5), (3, 0.0)])
def test_huge_geometry_tiny_against_oracle(B, ratio):
    """ViT-H's width (16 heads of 80 — the only LayerNorm width of this library that is a multiple of 80) at depth 2,
    patch 14 on inputs it does not divide (100 x 30 fbank -> 7 x 2 tokens, 60 x 60 frame -> 4 x 4), K-padded patch
    embedding (196 -> 200, 588 -> 592), every parameter's gradient against the live oracle."""
    d = dataclasses.replace(O.TINY, embed_dim=1280, heads=16, patch=14, audio_len=100, mel=30, img=60)
    assert (d.Ta, d.Tv) == (14, 16)
    model, sd, audio, imgs, plan, kw = _huge_case(d, B, 300 + B, ratio)
    ref, state = run_oracle(O.forward_single_pass, audio, imgs, sd, d, plan, **kw)
    model.mask_plan = plan
    out = model(audio.to(DEV), imgs.to(DEV), ratio, ratio, mae_loss_weight=0.0, contrast_loss_weight=1.0)
    ref_grads = {k: v.grad for k, v in state.items() if v.grad is not None}
    _check_contrastive_only(model, out, ref, ref_grads, LOSS_RTOL_TINY, f"huge-geometry tiny B={B} ratio={ratio}")
    with pytest.raises(RuntimeError, match="MAE branch needs a patch size dividing"):
        model(audio.to(DEV), imgs.to(DEV), ratio, ratio, mae_loss_weight=1.0, contrast_loss_weight=1.0)


def test_vit_huge_contrastive_against_fp32_oracle_on_gpu():
    """BASELINE config 5, ViT-H/14 at FULL width and depth (D 1280, 32 blocks, 16 heads of 80, patch 14, 657 / 256 tokens,
    164 / 64 kept at 75 %) against the fp32 oracle on the GPU: every parameter's gradient cosine >= 0.999; the loss here is
    the bare InfoNCE over 8 samples at temperature 0.05 (no MAE term beside it, 32 bf16 blocks in front of the x20 logits):
    5e-3 asserted (measured 1.2e-3 at B = 4; the stated bf16 bound is 2e-2, bench.py's step-0 gate at B = 16 measures 6e-6)."""
    d = dataclasses.replace(O.VIT_H, head_classes=64, dec_depth=1)     # the (unused) decoder kept minimal
    assert (d.Ta, d.Tv, d.embed_dim // d.heads) == (657, 256, 80)
    B = 8
    torch.cuda.empty_cache()
    model, sd, audio, imgs, plan, kw = _huge_case(d, B, 95)
    sd = {k: v for k, v in sd.items() if ".head." not in k}
    ref, state = run_oracle_gpu(O.forward_single_pass, audio, imgs, sd, d, plan, **kw)
    ref_vals = [float(r) for r in ref[:5]] + [None, None, float(ref[7])]
    ref_grads = {k: v.grad.detach().cpu() for k, v in state.items() if v.grad is not None}
    del ref, state
    torch.cuda.empty_cache()
    model.mask_plan = plan
    out = model(audio.to(DEV), imgs.to(DEV), 0.75, 0.75, mae_loss_weight=0.0, contrast_loss_weight=1.0)
    _check_contrastive_only(model, out, ref_vals, ref_grads, 5e-3, f"ViT-H/14 B={B}")
    del model, out
    torch.cuda.empty_cache()


def test_graphed_train_step_matches_eager():
    """GraphedTrainStep (forward + reverse pass + fused Adam in one CUDA graph, device-side step counter) against the
    same steps issued eagerly. Mask ratio 0 with the contrastive loss only makes the step independent of the mask draws
    (every token is kept; pooling and attention are permutation-invariant), so both runs follow the same trajectory."""
    from avsiam_b200 import GraphedTrainStep
    d = O.TINY
    B, K, W = 4, 4, 2
    audio, imgs = synth_inputs(B, d, 61)
    audio, imgs = audio.to(DEV), imgs.to(DEV)
    adam_kw = dict(lr=2e-4, weight_decay=5e-7, betas=(0.95, 0.999))
    runs = []
    for graphed in (False, True):
        model, sd = make_model(d, "single_pass")
        model.direct_grads = True
        opt = FusedAdam(model.parameters(), **adam_kw)
        losses = []
        if graphed:
            step = GraphedTrainStep(model, opt, 0.0, 0.0, mae_loss_weight=0.0, contrast_loss_weight=1.0, warmup=W)
            for _ in range(K):
                losses.append(float(step(audio, imgs)))
            assert step.launches_per_step > 20 and opt.steps_taken() == W + K
        else:
            for _ in range(W + K):
                out = model(audio, imgs, 0.0, 0.0, mae_loss_weight=0.0, contrast_loss_weight=1.0)
                opt.zero_grad()
                out[0].backward()
                opt.step()
                losses.append(float(out[0]))
            losses = losses[W:]
        runs.append((losses, model.arena.flat[:model.arena.n_hot].detach().cpu().clone(), sd))
    (l0, w0, sd), (l1, w1, _) = runs
    for a, b in zip(l0, l1):
        assert b == pytest.approx(a, rel=2e-3, abs=1e-4), (l0, l1)
    assert l0[-1] < l0[0]                                   # it trains
    model, _ = make_model(d, "single_pass")
    init = model.arena.flat[:model.arena.n_hot].detach().cpu()
    moved = (w0 - init).abs() > 1e-6
    assert cos((w1 - init)[moved], (w0 - init)[moved]) > 0.97
