"""Per-kernel parity tests (B200): every C-ABI entry against the oracle / a plain fp32 torch statement of the op.
Index work is bit-exact; floating point within the tolerance written beside each assert (bf16 operands)."""
import dataclasses
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from avsiam_b200 import ops  # noqa: E402
from oracle import avsiam_oracle as O  # noqa: E402

DEV = "cuda"


def rnd(*shape, scale=1.0, dtype=torch.float32, seed=None):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed if seed is not None else (hash(shape) % 1000) + 1)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def rel_err(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(354, 768, 768), (1000, 2304, 768), (256, 128, 200), (45, 512, 3072),
                                   # two-CTA kernel edges (N > 128, M > 128): a pair whose second CTA has no valid row
                                   # (M = 129, 300), ragged N / K tails, one more tile than clusters (74 on a B200),
                                   # many tiles per cluster with both TMEM buffers and every ring phase in play
                                   (129, 136, 72), (257, 264, 64), (300, 520, 200), (256, 256, 64), (19200, 256, 512),
                                   (4000, 3072, 768)])
def test_gemm_fwd_bias_gelu_resid(M, N, K):
    a, w = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=K ** -0.5, dtype=torch.bfloat16)
    bias, res = rnd(N), rnd(M, N, dtype=torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    pre = torch.empty_like(out)
    ref_pre = a.float() @ w.float().t() + bias
    ops.gemm(a, w, out, M, N, K, bias=bias, gelu=True, aux_out=pre)          # fc1: GELU class, second output = pre
    assert rel_err(pre, ref_pre) < 6e-3      # bf16 output rounding (2^-9) dominates
    assert rel_err(out, F.gelu(ref_pre)) < 6e-3
    ops.gemm(a, w, out, M, N, K, bias=bias, gelu=True)                       # no second output
    assert rel_err(out, F.gelu(ref_pre)) < 6e-3
    ops.gemm(a, w, out, M, N, K, bias=bias, resid=res)                       # proj / fc2: residual class
    assert rel_err(out, ref_pre + res.float()) < 6e-3
    ops.gemm(a, w, out, M, N, K, bias=bias, alpha=0.5)                       # plain class
    assert rel_err(out, 0.5 * ref_pre) < 6e-3
    with pytest.raises(RuntimeError):                                        # one epilogue class per launch
        ops.gemm(a, w, out, M, N, K, bias=bias, gelu=True, resid=res)


@pytest.mark.parametrize("env", [{"AVS_GEMM_2CTA": "0"}, {"AVS_GEMM_EW16": "0", "AVS_GEMM_MUL16": "0"}])
def test_gemm_fallback_kernels_in_a_subprocess(env):
    """The kernel choice is read from the environment once per process: the one-CTA kernel for every product
    (AVS_GEMM_2CTA=0) and the 8-warp GELU / product epilogues of the two-CTA kernel (AVS_GEMM_EW16=0, AVS_GEMM_MUL16=0) stay covered by running this
    file's GEMM tests in a child process with the switch set."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                        "gemm and not subprocess"], cwd=root, env={**os.environ, **env}, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_gemm_rowadd_alpha_patch_embed_epilogue():
    M, N, K, T = 300, 768, 256, 64
    a, w, bias, pos = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=K ** -0.5, dtype=torch.bfloat16), rnd(N), rnd(T, N)
    idx = torch.randint(0, T, (M,), device=DEV, dtype=torch.int32)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a, w, out, M, N, K, bias=bias, rowadd=pos, rowidx=idx, alpha=2.0)
    ref = 2.0 * (a.float() @ w.float().t() + bias + pos[idx.long()])
    assert rel_err(out, ref) < 6e-3
    ops.gemm(a, w, out, M, N, K, bias=bias, rowadd=pos, alpha=2.0)  # implicit m % T
    ref = 2.0 * (a.float() @ w.float().t() + bias + pos[torch.arange(M, device=DEV) % T])
    assert rel_err(out, ref) < 6e-3


def test_gemm_dgrad_dgelu_and_wgrad_accumulate():
    M, N, K = 708, 3072, 768  # y = x W^T, W [N,K]
    x, w, dy = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=K ** -0.5, dtype=torch.bfloat16), rnd(M, N, dtype=torch.bfloat16)
    # dgrad: dx[M,K] = dy[M,N] @ W[N,K]  -> reduction N; B operand = W read MN-major
    dx = torch.empty(M, K, dtype=torch.bfloat16, device=DEV)
    pre = rnd(M, K, dtype=torch.bfloat16)
    ops.gemm(dy, w, dx, M, K, N, b_major=ops.MAJOR_MN, dgelu_aux=pre)
    xr = pre.float().requires_grad_(True)
    F.gelu(xr).backward(dy.float() @ w.float())
    assert rel_err(dx, xr.grad) < 6e-3
    # wgrad: dW[N,K] += dy^T @ x  -> reduction over M; both operands MN-major; fp32 accumulate with split-K
    dw = torch.ones(N, K, dtype=torch.float32, device=DEV)
    ops.gemm(dy, x, dw, N, K, M, a_major=ops.MAJOR_MN, b_major=ops.MAJOR_MN, accumulate=True, split_k=0)
    ref = 1.0 + dy.float().t() @ x.float()
    assert rel_err(dw, ref) < 1e-4           # fp32 accumulate of exact bf16 products


def test_gemm_stored_gelu_derivative_pair():
    """fc1's epilogue stores gelu'(pre) instead of pre (aux_grad), fc2's dgrad epilogue multiplies by it (mul_aux): the
    pair must reproduce gelu / dgelu of the exact-erf GELU (timm Mlp at cav_mae_base.py:138-143)."""
    M, N, K = 708, 2048, 512
    x, w = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=K ** -0.5, dtype=torch.bfloat16)
    bias = rnd(N)
    act = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    dact = torch.empty_like(act)
    ops.gemm(x, w, act, M, N, K, bias=bias, gelu=True, aux_out=dact, aux_grad=True)
    pre = (x.float() @ w.float().t() + bias).requires_grad_(True)
    ref = F.gelu(pre)
    ref.sum().backward()
    assert rel_err(act, ref) < 6e-3
    assert rel_err(dact, pre.grad) < 6e-3                       # gelu'(pre), bf16 rounding dominates
    dy, w2 = rnd(M, K, dtype=torch.bfloat16), rnd(K, N, scale=K ** -0.5, dtype=torch.bfloat16)
    dh = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    cs = torch.zeros(N, device=DEV)
    ops.gemm(dy, w2, dh, M, N, K, b_major=ops.MAJOR_MN, mul_aux=dact, colsum=cs)
    want = (dy.float() @ w2.float()) * dact.float()
    assert rel_err(dh, want) < 6e-3
    assert torch.allclose(cs, want.sum(0), rtol=2e-3, atol=2e-2 * math.sqrt(M))
    with pytest.raises(RuntimeError):
        ops.gemm(dy, w2, dh, M, N, K, b_major=ops.MAJOR_MN, mul_aux=dact, dgelu_aux=dact)


def test_gemm_rejects_bad_input():
    a = rnd(8, 60, dtype=torch.bfloat16)  # pitch 60 not a multiple of 8
    w = rnd(16, 60, dtype=torch.bfloat16)
    out = torch.empty(8, 16, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError):
        ops.gemm(a, w, out, 8, 16, 60)


# ------------------------------------------------------------------------------------------------ masking
@pytest.mark.parametrize("N,L,ratio", [(7, 512, 0.75), (5, 196, 0.75), (3, 657, 0.4), (2, 40, 0.0), (4, 1, 0.0)])
def test_mask_argsort_bit_exact(N, L, ratio):
    noise = torch.rand(N, L, device=DEV)
    noise[0, : L // 2] = 1.1  # ties, as produced by structured masking (cav_mae_base.py:408,413)
    keep = O.len_keep_of(L, ratio)
    ids, ids_restore, mask = ops.mask_argsort(noise, keep)
    ref_ids = torch.argsort(noise.cpu(), dim=1, stable=True)
    assert torch.equal(ids.cpu().long(), ref_ids)
    x = torch.randn(N, L, 8)
    _, ref_mask, ref_restore = O.apply_masking(x, ref_ids, keep)
    assert torch.equal(ids_restore.cpu().long(), ref_restore)
    assert torch.equal(mask.cpu(), ref_mask)


def test_gather_rows_bit_exact_and_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "masking_unstructured.pt"), weights_only=False)
    x, ids = g["x"].to(DEV), g["ids_shuffle"].to(DEV).int()
    keep = O.len_keep_of(x.shape[1], g["mask_ratio"])
    out = ops.gather_rows(x, ids, keep)
    assert torch.equal(out.cpu(), g["x_masked"])          # reference's own torch.gather output
    ids2, ids_restore, mask = ops.mask_from_ids(ids, keep)
    assert torch.equal(ids2, ids)
    assert torch.equal(ids_restore.cpu().long(), g["ids_restore"])
    assert torch.equal(mask.cpu(), g["mask"])
    xb = torch.randn(9, 512, 768, device=DEV).to(torch.bfloat16)
    idb = torch.argsort(torch.rand(9, 512, device=DEV), dim=1).int()
    ob = ops.gather_rows(xb, idb, 128)
    ref = torch.gather(xb, 1, idb[:, :128].long().unsqueeze(-1).expand(-1, -1, 768))
    assert torch.equal(ob, ref)
    assert ops.gather_rows(xb, idb, 0).shape == (9, 0, 768)  # empty keep


def test_structured_masking_against_reference_golden(golden_dir):
    """random_masking_structured, modes 'time' / 'freq' / 'tf' (cav_mae_base.py:392-439): same random.seed, same noise
    as the UNMODIFIED reference method (tests/golden/masking_structured.pt, oracle/make_golden_masking.py).
    Bit-exact: everything the sort determines. When more tokens are forced to 1.1 than are removed ('order_free'), the
    reference keeps a few forced tokens picked by torch.argsort's unspecified tie order; there the un-forced part of
    the result (kept rows, their order, mask, ranks) is still bit-exact and the number of forced survivors equal."""
    import dataclasses
    import random
    from avsiam_b200 import CAVMAE_BASE, Dims
    cases = torch.load(os.path.join(golden_dir, "masking_structured.pt"), weights_only=False)
    models = {}
    for c in cases:
        t, f, N = c["t"], c["f"], c["N"]
        L = t * f
        if (t, f) not in models:
            d = dataclasses.replace(O.TINY, audio_len=16 * t, mel=16 * f)
            models[(t, f)] = CAVMAE_BASE(dims=Dims(**dataclasses.asdict(d)))
        model = models[(t, f)]
        random.seed(c["seed"])
        ids, ids_restore, mask, keep = model._structured_ids(N, L, c["ratio"], torch.device(DEV), mode=c["mode"],
                                                             noise=c["noise"])
        assert keep == c["x_masked"].shape[1]
        xm = ops.gather_rows(c["x"].to(DEV), ids, keep).cpu()
        mask, ids_restore = mask.cpu(), ids_restore.cpu().long()
        n_forced = c["n_forced"]
        if not c["order_free"]:
            assert torch.equal(xm, c["x_masked"]) and torch.equal(mask, c["mask"]), (c["mode"], c["ratio"])
        for i in range(N):
            free = L - int(n_forced[i])                      # tokens with their own random noise: ranks 0 .. free-1
            k_free = min(keep, free)
            assert torch.equal(xm[i, :k_free], c["x_masked"][i, :k_free])
            unforced = c["ids_restore"][i] < free
            assert torch.equal(ids_restore[i] < free, unforced)   # the same columns / rows were forced
            assert torch.equal(ids_restore[i][unforced], c["ids_restore"][i][unforced])
            assert torch.equal(mask[i][unforced], c["mask"][i][unforced])
            assert float(mask[i].sum()) == float(c["mask"][i].sum()) == L - keep


def test_structured_masking_rejects_unknown_mode():
    import dataclasses
    from avsiam_b200 import CAVMAE_BASE, Dims
    model = CAVMAE_BASE(dims=Dims(**dataclasses.asdict(O.TINY)))
    with pytest.raises(ValueError):
        model._structured_ids(2, O.TINY.Ta, 0.5, torch.device(DEV), mode="diagonal")


def test_patchify_matches_patch_embed_order():
    d = O.VIT_B
    B = 3
    audio, img = rnd(B, d.audio_len, d.mel), rnd(B, 3, d.img, d.img)
    ids_a = torch.argsort(torch.rand(B, d.Ta, device=DEV), dim=1).int()
    ids_v = torch.argsort(torch.rand(B, d.Tv, device=DEV), dim=1).int()
    ka, kv = 128, 49
    pa = torch.empty(B * ka, 256, dtype=torch.bfloat16, device=DEV)
    pv = torch.empty(B * kv, 768, dtype=torch.bfloat16, device=DEV)
    ops.patchify_audio(audio, ids_a, ka, 16, pa)
    ops.patchify_video(img, ids_v, kv, 16, pv)
    eye_a = torch.eye(256).reshape(256, 1, 16, 16)
    eye_v = torch.eye(768).reshape(768, 3, 16, 16)
    ref_a = O.patch_embed_audio(audio.cpu(), eye_a, torch.zeros(256), d)   # identity weights expose the vector order
    ref_v = O.patch_embed_video(img.cpu(), eye_v, torch.zeros(768), d)
    ref_a = torch.gather(ref_a, 1, ids_a[:, :ka].cpu().long().unsqueeze(-1).expand(-1, -1, 256)).reshape(B * ka, 256)
    ref_v = torch.gather(ref_v, 1, ids_v[:, :kv].cpu().long().unsqueeze(-1).expand(-1, -1, 768)).reshape(B * kv, 768)
    assert torch.equal(pa.cpu(), ref_a.to(torch.bfloat16))
    assert torch.equal(pv.cpu(), ref_v.to(torch.bfloat16))
    full = torch.empty(B * d.Ta, 256, dtype=torch.bfloat16, device=DEV)
    ops.patchify_audio(audio, None, d.Ta, 16, full)
    assert torch.equal(full.cpu(), O.patch_embed_audio(audio.cpu(), eye_a, torch.zeros(256), d).reshape(-1, 256).to(torch.bfloat16))


def test_patchify_patch14_matches_strided_conv_order():
    """ViT-H/14 geometry: patch 14 neither divides 1024 x 128 nor is a multiple of 8; the kernel must
    read exactly what Conv2d(kernel = stride = 14) reads (73 x 9 = 657 audio tokens from the first 1022 x 126 samples,
    16 x 16 video tokens) and zero the K padding (196 -> 200, 588 -> 592 columns). Checker: torch's own conv2d with
    identity weights (the reference's PatchEmbed is that Conv2d, cav_mae_base.py:96-100), and the oracle's restatement."""
    d = dataclasses.replace(O.VIT_B, patch=14)
    assert (d.ta, d.fa, d.Ta, d.Tv) == (73, 9, 657, 256)
    B = 2
    audio, img = rnd(B, d.audio_len, d.mel), rnd(B, 3, d.img, d.img)
    ids_a = torch.argsort(torch.rand(B, d.Ta, device=DEV), dim=1).int()
    ids_v = torch.argsort(torch.rand(B, d.Tv, device=DEV), dim=1).int()
    ka, kv = 164, 64
    pa = torch.full((B * ka, 200), float("nan"), dtype=torch.bfloat16, device=DEV)
    pv = torch.full((B * kv, 592), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.patchify_audio(audio, ids_a, ka, 14, pa)
    ops.patchify_video(img, ids_v, kv, 14, pv)
    eye_a = torch.eye(196).reshape(196, 1, 14, 14)
    eye_v = torch.eye(588).reshape(588, 3, 14, 14)
    conv_a = torch.nn.functional.conv2d(audio.cpu().unsqueeze(1).transpose(2, 3), eye_a, stride=14).flatten(2).transpose(1, 2)
    conv_v = torch.nn.functional.conv2d(img.cpu(), eye_v, stride=14).flatten(2).transpose(1, 2)
    assert torch.equal(conv_a, O.patch_embed_audio(audio.cpu(), eye_a, torch.zeros(196), d))
    assert torch.equal(conv_v, O.patch_embed_video(img.cpu(), eye_v, torch.zeros(588), d))
    ref_a = torch.gather(conv_a, 1, ids_a[:, :ka].cpu().long().unsqueeze(-1).expand(-1, -1, 196)).reshape(B * ka, 196)
    ref_v = torch.gather(conv_v, 1, ids_v[:, :kv].cpu().long().unsqueeze(-1).expand(-1, -1, 588)).reshape(B * kv, 588)
    assert torch.equal(pa[:, :196].cpu(), ref_a.to(torch.bfloat16))
    assert torch.equal(pv[:, :588].cpu(), ref_v.to(torch.bfloat16))
    assert float(pa[:, 196:].abs().max()) == 0.0 and float(pv[:, 588:].abs().max()) == 0.0


def test_decoder_restore_fwd_bwd():
    B, Ta, Tv, ka, kv, D = 3, 64, 36, 16, 9, 64
    torch.manual_seed(11)
    x = rnd(B, ka + kv, D, dtype=torch.bfloat16)
    ira = torch.argsort(torch.argsort(torch.rand(B, Ta, device=DEV), dim=1), dim=1).int()
    irv = torch.argsort(torch.argsort(torch.rand(B, Tv, device=DEV), dim=1), dim=1).int()
    mt, pos_a, pos_v, mod_a, mod_v = rnd(D), rnd(Ta, D), rnd(Tv, D), rnd(D), rnd(D)
    out = torch.empty(B, Ta + Tv, D, dtype=torch.bfloat16, device=DEV)
    ops.decoder_restore_fwd(x, ira, irv, mt, pos_a, pos_v, mod_a, mod_v, out, B, Ta, Tv, ka, kv, D)
    leaves = [t.clone().float().requires_grad_(True) for t in (x, mt, pos_a, pos_v, mod_a, mod_v)]
    xr, mtr, par, pvr, mar, mvr = leaves
    a_ = torch.cat([xr[:, :ka], mtr.expand(B, Ta - ka, D)], 1)
    a_ = torch.gather(a_, 1, ira.long().unsqueeze(-1).expand(-1, -1, D)) + par + mar
    v_ = torch.cat([xr[:, ka:], mtr.expand(B, Tv - kv, D)], 1)
    v_ = torch.gather(v_, 1, irv.long().unsqueeze(-1).expand(-1, -1, D)) + pvr + mvr
    ref = torch.cat([a_, v_], 1)
    assert torch.equal(out, ref.to(torch.bfloat16))       # pure index work + one rounding
    dout = rnd(B, Ta + Tv, D, dtype=torch.bfloat16)
    ref.backward(dout.float())
    dx = torch.full_like(x, float("nan"))
    grads = [torch.zeros_like(t) for t in (mt, pos_a, pos_v, mod_a, mod_v)]
    ops.decoder_restore_bwd(dout, ira, irv, dx, *grads, B, Ta, Tv, ka, kv, D)
    assert torch.equal(dx.float(), xr.grad)
    for g, l in zip(grads, leaves[1:]):   # fp32 atomics: the summation order differs from torch's
        assert torch.allclose(g, l.grad, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("M,D,eps", [(177 * 3, 768, 1e-5), (708 * 2, 512, 1e-5), (98, 1280, 1e-6), (40, 64, 1e-5),
                                     (1001, 128, 1e-5), (5, 1024, 1e-6), (77, 256, 1e-5)])
def test_layernorm_fwd_bwd(M, D, eps):
    x = rnd(M, D, scale=2.0, dtype=torch.bfloat16)
    gamma, beta = 1 + 0.1 * rnd(D), 0.1 * rnd(D)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, eps, y, mean, rstd, M, D)
    xr, gr, br = x.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), gr, br, eps)
    assert rel_err(y, ref) < 4e-3
    dy, res = rnd(M, D, dtype=torch.bfloat16), rnd(M, D, dtype=torch.bfloat16)
    ref.backward(dy.float())
    dx = torch.empty_like(x)
    dg, db, dbias = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.ones(D, device=DEV)
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, M, D, resid=res, dbias=dbias)
    assert rel_err(dx, xr.grad + res.float()) < 5e-3
    assert rel_err(dg, gr.grad) < 1e-3 and rel_err(db, br.grad) < 1e-3
    # column sums of the produced dx (bias gradient of the upstream Linear), accumulated in place
    assert torch.allclose(dbias, 1.0 + (xr.grad + res.float()).sum(0), rtol=2e-3, atol=2e-2 * math.sqrt(M))


@pytest.mark.parametrize("M,D,split", [(45312, 768, 32768), (1000, 768, 600), (777, 512, 1), (300, 1024, 299),
                                       (500, 128, 200), (64, 768, 0), (64, 768, 64), (3, 512, 2)])
def test_layernorm_two_affine_sets_one_launch(M, D, split):
    """Rows [0, split) with the audio LayerNorm set, rows [split, M) with the video set, in ONE launch
    (Block.forward's norm{1,2}_a / norm{1,2}_v switch, cav_mae_base.py:151-152,169-170,190-191)."""
    x = rnd(M, D, scale=2.0, dtype=torch.bfloat16)
    g0, b0, g1, b1 = 1 + 0.1 * rnd(D, seed=1), 0.1 * rnd(D, seed=2), 1 + 0.2 * rnd(D, seed=3), 0.3 * rnd(D, seed=4)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.layernorm_fwd2(x, g0, b0, split, g1, b1, 1e-5, y, mean, rstd, M, D)
    xr = x.float().requires_grad_(True)
    gs = [t.clone().requires_grad_(True) for t in (g0, b0, g1, b1)]
    ref = torch.cat([F.layer_norm(xr[:split], (D,), gs[0], gs[1], 1e-5), F.layer_norm(xr[split:], (D,), gs[2], gs[3], 1e-5)])
    assert rel_err(y, ref) < 4e-3
    dy, res = rnd(M, D, dtype=torch.bfloat16), rnd(M, D, dtype=torch.bfloat16)
    ref.backward(dy.float())
    dx = torch.empty_like(x)
    dg0, db0, dg1, db1, dbias = (torch.zeros(D, device=DEV) for _ in range(5))
    ops.layernorm_bwd2(dy, x, mean, rstd, g0, dg0, db0, split, g1, dg1, db1, dx, M, D, resid=res, dbias=dbias)
    assert rel_err(dx, xr.grad + res.float()) < 5e-3
    for got, want, rows in ((dg0, gs[0].grad, split), (db0, gs[1].grad, split), (dg1, gs[2].grad, M - split),
                            (db1, gs[3].grad, M - split)):
        if rows == 0:
            assert float(got.abs().max()) == 0.0
        else:
            assert rel_err(got, want) < 1e-3
    assert torch.allclose(dbias, (xr.grad + res.float()).sum(0), rtol=2e-3, atol=2e-2 * math.sqrt(M))


def test_layernorm_rowmap_and_pool_grad():
    n_seq, S, D, stride, off = 4, 5, 128, 12, 7
    M = n_seq * S
    x = rnd(M, D, dtype=torch.bfloat16)
    gamma, beta = 1 + 0.1 * rnd(D), 0.1 * rnd(D)
    ycat = torch.zeros(n_seq * stride, D, dtype=torch.bfloat16, device=DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, 1e-6, ycat, mean, rstd, M, D, seq_len=S, y_seq_stride=stride, y_off=off)
    xr = x.float().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), gamma, beta, 1e-6).reshape(n_seq, S, D)
    got = ycat.reshape(n_seq, stride, D)[:, off:off + S]
    assert rel_err(got, ref) < 4e-3
    pooled = torch.empty(n_seq, D, device=DEV)
    ops.seq_mean_fwd(ycat, pooled, n_seq, S, D, y_seq_stride=stride, y_off=off)
    assert torch.allclose(pooled, got.float().mean(1), rtol=1e-5, atol=1e-5)
    dycat = rnd(n_seq * stride, D, dtype=torch.bfloat16)
    dpool = rnd(n_seq, D)
    (ref * dycat.reshape(n_seq, stride, D)[:, off:off + S].float()).sum().backward(retain_graph=True)
    (ref.mean(1) * dpool).sum().backward()
    dx = torch.empty_like(x)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.layernorm_bwd(dycat, x, mean, rstd, gamma, dx, dg, db, M, D, dpool=dpool, pool_scale=1.0 / S, seq_len=S,
                      y_seq_stride=stride, y_off=off)
    assert rel_err(dx, xr.grad) < 5e-3


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("n_seq,S,H,hd", [(3, 49, 12, 64), (2, 128, 12, 64), (2, 177, 12, 64), (2, 708, 16, 32),
                                          (1, 64, 2, 32), (2, 1, 2, 64),
                                          # persistent single-tile tcgen05 kernels (head_dim 64, S <= 128): pair packing
                                          # with an odd sequence count, slot edges, more tiles than CTAs
                                          (1, 49, 2, 64), (5, 64, 2, 64), (3, 65, 2, 64), (2, 100, 4, 64), (3, 127, 2, 64),
                                          (7, 17, 2, 64), (64, 128, 12, 64), (101, 49, 12, 64),
                                          # tcgen05 forward range (head_dim 32, 256 <= S <= 768): unit / block edges
                                          (1, 256, 4, 32), (3, 257, 2, 32), (2, 300, 2, 32), (1, 511, 2, 32),
                                          (1, 640, 2, 32), (2, 768, 2, 32),
                                          # ViT-H/14 (BASELINE config 5): head_dim 80 at its kept-token counts 164 / 64,
                                          # block / tile edges, and its 913-token sequence at the decoder's head_dim 32
                                          (2, 164, 16, 80), (3, 64, 4, 80), (1, 657, 2, 80), (2, 65, 2, 80), (1, 7, 2, 80),
                                          (2, 228, 2, 80), (1, 913, 2, 32)])
def test_attention_fwd_bwd(n_seq, S, H, hd):
    D = H * hd
    qkv = rnd(n_seq * S, 3 * D, dtype=torch.bfloat16)
    out = torch.empty(n_seq * S, D, dtype=torch.bfloat16, device=DEV)
    lse2 = torch.empty(n_seq, H, S, device=DEV)
    ops.attention_fwd(qkv, out, lse2, n_seq, S, H, hd)
    q5 = qkv.float().reshape(n_seq, S, 3, H, hd).permute(2, 0, 3, 1, 4).detach().requires_grad_(True)
    q, k, v = q5[0], q5[1], q5[2]
    att = (q @ k.transpose(-2, -1)) * hd ** -0.5
    ref = (att.softmax(-1) @ v).transpose(1, 2).reshape(n_seq * S, D)
    assert rel_err(out, ref) < 8e-3
    ref_lse2 = torch.logsumexp(att, -1) / math.log(2)
    assert torch.allclose(lse2, ref_lse2, rtol=1e-4, atol=1e-3)
    dout = rnd(n_seq * S, D, dtype=torch.bfloat16)
    ref.backward(dout.float())
    ref_dqkv = q5.grad.permute(1, 3, 0, 2, 4).reshape(n_seq * S, 3 * D)
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty(n_seq, H, S, device=DEV)
    dbias = torch.ones(3 * D, device=DEV)
    ops.attention_bwd(qkv, out, dout, lse2, delta, dqkv, n_seq, S, H, hd, dbias=dbias)
    # fused qkv-bias gradient: 1 + column sums of dQKV (accumulated; fp32 sums of the pre-rounding values)
    ref_db = 1 + ref_dqkv.sum(0)
    assert float((dbias - ref_db).abs().max()) <= 2e-2 * float(ref_dqkv.abs().sum(0).max()) + 1e-3
    for i, name in enumerate("qkv"):
        got, want = dqkv[:, i * D:(i + 1) * D].float(), ref_dqkv[:, i * D:(i + 1) * D]
        if S == 1 and name in "qk":   # softmax over one key is constant: dQ = dK = 0 exactly in the reference
            assert float(got.abs().max()) < 2e-2 * float(dout.float().abs().max()), name
            continue
        e = rel_err(got, want)
        assert e < 2e-2, (name, e)   # bf16 P / dS operands in the tensor-core products


@pytest.mark.parametrize("n_seq,S", [(3, 49), (2, 128), (5, 80)])
def test_attention_small_writes_only_its_rows(n_seq, S):
    """The encoder groups are row slices of shared [tokens, 3D] / [tokens, D] buffers (audio rows, then video rows): the
    persistent kernels load 64-row boxes that reach into the neighbouring group and must never WRITE outside their own
    rows. Guard rows before and after the slice must keep their contents, and the result must not depend on them."""
    H, hd = 2, 64
    D, G = H * hd, 70
    rows = n_seq * S
    big_qkv = rnd(rows + 2 * G, 3 * D, dtype=torch.bfloat16, seed=5)
    big_out = torch.full((rows + 2 * G, D), 7.0, dtype=torch.bfloat16, device=DEV)
    big_dqkv = torch.full((rows + 2 * G, 3 * D), 9.0, dtype=torch.bfloat16, device=DEV)
    big_dout = rnd(rows + 2 * G, D, dtype=torch.bfloat16, seed=6)
    qkv, out, dqkv, dout = big_qkv[G:G + rows], big_out[G:G + rows], big_dqkv[G:G + rows], big_dout[G:G + rows]
    lse2 = torch.empty(n_seq, H, S, device=DEV)
    delta = torch.empty(n_seq, H, S, device=DEV)
    ops.attention_fwd(qkv, out, lse2, n_seq, S, H, hd)
    ops.attention_bwd(qkv, out, dout, lse2, delta, dqkv, n_seq, S, H, hd)
    for big, fill in ((big_out, 7.0), (big_dqkv, 9.0)):
        assert bool((big[:G] == fill).all()) and bool((big[G + rows:] == fill).all())
    # same inputs in a stand-alone buffer (different neighbours: zero fill) give the same result
    out2, dqkv2 = torch.empty(rows, D, dtype=torch.bfloat16, device=DEV), torch.empty(rows, 3 * D, dtype=torch.bfloat16, device=DEV)
    lse2b = torch.empty_like(lse2)
    ops.attention_fwd(qkv.contiguous(), out2, lse2b, n_seq, S, H, hd)
    ops.attention_bwd(qkv.contiguous(), out2, dout.contiguous(), lse2b, delta, dqkv2, n_seq, S, H, hd)
    assert torch.equal(out2, out) and torch.equal(dqkv2, dqkv) and torch.equal(lse2, lse2b)


@pytest.mark.parametrize("S", [708, 400])
def test_attention_fwd_rising_scores_rescale_path(S):
    """Scores that keep growing along the key axis (by far more than the 2^8 slack of the lazy running maximum) force
    the tcgen05 forward to raise its maximum and rescale O and the row sums in tensor memory in EVERY unit; sharp and
    flat rows are mixed so that lanes with and without a raise share a warp."""
    n_seq, H, hd = 2, 4, 32
    D = H * hd
    g = torch.Generator(device="cpu").manual_seed(11)
    q = torch.randn(n_seq, S, H, hd, generator=g)
    k = torch.randn(n_seq, S, H, hd, generator=g)
    v = torch.randn(n_seq, S, H, hd, generator=g)
    ramp = torch.linspace(0.2, 6.0, S).view(1, S, 1, 1)             # later keys are longer ...
    k = k * ramp
    q[:, ::3] = q[:, ::3] * 4.0                                       # ... and every third query row is sharp
    k[:, :, :, 0] += ramp[..., 0] * 3.0                               # a shared direction makes the growth systematic
    q[:, :, :, 0] = q[:, :, :, 0].abs() + 1.0
    qkv = torch.stack([q, k, v], 2).reshape(n_seq * S, 3 * D).to(torch.bfloat16).to(DEV)
    out = torch.empty(n_seq * S, D, dtype=torch.bfloat16, device=DEV)
    lse2 = torch.empty(n_seq, H, S, device=DEV)
    ops.attention_fwd(qkv, out, lse2, n_seq, S, H, hd)
    q5 = qkv.float().reshape(n_seq, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    att = (q5[0] @ q5[1].transpose(-2, -1)) * hd ** -0.5
    # the premise of the test: the running row maximum is raised by more than 8 (log2) after the first 128-key unit
    first = att[..., :128].max(-1).values
    assert float(((att.max(-1).values - first) / math.log(2) > 8).float().mean()) > 0.5
    ref = (att.softmax(-1) @ q5[2]).transpose(1, 2).reshape(n_seq * S, D)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 8e-3
    ref_lse2 = torch.logsumexp(att, -1) / math.log(2)
    err = (lse2 - ref_lse2).abs()
    assert float(err.max()) < 6e-3, (float(err.max()), float(err.mean()))   # row sums are taken over the bf16-rounded P


def test_attention_fwd_mixed_bound_and_exact_blocks():
    """Query blocks whose Cauchy-Schwarz score bound is small take the no-maximum path of the tcgen05 forward, the others
    the lazy-maximum path: alternate them inside one head (and inside one lane quarter's neighbourhood) and compare."""
    n_seq, S, H, hd = 2, 708, 2, 32
    D = H * hd
    g = torch.Generator(device="cpu").manual_seed(5)
    q = torch.randn(n_seq, S, H, hd, generator=g) * 0.5
    k = torch.randn(n_seq, S, H, hd, generator=g)
    v = torch.randn(n_seq, S, H, hd, generator=g)
    q[:, 128:256] *= 40.0            # query block 1: bound far above the limit
    q[:, 300:310] *= 60.0            # a few rows of block 2 (one lane quarter exact, the others fast)
    q[:, 640:] *= 25.0               # the ragged last block
    qkv = torch.stack([q, k, v], 2).reshape(n_seq * S, 3 * D).to(torch.bfloat16).to(DEV)
    out = torch.empty(n_seq * S, D, dtype=torch.bfloat16, device=DEV)
    lse2 = torch.empty(n_seq, H, S, device=DEV)
    ops.attention_fwd(qkv, out, lse2, n_seq, S, H, hd)
    q5 = qkv.float().reshape(n_seq, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    att = (q5[0] @ q5[1].transpose(-2, -1)) * hd ** -0.5
    ref = (att.softmax(-1) @ q5[2]).transpose(1, 2).reshape(n_seq * S, D)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 8e-3
    err = (lse2 - torch.logsumexp(att, -1) / math.log(2)).abs()
    assert float(err.max()) < 6e-3, float(err.max())


@pytest.mark.parametrize("M,N,K", [(1000, 512, 256), (4133, 2048, 512), (300, 200, 128)])
def test_gemm_fused_colsum_of_dgelu_output(M, N, K):
    """The dGELU-epilogue GEMM (fc2's dgrad) also accumulates the column sums of its output = the fc1 bias gradient
    (ragged M and N tails, several tiles per CTA, accumulation into a non-zero buffer)."""
    a = rnd(M, K, dtype=torch.bfloat16)
    w = rnd(K, N, dtype=torch.bfloat16) * K ** -0.5          # MN-major B: out[m, n] = sum_k a[m, k] w[k, n]
    z = rnd(M, N, dtype=torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    cs = torch.full((N,), 3.0, device=DEV)
    ops.gemm(a, w, out, M, N, K, b_major=ops.MAJOR_MN, dgelu_aux=z, colsum=cs)
    plain = torch.empty_like(out)
    ops.gemm(a, w, plain, M, N, K, b_major=ops.MAJOR_MN, dgelu_aux=z)
    assert torch.equal(out, plain)                            # the output itself is untouched
    ref = 3.0 + out.float().sum(0)
    scale = float(out.float().abs().sum(0).max())
    assert float((cs - ref).abs().max()) <= 4e-3 * scale + 1e-3   # fp32 sums of the pre-rounding values


# ------------------------------------------------------------------------------------------------ losses
def test_mae_loss_fwd_bwd():
    d = O.VIT_B
    B = 2
    audio, img = rnd(B, d.audio_len, d.mel), rnd(B, 3, d.img, d.img)
    pa, pv = rnd(B * d.Ta, 256, dtype=torch.bfloat16), rnd(B * d.Tv, 768, dtype=torch.bfloat16)
    ma = (torch.rand(B, d.Ta, device=DEV) < 0.75).float()
    mv = (torch.rand(B, d.Tv, device=DEV) < 0.75).float()
    loss = torch.zeros(2, device=DEV)
    ops.mae_loss_fwd(pa, audio, ma, 0, B, 16, 1, d.audio_len, d.mel, float(ma.sum()), loss[0:1])
    ops.mae_loss_fwd(pv, img, mv, 1, B, 16, 3, d.img, d.img, float(mv.sum()), loss[1:2])
    par = pa.float().cpu().reshape(B, d.Ta, 256).requires_grad_(True)
    pvr = pv.float().cpu().reshape(B, d.Tv, 768).requires_grad_(True)
    ra = O.mae_loss(O.patchify_target_audio(audio.cpu(), d), par, ma.cpu())
    rv = O.mae_loss(O.patchify_target_video(img.cpu(), d), pvr, mv.cpu())
    assert float(loss[0]) == pytest.approx(float(ra), rel=1e-4)
    assert float(loss[1]) == pytest.approx(float(rv), rel=1e-4)
    up = torch.tensor([3.0], device=DEV)
    (3.0 * ra).backward(); (3.0 * rv).backward()
    dpa, dpv = torch.empty_like(pa), torch.empty_like(pv)
    ops.mae_loss_bwd(pa, audio, ma, 0, B, 16, 1, d.audio_len, d.mel, float(ma.sum()), up, dpa)
    ops.mae_loss_bwd(pv, img, mv, 1, B, 16, 3, d.img, d.img, float(mv.sum()), up, dpv)
    assert rel_err(dpa.cpu(), par.grad.reshape(-1, 256)) < 4e-3
    assert rel_err(dpv.cpu(), pvr.grad.reshape(-1, 768)) < 4e-3


@pytest.mark.parametrize("N,D,bidirect", [(10, 768, True), (300, 128, True), (64, 768, False), (520, 768, True), (2048, 768, True)])
def test_infonce_fwd_bwd(N, D, bidirect):
    ea, ev = rnd(N, D, seed=3), rnd(N, D, seed=4)
    ev = ev + 0.5 * ea  # correlated pairs => non-trivial accuracy
    ws = ops.infonce_workspace(N, D, DEV)
    res = torch.zeros(2, device=DEV)
    ops.infonce_fwd(ea, ev, 0.05, bidirect, ws, res[0:1], res[1:2])
    ar, vr = ea.cpu().requires_grad_(True), ev.cpu().requires_grad_(True)
    loss, acc = O.contrastive(ar, vr, bidirect=bidirect)
    assert float(res[0]) == pytest.approx(float(loss), rel=1e-4, abs=1e-5)   # fp32 path: 1e-4
    assert float(res[1]) == pytest.approx(float(acc), abs=1e-6)
    (0.7 * 2.0 * loss).backward()
    row0, rows = N // 3, N - N // 3 - 1
    d_ea, d_ev = torch.empty(rows, D, device=DEV), torch.empty(rows, D, device=DEV)
    scratch = torch.empty(2 * rows * D, device=DEV)
    up = torch.tensor([2.0], device=DEV)
    ops.infonce_bwd(N, D, 0.05, bidirect, 0.7, up, ws, row0, rows, scratch, d_ea, d_ev)
    assert rel_err(d_ea.cpu(), ar.grad[row0:row0 + rows]) < 1e-3
    assert rel_err(d_ev.cpu(), vr.grad[row0:row0 + rows]) < 1e-3


# ------------------------------------------------------------------------------------------------ optimizer
def test_adam_matches_torch_optim():
    n = 4096 * 5
    p0, g1, g2 = rnd(n, seed=1), rnd(n, seed=2), rnd(n, seed=3)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=2e-4, weight_decay=5e-7, betas=(0.95, 0.999))
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    shadow = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    for step, g in enumerate((g1, g2), 1):
        ref_p.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, shadow, 2e-4, 0.95, 0.999, 1e-8, 5e-7, step)
    assert torch.allclose(p, ref_p.data, rtol=1e-5, atol=1e-7)
    assert torch.equal(shadow, p.to(torch.bfloat16))
    # found_inf skips the step; inv_scale unscales
    flag = torch.ones(1, device=DEV)
    before = p.clone()
    ops.adam_step(p, g1, m, v, None, 2e-4, 0.95, 0.999, 1e-8, 5e-7, 3, found_inf=flag)
    assert torch.equal(p, before)
    flag.zero_()
    bad = g1.clone(); bad[7] = float("inf")
    ops.found_inf(bad, flag)
    assert float(flag) == 1.0


def test_colsum_and_cast():
    M, N = 1234, 2304
    dy = rnd(M, N, dtype=torch.bfloat16)
    out = torch.ones(N, device=DEV)
    ops.colsum(dy, out, M, N)
    assert torch.allclose(out, 1 + dy.float().sum(0), rtol=1e-4, atol=1e-3)
    src = rnd(100003)
    dst = torch.empty(100003, dtype=torch.bfloat16, device=DEV)
    ops.cast_f32_to_bf16(src, dst)
    assert torch.equal(dst, src.to(torch.bfloat16))
    table = torch.zeros(16, 64, device=DEV)
    idx = torch.randint(0, 16, (200,), device=DEV, dtype=torch.int32)
    d2 = rnd(200, 64, dtype=torch.bfloat16)
    ops.scatter_add_rows(d2, idx, table, 2.0)
    ref = torch.zeros(16, 64, device=DEV).index_add_(0, idx.long(), 2.0 * d2.float())
    assert torch.allclose(table, ref, rtol=1e-5, atol=1e-5)


def test_attention_fwd_persistent_pingpong_variant():
    """The experimental persistent forward (attn_fwd_pp_kernel: one CTA per SM, two softmax groups passing a MUFU token,
    TMA-fed K/V double-buffered across heads) is selected by an environment variable read once per process, so it is
    checked in a child process against torch's SDPA on the decoder shape."""
    import subprocess
    import sys
    code = r'''
import torch, torch.nn.functional as F
from avsiam_b200 import ops
torch.manual_seed(0)
n_seq, S, H, hd = 3, 708, 16, 32
D = H * hd
qkv = (torch.randn(n_seq * S, 3 * D, device="cuda") * 0.7).to(torch.bfloat16)
out = torch.empty(n_seq * S, D, dtype=torch.bfloat16, device="cuda")
lse = torch.empty(n_seq, H, S, dtype=torch.float32, device="cuda")
ops.attention_fwd(qkv, out, lse, n_seq, S, H, hd)
q, k, v = (qkv.float().view(n_seq, S, 3, H, hd).permute(2, 0, 3, 1, 4))
ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n_seq * S, D)
err = float((out.float() - ref).norm() / ref.norm())
assert err < 1e-2, err
ref_lse = torch.logsumexp((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1) * 1.4426950408889634   # log2 domain
assert float((lse - ref_lse).abs().max()) < 2e-2
print("ok", err)
'''
    env = dict(os.environ, AVS_ATTN_FWD_VARIANT="2")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
