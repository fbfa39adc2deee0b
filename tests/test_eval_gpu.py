"""Evaluation path on the GPU against
  (1) golden vectors from the reference's own get_sim_mat / compute_metrics (src/retrieval.py:27-52) and from torch's
      BCEWithLogitsLoss / CrossEntropyLoss (traintest_ft_base.py:106-109)            tests/golden/eval_path.pt
  (2) the CPU oracle on larger seeded inputs and through the validation loops.
Similarities to fp32 rounding (2e-6); ranks / recall / median rank exact; losses 1e-6 relative; gradients 1e-7."""
import dataclasses
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from avsiam_b200 import CAVMAEFT_BASE, Dims, evaluate, losses  # noqa: E402
from oracle import avsiam_oracle as O  # noqa: E402
from oracle import eval_oracle as E  # noqa: E402
from oracle import stats_oracle as S  # noqa: E402
from oracle.make_golden_eval import synth_features, synth_logits  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "eval_path.pt"), weights_only=False)


def test_cosine_sim_and_metrics_match_reference_golden(golden):
    for c in golden["retrieval"]:
        a, v = synth_features(c["seed"], c["n"], c["d"], c["dup"], c["noise"])
        sim = evaluate.get_sim_mat(torch.from_numpy(a).to(DEV), torch.from_numpy(v).to(DEV))
        assert torch.allclose(sim.cpu(), c["sim"], atol=2e-6, rtol=0)
        # metrics on the reference's own matrix: exact (ties included)
        assert evaluate.compute_metrics(c["sim"].to(DEV)) == c["metrics"]
        # and end to end on ours: duplicated candidate rows stay exact ties, so the listing is the same
        ours = evaluate.compute_metrics(sim)
        assert ours["MR"] == c["metrics"]["MR"]
        for k in ("R1", "R5", "R10"):
            assert ours[k] == pytest.approx(c["metrics"][k], abs=1.5 / c["n"])


def test_retrieval_large_against_oracle():
    a, v = synth_features(5, 1545, 768, 4, 5.0)                    # VGGSound retrieval set: 309 classes x 5 clips
    sim = evaluate.get_sim_mat(torch.from_numpy(a).to(DEV), torch.from_numpy(v).to(DEV))
    ref = E.sim_mat(a, v)
    assert np.abs(sim.cpu().numpy() - ref).max() < 3e-6
    assert evaluate.compute_metrics(torch.from_numpy(ref).float().to(DEV)) == E.compute_metrics(ref.astype(np.float32))
    # rectangular + ragged tile edges
    r = evaluate.get_sim_mat(torch.from_numpy(a[:70]).to(DEV), torch.from_numpy(v[:131, :]).to(DEV))
    assert np.abs(r.cpu().numpy() - ref[:70, :131]).max() < 3e-6


def test_classification_losses_match_torch_golden(golden):
    for c in golden["loss"]:
        x, y = synth_logits(c["seed"], c["B"], c["C"], c["smooth"])
        for name, fn in (("bce", losses.bce_with_logits), ("ce", losses.cross_entropy)):
            xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
            loss = fn(xt, torch.from_numpy(y).to(DEV))
            (3.0 * loss).backward()                                 # upstream gradient is applied
            assert float(loss.detach()) == pytest.approx(c[name], rel=2e-6)
            assert torch.allclose(xt.grad.cpu() / 3.0, c[name + "_grad"], atol=1e-7, rtol=1e-5)
        ol, og = E.bce_with_logits(x, y)
        assert ol == pytest.approx(c["bce"], rel=1e-9)


def test_losses_extreme_logits_and_no_grad():
    x = torch.tensor([[80.0, -80.0, 0.0, 30.0]], device=DEV)
    y = torch.tensor([[1.0, 0.0, 1.0, 0.0]], device=DEV)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(x, y)
    assert float(losses.bce_with_logits(x, y)) == pytest.approx(float(ref), rel=1e-6)
    ref = torch.nn.functional.cross_entropy(x, y)
    assert float(losses.cross_entropy(x, y)) == pytest.approx(float(ref), rel=1e-6)
    assert losses.loss_fn("BCE") is losses.bce_with_logits and losses.loss_fn("CE") is losses.cross_entropy
    with pytest.raises(ValueError):
        losses.loss_fn("MSE")
    with pytest.raises(RuntimeError):
        losses.bce_with_logits(x.cpu(), y.cpu())                    # no CPU path


class _Sampler:
    def __init__(self, n):
        self.dataset = range(n)


def _ft_model(label_dim):
    d = O.TINY
    m = CAVMAEFT_BASE(label_dim=label_dim, dims=Dims(**dataclasses.asdict(d)))
    m.load_state_dict(O.with_aliases(O.init_ft_state(d, label_dim, seed=4)), strict=True)
    return m.to(DEV), d


def _loader(d, n_batches, B, T, C, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        a = torch.randn(B, d.audio_len, d.mel, generator=g)
        v = torch.randn(B, T, 3, d.img, d.img, generator=g)
        y = (torch.rand(B, C, generator=g) < 0.3).float()
        y[:, 0] = 1.0
        out.append((a, v, y))
    return out


def test_validate_mlp_matches_manual_pipeline():
    C = 7
    model, d = _ft_model(C)
    loader = _loader(d, n_batches=3, B=4, T=10, C=C, seed=1)
    stats, loss = evaluate.validate_mlp(model, loader, _Sampler(11), "mm_grad")       # 12 fed, 11 real samples
    stats2, probs, target = evaluate.validate_mlp(model, loader, _Sampler(11), "mm_grad", output_pred=True)
    assert probs.shape == (11, 10, C) and target.shape == (11, C)
    # manual: same model outputs, the oracle's statistics and loss
    with torch.no_grad():
        outs = [model(a.to(DEV), v.to(DEV), "mm_grad", is_eval=True).float().cpu() for a, v, _ in loader]
    logits = torch.cat(outs)[:11]
    tgt = torch.cat([y for _, _, y in loader])[:11]
    p = torch.sigmoid(logits).mean(1).numpy()
    ap, auc, acc = S.calculate_stats(p, tgt.numpy())
    assert np.abs(stats["AP"].cpu().numpy() - ap).max() < 1e-5
    assert np.abs(stats["auc"].cpu().numpy() - auc).max() < 1e-5
    assert float(stats["acc"]) == pytest.approx(acc, abs=1e-6)
    ref_loss = np.mean([E.bce_with_logits(o.mean(1).numpy(), y.numpy())[0] for o, (_, _, y) in zip(outs, loader)])
    assert loss == pytest.approx(ref_loss, rel=1e-5)
    assert torch.equal(stats["AP"], stats2["AP"])


def test_retrieval_result_runs_through_model():
    model, d = _ft_model(5)
    loader = _loader(d, n_batches=2, B=3, T=6, C=5, seed=2)
    r1, r5, r10, mr = evaluate.get_retrieval_result(model, loader, "audio")
    with torch.no_grad():
        fa, fv = zip(*[model(a.to(DEV), v.to(DEV), "retrieval") for a, v, _ in loader])
    fa = torch.cat([f.float().mean(1) for f in fa]).cpu().numpy()
    fv = torch.cat([f.float().mean(1) for f in fv]).cpu().numpy()
    ref = E.compute_metrics(E.sim_mat(fa, fv).astype(np.float32))
    assert (r1, r5, r10, mr) == (ref["R1"], ref["R5"], ref["R10"], ref["MR"])
    r = evaluate.get_retrieval_result(model, loader, "video")
    ref = E.compute_metrics(E.sim_mat(fv, fa).astype(np.float32))
    assert r == (ref["R1"], ref["R5"], ref["R10"], ref["MR"])
