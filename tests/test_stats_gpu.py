"""GPU evaluation statistics vs scikit-learn's own outputs (tests/golden/eval_stats.pt, the calls of
src/utilities/stats.py:11-68) and vs the oracle on a larger sparse multi-label problem. AP / AUC to fp32 rounding."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import avsiam_b200  # noqa: E402
from oracle import stats_oracle as S  # noqa: E402
from oracle.make_golden_stats import synth_eval  # noqa: E402


def test_stats_match_sklearn_golden(golden_dir):
    for c in torch.load(os.path.join(golden_dir, "eval_stats.pt"), weights_only=False):
        o, t = synth_eval(c["seed"], c["n"], c["c"], c["ties"])
        st = avsiam_b200.calculate_stats(torch.from_numpy(o).cuda(), torch.from_numpy(t).cuda())
        assert torch.allclose(st["AP"].cpu().double(), c["AP"], atol=2e-6)
        assert torch.allclose(st["auc"].cpu().double(), c["auc"], atol=2e-6)
        assert float(st["acc"]) == pytest.approx(c["acc"], abs=1e-7)


def test_stats_large_sparse_and_degenerate_classes():
    o, t = synth_eval(9, 3000, 64, True)
    t[:, 5] = 0.0                      # a class without positives
    t[:, 6] = 1.0                      # a class without negatives
    st = avsiam_b200.calculate_stats(torch.from_numpy(o).cuda(), torch.from_numpy(t).cuda())
    ap, auc, acc = S.calculate_stats(o, t)
    keep = np.ones(64, bool); keep[[5, 6]] = False
    assert np.abs(st["AP"].cpu().numpy()[keep] - ap[keep]).max() < 2e-6
    assert np.abs(st["auc"].cpu().numpy()[keep] - auc[keep]).max() < 2e-6
    assert float(st["auc"][5]) == -1.0 and float(st["auc"][6]) == -1.0 and float(st["AP"][5]) == 0.0
    assert float(st["acc"]) == pytest.approx(acc, abs=1e-7)
    ref_list = avsiam_b200.stats.as_reference_list(st)
    assert len(ref_list) == 64 and set(ref_list[0]) == {"AP", "auc", "acc"}
