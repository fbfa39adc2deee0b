"""TEST INFRASTRUCTURE ONLY — golden vectors for the evaluation statistics, produced by the scikit-learn calls the
reference makes (src/utilities/stats.py:11-68).      python -m oracle.make_golden_stats
"""
import os
import sys

import numpy as np
import torch
from sklearn import metrics

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.make_golden import GOLDEN_DIR  # noqa: E402


def synth_eval(seed: int, n: int, c: int, ties: bool):
    g = np.random.default_rng(seed)
    target = (g.random((n, c)) < 0.03).astype(np.float32)
    target[np.arange(n), g.integers(0, c, n)] = 1.0               # at least one label per sample
    output = (g.standard_normal((n, c)) + 1.5 * target).astype(np.float32)
    if ties:
        output = np.round(output * 4) / 4                         # heavy score ties
    output = 1.0 / (1.0 + np.exp(-output))                        # the loop applies sigmoid before calculate_stats
    return output.astype(np.float32), target


def main():
    cases = []
    for seed, n, c, ties in ((1, 700, 37, False), (2, 513, 20, True), (3, 64, 5, True)):
        output, target = synth_eval(seed, n, c, ties)
        ap = np.array([metrics.average_precision_score(target[:, k], output[:, k], average=None) for k in range(c)])
        auc = np.array([metrics.roc_auc_score(target[:, k], output[:, k], average=None) for k in range(c)])
        acc = metrics.accuracy_score(np.argmax(target, 1), np.argmax(output, 1))
        cases.append({"seed": seed, "n": n, "c": c, "ties": ties, "AP": torch.from_numpy(ap), "auc": torch.from_numpy(auc),
                      "acc": float(acc)})
        print(f"[golden-stats] seed={seed}: mAP={ap.mean():.6f} mAUC={auc.mean():.6f} acc={acc:.4f}")
    torch.save(cases, os.path.join(GOLDEN_DIR, "eval_stats.pt"))


if __name__ == "__main__":
    main()
