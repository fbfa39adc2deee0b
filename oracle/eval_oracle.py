"""TEST INFRASTRUCTURE ONLY — CPU restatement of the evaluation-path arithmetic.

  sim_mat / compute_metrics  follow src/retrieval.py:27-52 (cosine similarity per pair; rank of the diagonal element in
                             each row's descending sort, every tied position listed; R@k = share of listed positions
                             below k; MR = median position + 1)
  bce_with_logits            nn.BCEWithLogitsLoss() as built at src/traintest_ft_base.py:106-107 (mean over B*C)
  cross_entropy_prob         nn.CrossEntropyLoss() (:108-109) applied to float label vectors (dataloader.py:497-503):
                             mean over the batch of -sum_c y_c log_softmax(x)_c
  distributed_concat         src/traintest_cavmae_base.py:21-26

tests/test_oracle_golden.py pins this file to tests/golden/eval_path.pt, which oracle/make_golden_eval.py produced by
executing the reference's own get_sim_mat / compute_metrics (function bodies loaded from /root/reference/src/retrieval.py
at generation time) and torch's loss modules.  Product code never imports this module.
"""
from __future__ import annotations

import numpy as np


def sim_mat(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    na = np.sqrt((a * a).sum(1, dtype=np.float32))
    nb = np.sqrt((b * b).sum(1, dtype=np.float32))
    return ((a @ b.T) / (na[:, None] * nb[None, :])).astype(np.float64)


def diagonal_positions(x: np.ndarray) -> np.ndarray:
    """Sorted positions (descending order) that hold each row's diagonal value, row after row."""
    out = []
    for i in range(x.shape[0]):
        d = x[i, i]
        g = int((x[i] > d).sum())
        e = int((x[i] == d).sum())
        out.extend(range(g, g + e))
    return np.asarray(out, np.int64)


def compute_metrics(x: np.ndarray) -> dict:
    ind = diagonal_positions(np.asarray(x))
    return {"R1": float((ind == 0).sum()) / len(ind), "R5": float((ind < 5).sum()) / len(ind),
            "R10": float((ind < 10).sum()) / len(ind), "MR": float(np.median(ind) + 1)}


def bce_with_logits(x: np.ndarray, y: np.ndarray):
    """-> (loss, dloss/dx) in float64."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    loss = np.maximum(x, 0) - x * y + np.log1p(np.exp(-np.abs(x)))
    sig = 1.0 / (1.0 + np.exp(-x))
    return float(loss.mean()), (sig - y) / x.size


def cross_entropy_prob(x: np.ndarray, y: np.ndarray):
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    z = x - x.max(1, keepdims=True)
    lse = np.log(np.exp(z).sum(1, keepdims=True))
    logp = z - lse
    loss = -(y * logp).sum(1).mean()
    return float(loss), (np.exp(logp) * y.sum(1, keepdims=True) - y) / x.shape[0]


def distributed_concat(per_rank, num_total_examples: int) -> np.ndarray:
    return np.concatenate(list(per_rank), 0)[:num_total_examples]
