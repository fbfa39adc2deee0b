"""TEST INFRASTRUCTURE ONLY — golden vectors for the evaluation path.       python -m oracle.make_golden_eval

Retrieval: the reference's own `get_sim_mat` / `compute_metrics` (src/retrieval.py:27-52).  The module cannot be
imported (its top level loads checkpoints from the authors' file system), so only those function definitions are
compiled out of the file, in memory, at generation time.  Losses: torch's nn.BCEWithLogitsLoss / nn.CrossEntropyLoss,
the modules traintest_ft_base.py:106-109 instantiates.
"""
import ast
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.make_golden import GOLDEN_DIR  # noqa: E402

REF_RETRIEVAL = "/root/reference/src/retrieval.py"


def reference_retrieval_functions():
    tree = ast.parse(open(REF_RETRIEVAL).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef)
            and n.name in ("get_similarity", "get_sim_mat", "compute_metrics")]
    ns = {"np": np, "dot": np.dot, "norm": np.linalg.norm}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF_RETRIEVAL, "exec"), ns)
    return ns["get_sim_mat"], ns["compute_metrics"]


def synth_features(seed: int, n: int, d: int, dup: int, noise: float = 1.6):
    """Paired audio / video features: video = audio + noise, so the diagonal usually — not always — wins.
    `dup` > 0 copies some video rows onto others to create exact score ties with the diagonal."""
    g = np.random.default_rng(seed)
    a = g.standard_normal((n, d)).astype(np.float32)
    v = (a + noise * g.standard_normal((n, d))).astype(np.float32)
    for k in range(dup):
        v[(7 * k + 3) % n] = v[(7 * k + 4) % n]
    return a, v


def synth_logits(seed: int, B: int, C: int, smooth: float):
    g = np.random.default_rng(seed)
    x = (3.0 * g.standard_normal((B, C))).astype(np.float32)
    y = np.zeros((B, C), np.float32)
    for b in range(B):
        for c in g.choice(C, size=int(g.integers(1, 4)), replace=False):
            y[b, c] = 1.0 - smooth                                   # dataloader.py:497-503 (label_smooth)
    return x, y


RETRIEVAL_CASES = ((11, 96, 64, 0, 1.6), (12, 130, 48, 9, 6.0), (13, 33, 768, 2, 1.6))
LOSS_CASES = ((21, 16, 527, 0.0), (22, 9, 309, 0.1), (23, 3, 5, 0.0))


def main():
    get_sim_mat, compute_metrics = reference_retrieval_functions()
    out = {"retrieval": [], "loss": []}
    for seed, n, d, dup, noise in RETRIEVAL_CASES:
        a, v = synth_features(seed, n, d, dup, noise)
        sim = get_sim_mat(torch.from_numpy(a), torch.from_numpy(v))          # the reference feeds CPU torch tensors
        m = compute_metrics(sim)
        out["retrieval"].append({"seed": seed, "n": n, "d": d, "dup": dup, "noise": noise, "sim": torch.from_numpy(sim).float(),
                                 "metrics": {k: float(x) for k, x in m.items()}})
        print(f"[golden-eval] retrieval seed={seed}: {m}")
    for seed, B, C, smooth in LOSS_CASES:
        x, y = synth_logits(seed, B, C, smooth)
        rec = {"seed": seed, "B": B, "C": C, "smooth": smooth}
        for name, fn in (("bce", torch.nn.BCEWithLogitsLoss()), ("ce", torch.nn.CrossEntropyLoss())):
            xt = torch.from_numpy(x).double().requires_grad_(True)
            loss = fn(xt, torch.from_numpy(y).double())
            loss.backward()
            rec[name] = float(loss.detach())
            rec[name + "_grad"] = xt.grad.float()
        out["loss"].append(rec)
        print(f"[golden-eval] loss seed={seed}: bce={rec['bce']:.6f} ce={rec['ce']:.6f}")
    torch.save(out, os.path.join(GOLDEN_DIR, "eval_path.pt"))


if __name__ == "__main__":
    main()
