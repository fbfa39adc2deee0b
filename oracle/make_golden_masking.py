"""TEST INFRASTRUCTURE ONLY — structured-masking fixture from the UNMODIFIED reference method.

    python -m oracle.make_golden_masking        (only works where /root/reference exists)

Runs the reference's own `CAVMAE_BASE.random_masking_structured` (cav_mae_base.py:392-439) for the modes 'time', 'freq'
and 'tf' with Python's `random` seeded (the time-column / frequency-row draws) and `torch.rand` fed a recorded noise
tensor, and stores its outputs (x_masked, mask, ids_restore). The CUDA path is then run on the GPU box with the same
`random.seed` and the same noise and must reproduce them bit for bit (tests/test_kernels_gpu.py).
`order_free` marks cases in which more tokens are forced to 1.1 than are removed: there the kept set depends on how
the sort orders equal keys, which torch.argsort leaves unspecified; the fixture records what the reference produced
here and whether that equals the stable (lower index first) order the CUDA kernel defines.
"""
from __future__ import annotations

import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    torch.manual_seed(0)
    model = ref_shim.build_reference_model()
    cases = []
    real_rand = torch.rand
    for ci, (mode, ratio, t, f, N) in enumerate([
            ("time", 0.75, 64, 8, 3), ("freq", 0.75, 64, 8, 3), ("tf", 0.75, 64, 8, 4), ("tf", 0.2, 64, 8, 2),
            ("tf", 0.4, 64, 8, 2), ("tf", 0.6000000000000001, 64, 8, 3), ("tf", 0.8, 64, 8, 3), ("tf", 0.0, 64, 8, 2),
            ("time", 0.5, 16, 2, 5), ("freq", 0.5, 16, 2, 5), ("tf", 0.75, 16, 2, 5)]):
        L, D = t * f, 8
        g = torch.Generator().manual_seed(500 + ci)
        x = torch.randn(N, L, D, generator=g)
        noise = real_rand(N, L, generator=g)
        seed = 9000 + ci
        random.seed(seed)
        torch.rand = lambda *a, **k: noise.clone()
        try:
            xm, mask, ids_restore = model.random_masking_structured(x, ratio, t=t, f=f, mode=mode)
        finally:
            torch.rand = real_rand
        # what a stable sort of the same forced noise gives (the CUDA kernel's definition of ties)
        len_keep = int(L * (1 - ratio))
        forced = (mask.sum(1) * 0).long()
        ids_shuffle_ref = torch.argsort(ids_restore, dim=1)
        kept_ref = ids_shuffle_ref[:, :len_keep]
        # reconstruct the forced noise through the kept / removed structure is not needed: count forced from the draws
        random.seed(seed)
        nz = noise.clone().reshape(N, f, t)
        if mode == "time":
            for i in range(N):
                for k in random.sample(range(t), int(t * ratio)):
                    nz[i, :, k] = 1.1
        elif mode == "freq":
            for i in range(N):
                for k in random.sample(range(f), int(f * ratio)):
                    nz[i, k, :] = 1.1
        else:
            for i in range(N):
                for k in random.sample(range(t), int(t * ratio * 0.7)):
                    nz[i, :, k] = 1.1
            for i in range(N):
                for k in random.sample(range(f), int(f * ratio * 0.7)):
                    nz[i, k, :] = 1.1
        nz = nz.reshape(N, L)
        n_forced = (nz == 1.1).sum(1)
        order_free = bool((n_forced > L - len_keep).any())
        stable = torch.argsort(nz, dim=1, stable=True)
        same_as_stable = bool(torch.equal(torch.argsort(stable, dim=1), ids_restore))
        cases.append({"mode": mode, "ratio": ratio, "t": t, "f": f, "N": N, "seed": seed, "x": x, "noise": noise,
                      "x_masked": xm, "mask": mask, "ids_restore": ids_restore, "order_free": order_free,
                      "same_as_stable": same_as_stable, "n_forced": n_forced})
        print(f"[golden] {mode} r={ratio:.2f} t={t} f={f} N={N}: keep={len_keep} forced={n_forced.tolist()} "
              f"order_free={order_free} same_as_stable={same_as_stable}")
    torch.save(cases, os.path.join(GOLDEN_DIR, "masking_structured.pt"))
    print("[golden] wrote masking_structured.pt")


if __name__ == "__main__":
    main()
