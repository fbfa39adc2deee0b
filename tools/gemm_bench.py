"""Developer aid: time the fused-epilogue GEMM shapes of the decoder / encoder blocks alone (CUDA events), sweeping the
epilogue tuning knobs (input-tile ring depth, TMEM load prefetch).  python tools/gemm_bench.py [variant-substring]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import _lib, ops


def t(fn, n=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


only = sys.argv[1] if len(sys.argv) > 1 else ""
tunings = [(2, 0), (2, 1), (3, 0), (3, 1), (4, 1)]
print("tuning columns (in_depth, tmem_prefetch):", tunings)
for (M, N, K, name) in [(181248, 2048, 512, "dec fc1/fc2"), (45312, 3072, 768, "enc fc1/fc2")]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.randn(M, N, device="cuda").bfloat16()
    dy = torch.randn(M, K, device="cuda").bfloat16()   # stands in for the [M, N_out] upstream gradient (K-major)
    w2 = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    dh = torch.empty_like(out)
    cs = torch.zeros(N, device="cuda")
    # fc2 forward: [M, N] x [K, N]^T -> [M, K] + residual
    wf2 = (torch.randn(K, N, device="cuda") * 0.02).bfloat16()
    res = torch.randn(M, K, device="cuda").bfloat16()
    o2 = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    b2 = torch.randn(K, device="cuda")
    runs = {
        "fc1 plain+bias": (lambda: ops.gemm(x, w, out, M, N, K, bias=bias), 2.0 * M * N * K),
        "fc1 gelu-noaux": (lambda: ops.gemm(x, w, out, M, N, K, bias=bias, gelu=True), 2.0 * M * N * K),
        "fc1 gelu+aux": (lambda: ops.gemm(x, w, out, M, N, K, bias=bias, gelu=True, aux_out=dh), 2.0 * M * N * K),
        "fc1 resid": (lambda: ops.gemm(x, w, out, M, N, K, bias=bias, resid=pre), 2.0 * M * N * K),
        # dgelu GEMM: dH[M, N] = (dY[M, K] W2[K(out), N]) * gelu'(pre): B operand MN-major
        "fc2 dgrad dgelu": (lambda: ops.gemm(dy, w2, dh, M, N, K, b_major=1, dgelu_aux=pre), 2.0 * M * N * K),
        "fc2 dgrad dgelu+colsum": (lambda: ops.gemm(dy, w2, dh, M, N, K, b_major=1, dgelu_aux=pre, colsum=cs), 2.0 * M * N * K),
        "fc2 fwd bias+resid": (lambda: ops.gemm(pre, wf2, o2, M, K, N, bias=b2, resid=res), 2.0 * M * N * K),
        "fc2 fwd bias": (lambda: ops.gemm(pre, wf2, o2, M, K, N, bias=b2), 2.0 * M * N * K),
    }
    for k, (fn, flops) in runs.items():
        if only and only not in k:
            continue
        cells = []
        for (depth, pf) in tunings:
            _lib.lib().avs_gemm_set_tuning(depth, pf)
            ms = t(fn)
            cells.append(f"{ms:.3f} ms {flops / ms / 1e9:5.0f}")
        print(f"{name} M={M} N={N} K={K} {k:24s}: " + " | ".join(cells), flush=True)
_lib.lib().avs_gemm_set_tuning(3, 1)
