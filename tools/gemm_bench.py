"""Developer aid: time the fused-epilogue GEMM shapes of the decoder / encoder MLP alone (CUDA events)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import ops

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

only = sys.argv[1] if len(sys.argv) > 1 else ""
for (M, N, K, name) in [(181248, 2048, 512, "dec fc1"), (45312, 3072, 768, "enc fc1")]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    dy = torch.randn(M, K, device="cuda").bfloat16()   # stands in for the [M, N_out] upstream gradient (K-major)
    w2 = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    dh = torch.empty_like(out)
    runs = {
        "plain+bias": lambda: ops.gemm(x, w, out, M, N, K, bias=bias),
        "plain": lambda: ops.gemm(x, w, out, M, N, K),
        "gelu-noaux": lambda: ops.gemm(x, w, out, M, N, K, bias=bias, gelu=True),
        "gelu": lambda: ops.gemm(x, w, out, M, N, K, bias=bias, gelu=True, aux_out=pre),
        # dgelu GEMM: dH[M, N] = (dY[M, K] W2[K(out), N]) * gelu'(pre): B operand MN-major
        "resid": lambda: ops.gemm(x, w, out, M, N, K, bias=bias, resid=pre),
        "dgelu": lambda: ops.gemm(dy, w2, dh, M, N, K, b_major=1, dgelu_aux=pre),
    }
    for k, fn in runs.items():
        if only and only != k: continue
        ms = t(fn)
        print(f"{name} M={M} N={N} K={K} {k:10s}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.0f} TF/s")
