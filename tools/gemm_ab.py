"""Developer aid: A/B of two builds of libavsiam_b200.so on the same box (ctypes, raw C-ABI): the fused-epilogue GEMM shapes.
    python tools/gemm_ab.py tools/ab/libavsiam_b200_r1.so avsiam_b200/libavsiam_b200.so"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200._lib import GemmEpilogue, SIGNATURES, EPI_GELU, EPI_DGELU


def load(path):
    l = ctypes.CDLL(os.path.abspath(path))
    l.avs_gemm_bf16.argtypes = SIGNATURES["avs_gemm_bf16"]
    l.avs_gemm_bf16.restype = ctypes.c_int
    l.avs_last_error.restype = ctypes.c_char_p
    return l


def gemm(l, a, b, c, M, N, K, b_major=0, bias=None, gelu=False, aux_out=None, resid=None, dgelu_aux=None, colsum=None,
         aux_grad=False, mul_aux=None):
    e = GemmEpilogue()
    e.flags = (EPI_GELU if gelu else 0) | (EPI_DGELU if dgelu_aux is not None else 0) | (16 if aux_grad else 0) | \
              (32 if mul_aux is not None else 0)
    if mul_aux is not None:
        dgelu_aux = mul_aux
    e.alpha = 1.0
    e.bias = bias.data_ptr() if bias is not None else None
    e.resid = resid.data_ptr() if resid is not None else None
    e.ld_resid = resid.stride(0) if resid is not None else 0
    aux = aux_out if aux_out is not None else dgelu_aux
    e.aux_in = dgelu_aux.data_ptr() if dgelu_aux is not None else None
    e.aux_out = aux_out.data_ptr() if aux_out is not None else None
    e.ld_aux = aux.stride(0) if aux is not None else 0
    e.colsum = colsum.data_ptr() if colsum is not None else None
    rc = l.avs_gemm_bf16(a.data_ptr(), a.stride(0), 0, b.data_ptr(), b.stride(0), b_major, c.data_ptr(), c.stride(0), M, N,
                         K, ctypes.byref(e), 1, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(l.avs_last_error())


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


def wgrad(l, dy, x, dw, n_out, k_in, M):
    """dW[n_out, k_in] += dY[M, n_out]^T X[M, k_in]: both operands MN-major, fp32 red.global.add, automatic split-K."""
    e = GemmEpilogue()
    e.flags = 8
    e.alpha = 1.0
    rc = l.avs_gemm_bf16(dy.data_ptr(), dy.stride(0), 1, x.data_ptr(), x.stride(0), 1, dw.data_ptr(), dw.stride(0), n_out,
                         k_in, M, ctypes.byref(e), 0, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(l.avs_last_error())


libs = [(p, load(p)) for p in sys.argv[1:]]
print("columns:", [p for p, _ in libs])
if os.environ.get("AB_WGRAD", "1") != "0":
    for (tokens, n_out, k_in, name) in [(181248, 512, 2048, "dec fc2 wgrad"), (181248, 2048, 512, "dec fc1 wgrad"),
                                         (181248, 1536, 512, "dec qkv wgrad"), (181248, 512, 512, "dec proj wgrad"),
                                         (45312, 768, 3072, "enc fc2 wgrad"), (45312, 3072, 768, "enc fc1 wgrad"),
                                         (45312, 2304, 768, "enc qkv wgrad"), (45312, 768, 768, "enc proj wgrad")]:
        dy = torch.randn(tokens, n_out, device="cuda").bfloat16()
        x = torch.randn(tokens, k_in, device="cuda").bfloat16()
        dw = torch.zeros(n_out, k_in, device="cuda")
        cells = []
        for rep in range(2):
            for p, l in libs:
                ms = t(lambda: wgrad(l, dy, x, dw, n_out, k_in, tokens))
                cells.append(f"{ms:.3f} ({2.0 * tokens * n_out * k_in / ms / 1e9:.0f})")
        print(f"{name} M={n_out} N={k_in} K={tokens}: " + " | ".join(cells), flush=True)
        del dy, x, dw
if os.environ.get("AB_WGRAD_ONLY"):
    sys.exit(0)
for (M, N, K, name) in [(181248, 2048, 512, "dec"), (45312, 3072, 768, "enc")]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.randn(M, N, device="cuda").bfloat16()
    dy = torch.randn(M, K, device="cuda").bfloat16()
    w2 = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    dh = torch.empty_like(out)
    cs = torch.zeros(N, device="cuda")
    wf2 = (torch.randn(K, N, device="cuda") * 0.02).bfloat16()
    res = torch.randn(M, K, device="cuda").bfloat16()
    o2 = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    b2 = torch.randn(K, device="cuda")
    runs = {
        "fc1 plain+bias": lambda l: gemm(l, x, w, out, M, N, K, bias=bias),
        "fc1 gelu+aux": lambda l: gemm(l, x, w, out, M, N, K, bias=bias, gelu=True, aux_out=dh),
        "fc1 resid": lambda l: gemm(l, x, w, out, M, N, K, bias=bias, resid=pre),
        "fc2 dgrad dgelu": lambda l: gemm(l, dy, w2, dh, M, N, K, b_major=1, dgelu_aux=pre),
        "fc2 dgrad dgelu+colsum": lambda l: gemm(l, dy, w2, dh, M, N, K, b_major=1, dgelu_aux=pre, colsum=cs),
        "fc1 gelu+GRADaux (new)": lambda l: gemm(l, x, w, out, M, N, K, bias=bias, gelu=True, aux_out=dh, aux_grad=True),
        "fc2 dgrad MULaux (new)": lambda l: gemm(l, dy, w2, dh, M, N, K, b_major=1, mul_aux=pre),
        "fc2 dgrad MULaux+colsum": lambda l: gemm(l, dy, w2, dh, M, N, K, b_major=1, mul_aux=pre, colsum=cs),
        "fc2 fwd bias+resid": lambda l: gemm(l, pre, wf2, o2, M, K, N, bias=b2, resid=res),
        "fc2 fwd bias": lambda l: gemm(l, pre, wf2, o2, M, K, N, bias=b2),
    }
    for k, fn in runs.items():
        cells = []
        for rep in range(2):
            for p, l in libs:
                try:
                    ms = t(lambda: fn(l))
                    cells.append(f"{ms:.3f}")
                except RuntimeError:
                    cells.append("  n/a")
        print(f"{name} M={M} N={N} K={K} {k:24s}: " + " | ".join(cells) + "   (lib order repeated twice)", flush=True)
