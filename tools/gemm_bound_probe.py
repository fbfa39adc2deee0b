"""Developer aid: what bounds the epilogue-heavy GEMM shapes?  Needs a -DAVS_GEMM_DEBUG build of the library
(tools/ab/libavsiam_b200_dbg.so): times the fc1 GELU (+ stored derivative) and fc2-dgrad (x stored derivative) shapes
with the TMA stores and / or the in-stream loads switched off (results are wrong by design in those modes).
    python tools/gemm_bound_probe.py tools/ab/libavsiam_b200_dbg.so"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200._lib import GemmEpilogue, SIGNATURES, EPI_GELU


def load(path):
    l = ctypes.CDLL(os.path.abspath(path))
    l.avs_gemm_bf16.argtypes = SIGNATURES["avs_gemm_bf16"]
    l.avs_gemm_bf16.restype = ctypes.c_int
    l.avs_last_error.restype = ctypes.c_char_p
    return l


def gemm(l, a, b, c, M, N, K, b_major=0, bias=None, gelu=False, aux_out=None, aux_grad=False, mul_aux=None, colsum=None,
         resid=None):
    e = GemmEpilogue()
    e.flags = (EPI_GELU if gelu else 0) | (16 if aux_grad else 0) | (32 if mul_aux is not None else 0)
    e.alpha = 1.0
    e.bias = bias.data_ptr() if bias is not None else None
    e.resid = resid.data_ptr() if resid is not None else None
    e.ld_resid = resid.stride(0) if resid is not None else 0
    aux = aux_out if aux_out is not None else mul_aux
    e.aux_in = mul_aux.data_ptr() if mul_aux is not None else None
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


l = load(sys.argv[1])
for (M, N, K, name) in [(181248, 2048, 512, "dec"), (45312, 3072, 768, "enc")]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    aux = torch.randn(M, N, device="cuda").bfloat16()
    dy = torch.randn(M, K, device="cuda").bfloat16()
    w2 = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    dh = torch.empty_like(out)
    cs = torch.zeros(N, device="cuda")
    runs = {
        "fc1 plain+bias": lambda: gemm(l, x, w, out, M, N, K, bias=bias),
        "fc1 gelu (one output)": lambda: gemm(l, x, w, out, M, N, K, bias=bias, gelu=True),
        "fc1 gelu+GRADaux": lambda: gemm(l, x, w, out, M, N, K, bias=bias, gelu=True, aux_out=dh, aux_grad=True),
        "fc2 dgrad MULaux": lambda: gemm(l, dy, w2, dh, M, N, K, b_major=1, mul_aux=aux),
        "fc2 dgrad MULaux+colsum": lambda: gemm(l, dy, w2, dh, M, N, K, b_major=1, mul_aux=aux, colsum=cs),
        "fc2 dgrad plain": lambda: gemm(l, dy, w2, dh, M, N, K, b_major=1),
    }
    for k, fn in runs.items():
        cells = []
        for variant in (0, 2, 4, 6, 0):
            l.avs_debug_set_desc_variant(variant)
            cells.append(f"{t(fn):.3f}")
        l.avs_debug_set_desc_variant(0)
        fl = 2.0 * M * N * K
        print(f"{name} M={M} N={N} K={K} {k:24s}: normal {cells[0]} | no-store {cells[1]} | no-in {cells[2]} | neither "
              f"{cells[3]} | normal {cells[4]}   ({fl / float(cells[0]) / 1e9:.0f} TF/s normal)", flush=True)
