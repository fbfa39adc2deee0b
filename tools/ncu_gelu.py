"""Developer aid: one warm-up + one launch of the fc1 epilogue variants at the decoder MLP shape through a given build of
the library (raw C-ABI), for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:gemm --launch-skip 4 --launch-count 4 -o gpurun_out/x python tools/ncu_gelu.py LIB"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200._lib import GemmEpilogue, SIGNATURES, EPI_GELU

l = ctypes.CDLL(os.path.abspath(sys.argv[1]))
l.avs_gemm_bf16.argtypes = SIGNATURES["avs_gemm_bf16"]
l.avs_gemm_bf16.restype = ctypes.c_int
l.avs_last_error.restype = ctypes.c_char_p
M, N, K = 181248, 2048, 512
x = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
aux = torch.empty_like(out)


def gemm(gelu, with_aux):
    e = GemmEpilogue()
    e.flags = (EPI_GELU if gelu else 0) | (16 if with_aux else 0)
    e.alpha = 1.0
    e.bias = bias.data_ptr()
    if with_aux:
        e.aux_out, e.ld_aux = aux.data_ptr(), aux.stride(0)
    rc = l.avs_gemm_bf16(x.data_ptr(), x.stride(0), 0, w.data_ptr(), w.stride(0), 0, out.data_ptr(), out.stride(0), M, N, K,
                         ctypes.byref(e), 1, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(l.avs_last_error())


dy = torch.randn(M, K, device="cuda").bfloat16()
w2 = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
dh = torch.empty_like(out)
cs = torch.zeros(N, device="cuda")


def dgrad_mul():
    """fc2 dgrad x stored derivative + fused column sums (B operand MN-major)."""
    e = GemmEpilogue()
    e.flags = 32
    e.alpha = 1.0
    e.aux_in, e.ld_aux = aux.data_ptr(), aux.stride(0)
    e.colsum = cs.data_ptr()
    rc = l.avs_gemm_bf16(dy.data_ptr(), dy.stride(0), 0, w2.data_ptr(), w2.stride(0), 1, dh.data_ptr(), dh.stride(0), M, N, K,
                         ctypes.byref(e), 1, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(l.avs_last_error())


for _ in range(2):
    gemm(False, False)    # plain + bias
    gemm(True, False)     # GELU, one output
    gemm(True, True)      # GELU + stored derivative
    dgrad_mul()
torch.cuda.synchronize()
print("ok")
