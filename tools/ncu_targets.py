"""Developer aid: one launch (after one warm-up) of each kernel whose design DESIGN.md argues from an ncu capture —
the GELU(+derivative) and product(+column-sum) GEMM epilogues at the decoder MLP shape, the decoder attention forward /
backward, the encoder single-tile attention.

    ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attn_' -o gpurun_out/r02_targets \\
        python tools/ncu_targets.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import ops

M, N, K = 181248, 2048, 512
x = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
bias = torch.randn(N, device="cuda")
act = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
dact = torch.empty_like(act)
dy = torch.randn(M, K, device="cuda").bfloat16()
w2 = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
dh = torch.empty_like(act)
cs = torch.zeros(N, device="cuda")
for _ in range(2):
    ops.gemm(x, w, act, M, N, K, bias=bias, gelu=True, aux_out=dact, aux_grad=True)          # fc1 forward
    ops.gemm(dy, w2, dh, M, N, K, b_major=ops.MAJOR_MN, mul_aux=dact, colsum=cs)             # fc2 dgrad
del x, w, act, dact, dy, w2, dh
for (n_seq, S, H, hd) in [(256, 708, 16, 32), (256, 128, 12, 64), (256, 49, 12, 64)]:
    D = H * hd
    qkv = torch.randn(n_seq * S, 3 * D, device="cuda").bfloat16()
    out = torch.empty(n_seq * S, D, device="cuda", dtype=torch.bfloat16)
    dout = torch.randn_like(out)
    lse = torch.empty(n_seq, H, S, device="cuda")
    delta = torch.empty_like(lse)
    dqkv = torch.empty_like(qkv)
    dbias = torch.zeros(3 * D, device="cuda")
    for _ in range(2):
        ops.attention_fwd(qkv, out, lse, n_seq, S, H, hd)
        ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, n_seq, S, H, hd, dbias=dbias)
torch.cuda.synchronize()
print("ok")
