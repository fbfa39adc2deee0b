"""Developer aid: decoder attention (S = 708, head_dim 32) forward / backward alone, 20 iterations."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import ops
n_seq, S, H, hd = 256, 708, 16, 32
D = H * hd
qkv = torch.randn(n_seq * S, 3 * D, device="cuda").bfloat16()
out = torch.empty(n_seq * S, D, device="cuda", dtype=torch.bfloat16)
dout = torch.randn_like(out)
lse = torch.empty(n_seq, H, S, device="cuda"); delta = torch.empty_like(lse); dqkv = torch.empty_like(qkv)
dbias = torch.zeros(3 * D, device="cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
f = t(lambda: ops.attention_fwd(qkv, out, lse, n_seq, S, H, hd))
b = t(lambda: ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, n_seq, S, H, hd, dbias=dbias))
print(f"{os.environ.get('AVSIAM_B200_LIB', 'default')}: fwd {f:.3f} ms  bwd {b:.3f} ms")
