"""Developer aid: time the attention kernels alone on the bench shapes (CUDA events, 10 iterations)."""
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

for (n_seq, S, H, hd) in [(256, 708, 16, 32), (256, 128, 12, 64), (256, 49, 12, 64), (256, 177, 12, 64)]:
    D = H * hd
    qkv = torch.randn(n_seq * S, 3 * D, device="cuda").bfloat16()
    out = torch.empty(n_seq * S, D, device="cuda", dtype=torch.bfloat16)
    dout = torch.randn_like(out)
    lse = torch.empty(n_seq, H, S, device="cuda")
    delta = torch.empty_like(lse)
    dqkv = torch.empty_like(qkv)
    f = t(lambda: ops.attention_fwd(qkv, out, lse, n_seq, S, H, hd))
    b = t(lambda: ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, n_seq, S, H, hd))
    fl = 4.0 * n_seq * H * S * S * hd
    print(f"S={S} H={H} hd={hd}: fwd {f:.3f} ms ({fl/f/1e9:.0f} TF/s)  bwd {b:.3f} ms ({2.5*fl/b/1e9:.0f} TF/s)")
