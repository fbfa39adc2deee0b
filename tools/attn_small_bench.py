"""Developer aid: the persistent single-tile attention kernels alone (encoder shapes), CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10


def t(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (n_seq, S, H, hd) in [(256, 128, 12, 64), (256, 49, 12, 64)]:
    D = H * hd
    qkv = torch.randn(n_seq * S, 3 * D, device="cuda").bfloat16()
    out = torch.empty(n_seq * S, D, device="cuda", dtype=torch.bfloat16)
    dout = torch.randn_like(out)
    lse = torch.empty(n_seq, H, S, device="cuda")
    delta = torch.empty_like(lse)
    dqkv = torch.empty_like(qkv)
    dbias = torch.zeros(3 * D, device="cuda")
    f = t(lambda: ops.attention_fwd(qkv, out, lse, n_seq, S, H, hd), iters)
    b = t(lambda: ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, n_seq, S, H, hd, dbias=dbias), iters)
    fl = 4.0 * n_seq * H * S * S * hd
    byt_f, byt_b = 4.0 * n_seq * S * D * 2, 8.0 * n_seq * S * D * 2
    print(f"S={S} H={H} hd={hd}: fwd {f*1e3:.1f} us ({fl/f/1e9:.0f} TF/s, {byt_f/f/1e9:.2f} TB/s)  "
          f"bwd {b*1e3:.1f} us ({2.5*fl/b/1e9:.0f} TF/s, {byt_b/b/1e9:.2f} TB/s)")
