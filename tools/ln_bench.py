"""Developer aid: LayerNorm forward / backward bandwidth at the call sizes of the step."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import ops
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, D) in [(181248, 512), (45312, 768), (32768, 768), (12544, 768)]:
    # rotate over several buffers so the data does not sit in the 126 MB L2
    nb = max(2, int(400e6 // (M * D * 2)) + 1)
    xs = [torch.randn(M, D, device="cuda").bfloat16() for _ in range(nb)]
    y = torch.empty_like(xs[0]); dy = torch.randn_like(xs[0]); res = torch.randn_like(xs[0]); dx = torch.empty_like(xs[0])
    g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    dg, db, dbias = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    i = [0]
    def fwd():
        i[0] += 1
        ops.layernorm_fwd(xs[i[0] % nb], g, b, 1e-5, y, mean, rstd, M, D)
    def bwd():
        i[0] += 1
        ops.layernorm_bwd(dy, xs[i[0] % nb], mean, rstd, g, dx, dg, db, M, D, resid=res, dbias=dbias)
    f = t(fwd); bw = t(bwd)
    print(f"[{M},{D}] fwd {f*1e3:6.1f} us = {M*D*4/f/1e6:5.0f} GB/s   bwd {bw*1e3:6.1f} us = {M*D*8/bw/1e6:5.0f} GB/s")
