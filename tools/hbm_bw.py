"""Developer aid: pure-write / pure-read / copy HBM bandwidth on this GPU (torch fill_, sum, copy_; CUDA events)."""
import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N = 1 << 30
a = torch.empty(N, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
w = t(lambda: a.fill_(1.0)); print(f"write-only  {2*N/w/1e6:8.0f} GB/s")
r = t(lambda: a.view(torch.int16).max()); print(f"read-only   {2*N/r/1e6:8.0f} GB/s")
c = t(lambda: b.copy_(a)); print(f"copy (r+w)  {4*N/c/1e6:8.0f} GB/s")
# elementwise 1-read-1-write at the LayerNorm call sizes of the step (bf16 [M, D])
for (M, D) in [(181248, 512), (45312, 768), (32768, 768)]:
    x = torch.randn(M, D, device="cuda").bfloat16(); y = torch.empty_like(x)
    ms = t(lambda: torch.mul(x, 2.0, out=y), n=20)
    print(f"elementwise bf16 [{M},{D}] ({x.numel()*2/1e6:.0f} MB in, same out): {ms*1e3:.1f} us = {x.numel()*4/ms/1e6:.0f} GB/s")
