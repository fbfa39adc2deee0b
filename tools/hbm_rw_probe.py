"""Developer aid: HBM bandwidth by direction on this box — write-only (fill), read-only (sum), copy — with torch's own
kernels over 4 GiB buffers (far beyond the 126 MB L2), CUDA events, best of 5."""
import torch

n = 1 << 31   # bf16 elements = 4 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")


def t(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


gb = n * 2 / 1e9
ms = t(lambda: a.fill_(1.0)); print(f"write-only fill : {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
ms = t(lambda: a.zero_()); print(f"write-only zero : {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
ai = a.view(torch.int32)
ms = t(lambda: ai.sum()); print(f"read-only sum   : {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy (r+w bytes): {ms:.3f} ms  {2 * gb / ms * 1e3:.0f} GB/s")
