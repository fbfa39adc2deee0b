"""Developer aid: throughput of the GPU loader transforms (frames resize+normalise, fbank augmentation) and of the
retrieval similarity / rank kernels at AudioSet / VGGSound evaluation sizes."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import augment as A, evaluate as E


def timeit(fn, n=50):
    for _ in range(20): fn()                    # long enough for the clocks to leave the idle state
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


N, H, W = 256, 360, 640
u8 = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, device="cuda")
ms, out = timeit(lambda: A.preprocess_frames(u8))
alg = u8.numel() + out.numel() * 4
print(f"preprocess_frames {N} x 3x{H}x{W} u8 -> 224x224: {ms:.3f} ms = {N/ms*1e3:.0f} frames/s, "
      f"{alg/ms*1e-6:.0f} GB/s algorithmic (uint8 in + fp32 out)")
B = 256
x = torch.randn(B, 1024, 128, device="cuda")
d = A.draw_augment_params(B, freqm=48, timem=192, noise=True)
ms, out = timeit(lambda: A.augment_fbank(x, d))
print(f"augment_fbank B={B}: {ms:.3f} ms, {3 * x.numel() * 4 / ms * 1e-6:.0f} GB/s (read fbank + noise, write out)")
for n in (1545, 2635, 15446):
    a, v = torch.randn(n, 768, device="cuda"), torch.randn(n, 768, device="cuda")
    ms, sim = timeit(lambda: E.get_sim_mat(a, v), 5)
    ms2, _ = timeit(lambda: E.compute_metrics(sim), 5)
    print(f"retrieval n={n}: get_sim_mat {ms:.3f} ms ({2 * n * n * 768 / ms * 1e-9:.1f} TF/s fp32), "
          f"compute_metrics {ms2:.3f} ms incl. the host read")
