"""Developer aid: throughput of the GPU fbank front end on AudioSet-shaped clips (10 s @ 16 kHz)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsiam_b200
B = 256
wav = torch.randn(B, 160000, device="cuda") * 0.1
for _ in range(3): avsiam_b200.wav2fbank(wav)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = avsiam_b200.wav2fbank(wav)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = (wav.numel() * 4 + out.numel() * 4) / 1e9
print(f"wav2fbank B={B} x 10 s clips: {ms:.3f} ms per batch = {B/ms*1e3:.0f} clips/s, {gb/ms*1e3:.0f} GB/s algorithmic")
