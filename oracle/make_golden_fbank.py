"""TEST INFRASTRUCTURE ONLY — golden vectors for the audio front end, produced by the dependency the reference
calls (torchaudio.compliance.kaldi.fbank with the arguments of src/dataloader.py:323).

    python -m oracle.make_golden_fbank       (needs torchaudio; the build container has 2.11)
"""
import os
import sys

import torch
import torchaudio

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.make_golden import GOLDEN_DIR  # noqa: E402


def synth_wave(seed: int, n: int) -> torch.Tensor:
    """Deterministic test signal: a few chirps + noise bursts with a wide dynamic range, in [-1, 1]."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n, dtype=torch.float64) / 16000.0
    w = 0.3 * torch.sin(2 * torch.pi * (200.0 + 900.0 * t) * t) + 0.1 * torch.sin(2 * torch.pi * 3100.0 * t)
    w = w + 0.05 * torch.randn(n, generator=g, dtype=torch.float64) * (torch.sin(2 * torch.pi * 0.7 * t) > 0)
    w[n // 2: n // 2 + 4000] = 0.0                       # a stretch of digital silence (log floor)
    w = w + 0.02                                         # DC offset
    return w.clamp(-1, 1).float()


def main():
    cases = []
    for seed, n in ((2, 48017), (3, 400), (5, 33123)):
        w = synth_wave(seed, n)
        wave = (w - w.mean()).unsqueeze(0)               # dataloader.py:287
        fb = torchaudio.compliance.kaldi.fbank(wave, htk_compat=True, sample_frequency=16000, use_energy=False,
                                               window_type='hanning', num_mel_bins=128, dither=0.0, frame_shift=10)
        cases.append({"seed": seed, "n": n, "fbank": fb.clone()})
        print(f"[golden-fbank] seed={seed} n={n}: frames={fb.shape[0]} mean={float(fb.mean()):.4f} min={float(fb.min()):.3f}")
    torch.save(cases, os.path.join(GOLDEN_DIR, "fbank_kaldi.pt"))


if __name__ == "__main__":
    main()
