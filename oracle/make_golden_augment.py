"""TEST INFRASTRUCTURE ONLY — golden vectors for the loader transforms, from the library calls the reference makes
(src/dataloader.py:152-155 torchvision Resize+Normalize; :491-516 torchaudio FrequencyMasking / TimeMasking, noise,
roll).                                                                  python -m oracle.make_golden_augment
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.make_golden import GOLDEN_DIR  # noqa: E402

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)           # timm IMAGENET_DEFAULT_MEAN / STD
FRAME_CASES = ((31, 2, 360, 640), (32, 3, 180, 320), (33, 1, 96, 128), (34, 2, 224, 224), (35, 1, 301, 227))
AUDIO_CASES = ((41, 96, 128, 48, 24, True), (42, 96, 128, 48, 0, False), (43, 64, 128, 0, 30, True),
               (44, 32, 128, 0, 0, True))
SUB = (slice(None), slice(None), slice(3, None, 7), slice(1, None, 5))   # stored sub-grid of the 224 x 224 output


def synth_frames(seed: int, n: int, h: int, w: int) -> np.ndarray:
    """Smooth gradients + texture + hard edges, uint8 [n, 3, h, w]."""
    g = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    out = np.zeros((n, 3, h, w))
    for i in range(n):
        for c in range(3):
            fx, fy, ph = g.uniform(1, 9), g.uniform(1, 9), g.uniform(0, 6.28)
            img = 0.5 + 0.3 * np.sin(6.28 * (fx * xx + fy * yy) + ph) + 0.15 * g.standard_normal((h, w))
            img[h // 3: h // 2, w // 4: w // 2] = g.uniform(0, 1)
            out[i, c] = img
    return np.clip(out * 255, 0, 255).astype(np.uint8)


def synth_fbank(seed: int, T: int, F: int) -> np.ndarray:
    return (np.random.default_rng(seed).standard_normal((T, F)) * 4.0 - 5.0).astype(np.float32)


def main():
    from PIL import Image
    from torchvision.transforms import Compose, Normalize, Resize
    import torchaudio
    out = {"frames": [], "audio": []}
    tf = Compose([Resize([224, 224], interpolation=Image.BICUBIC, antialias=True), Normalize(MEAN, STD)])
    for seed, n, h, w in FRAME_CASES:
        u8 = torch.from_numpy(synth_frames(seed, n, h, w))
        y = tf(u8 / 255)
        out["frames"].append({"seed": seed, "n": n, "h": h, "w": w, "sub": y[SUB].clone(),
                              "sum": float(y.double().sum()), "abs_sum": float(y.double().abs().sum())})
        print(f"[golden-augment] frames seed={seed} {h}x{w}: sum={float(y.sum()):.4f}")
    for seed, T, F, freqm, timem, noise in AUDIO_CASES:
        torch.manual_seed(seed)
        np.random.seed(seed)
        fbank = torch.from_numpy(synth_fbank(seed, T, F))
        x = fbank.transpose(0, 1).unsqueeze(0)
        if freqm != 0:
            x = torchaudio.transforms.FrequencyMasking(freqm)(x)
        if timem != 0:
            x = torchaudio.transforms.TimeMasking(timem)(x)
        x = x.squeeze(0).transpose(0, 1)
        x = (x - (-5.081)) / 4.4849
        if noise:
            x = x + torch.rand(x.shape[0], x.shape[1]) * np.random.rand() / 10
            x = torch.roll(x, np.random.randint(-T, T), 0)
        out["audio"].append({"seed": seed, "T": T, "F": F, "freqm": freqm, "timem": timem, "noise": noise,
                             "out": x.contiguous().clone()})
        print(f"[golden-augment] audio seed={seed}: zeros->{float((x == (0 + 5.081) / 4.4849).float().mean()):.3f}")
    torch.save(out, os.path.join(GOLDEN_DIR, "augment.pt"))


if __name__ == "__main__":
    main()
