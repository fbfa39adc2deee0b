"""TEST INFRASTRUCTURE ONLY — CPU restatement of the loader's per-sample transforms (second half).

  resize_normalize   `my_normalize(image / 255)` (src/dataloader.py:152-155,455-456): torchvision Resize([224,224],
                     BICUBIC, antialias=True) on a float tensor = ATen _upsample_bicubic2d_aa (separable; per output
                     pixel a window of `2*support+1` taps of the PIL cubic kernel a = -0.5 at 1/scale spacing, weights
                     normalised to sum 1; horizontal pass first), then Normalize(mean, std).  Restated here as two dense
                     weight matrices applied in float64.
  spec_augment_bands torchaudio FrequencyMasking / TimeMasking (mask_along_axis, p = 1): value = rand*param,
                     min = rand*(size - value), band = [floor(min), floor(min) + floor(value)) filled with 0.0
  augment_fbank      dataloader.py:491-516: bands -> (x - mean)/std -> + rand(T,F) * np.random.rand() / 10 -> roll
  mix_frames         dataloader.py:419-420

tests/test_oracle_golden.py pins this file to tests/golden/augment.pt — torchvision's / torchaudio's own outputs on the
calls the reference makes, produced by oracle/make_golden_augment.py in the build container (torchvision 0.26,
torchaudio 2.11).  Product code never imports this module.
"""
from __future__ import annotations

import math

import numpy as np


def _cubic(x: np.ndarray, a: float = -0.5) -> np.ndarray:
    x = np.abs(x)
    return np.where(x < 1, ((a + 2) * x - (a + 3)) * x * x + 1,
                    np.where(x < 2, ((a * x - 5 * a) * x + 8 * a) * x - 4 * a, 0.0))


def aa_matrix(in_size: int, out_size: int) -> np.ndarray:
    """Dense [out, in] matrix of the antialiased bicubic resampling along one axis."""
    scale = np.float32(in_size) / np.float32(out_size)              # ATen computes the geometry in float32
    support = np.float32(2.0) * scale if scale >= 1 else np.float32(2.0)
    inv = np.float32(1.0) / scale if scale >= 1 else np.float32(1.0)
    cap = int(math.ceil(float(support))) * 2 + 1
    M = np.zeros((out_size, in_size))
    for i in range(out_size):
        center = np.float32(scale * np.float32(i + 0.5))
        lo = max(int(np.float32(center - support + np.float32(0.5))), 0)
        hi = min(int(np.float32(center + support + np.float32(0.5))), in_size)
        n = min(max(hi - lo, 0), cap)
        j = np.arange(lo, lo + n)
        w = _cubic(((j - center + np.float32(0.5)) * inv).astype(np.float32).astype(np.float64))
        M[i, j] = w / w.sum() if w.sum() != 0 else w
    return M


def resize_normalize(frames_u8: np.ndarray, size: int, mean, std) -> np.ndarray:
    """uint8 [N, C, H, W] -> float64 [N, C, size, size]."""
    x = frames_u8.astype(np.float64) / 255.0
    Wx = aa_matrix(frames_u8.shape[3], size)
    Wy = aa_matrix(frames_u8.shape[2], size)
    y = Wy @ (x @ Wx.T)                                   # horizontal pass, then vertical
    m = np.asarray(mean, np.float64).reshape(1, -1, 1, 1)
    s = np.asarray(std, np.float64).reshape(1, -1, 1, 1)
    return (y - m) / s


def mix_frames(a: np.ndarray, b: np.ndarray, w: np.ndarray) -> np.ndarray:
    w = np.asarray(w, np.float32).reshape(-1, 1, 1, 1)
    return (w * a.astype(np.float32) + (np.float32(1) - w) * b.astype(np.float32)).astype(np.float32)


def band(u_value: float, u_min: float, mask_param: int, size: int):
    """[start, end) from the two uniform draws, float32 arithmetic like torch.rand(1) * param."""
    if mask_param < 1:
        return 0, 0
    value = np.float32(u_value) * np.float32(mask_param)
    mn = np.float32(u_min) * (np.float32(size) - value)
    return int(mn), int(mn) + int(value)


def augment_fbank(x: np.ndarray, f_band, t_band, mean: float, std: float, noise=None, r: float = 0.0, shift: int = 0,
                  skip_norm: bool = False) -> np.ndarray:
    """x float32 [T, F] -> float32 [T, F], every step in float32 like the torch CPU ops."""
    y = x.astype(np.float32).copy()
    y[:, f_band[0]:f_band[1]] = 0.0
    y[t_band[0]:t_band[1], :] = 0.0
    if not skip_norm:
        y = ((y - np.float32(mean)) / np.float32(std)).astype(np.float32)
    if noise is not None:
        y = (y + (noise.astype(np.float32) * np.float32(r)) / np.float32(10)).astype(np.float32)
        y = np.roll(y, shift, axis=0)
    return y
