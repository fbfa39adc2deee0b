"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's audio front end.

The reference computes its input spectrogram with a third-party dependency that is not under /root/reference:
    torchaudio.compliance.kaldi.fbank(waveform, htk_compat=True, sample_frequency=sr, use_energy=False,
                                      window_type='hanning', num_mel_bins=128, dither=0.0, frame_shift=10)
(src/dataloader.py:323; requirements.txt pins torchaudio==2.0.2), preceded by `waveform - waveform.mean()` (:287)
and followed by pad/crop to target_length (:331-339) and `(fbank - norm_mean) / norm_std` (:506).
This file restates kaldi's published algorithm in numpy (float64 internally, so it is the more exact of the two);
tests/test_oracle_golden.py pins it to the outputs of torchaudio itself (tests/golden/fbank_kaldi.pt, written by
oracle/make_golden_fbank.py in the build container, where torchaudio 2.11 is installed).
"""
from __future__ import annotations

import math

import numpy as np

SAMPLE_RATE = 16000
FRAME_LEN = 400        # 25 ms
FRAME_SHIFT = 160      # 10 ms
NFFT = 512             # round_to_power_of_two
NUM_MEL = 128
LOW_FREQ, HIGH_FREQ = 20.0, 8000.0   # kaldi defaults: low_freq=20, high_freq=0 -> Nyquist
PREEMPH = 0.97
EPS = float(np.finfo(np.float32).eps)  # torchaudio: EPSILON = torch.finfo(torch.float).eps


def mel_scale(f):
    return 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_banks(num_bins: int = NUM_MEL, nfft: int = NFFT, sample_freq: float = SAMPLE_RATE, low: float = LOW_FREQ,
              high: float = HIGH_FREQ) -> np.ndarray:
    """kaldi get_mel_banks without VTLN: [num_bins, nfft/2 + 1] triangular weights on the mel axis (the last FFT bin
    gets zero weight: torchaudio pads one zero column)."""
    nbins_fft = nfft // 2
    width = sample_freq / nfft
    mlo, mhi = mel_scale(low), mel_scale(high)
    delta = (mhi - mlo) / (num_bins + 1)
    b = np.arange(num_bins, dtype=np.float64)[:, None]
    left, center, right = mlo + b * delta, mlo + (b + 1) * delta, mlo + (b + 2) * delta
    mel = mel_scale(width * np.arange(nbins_fft, dtype=np.float64))[None, :]
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    w = np.maximum(0.0, np.minimum(up, down))
    return np.concatenate([w, np.zeros((num_bins, 1))], axis=1)


def povey_free_hann(n: int = FRAME_LEN) -> np.ndarray:
    """window_type='hanning': torch.hann_window(n, periodic=False)."""
    return 0.5 - 0.5 * np.cos(2.0 * math.pi * np.arange(n, dtype=np.float64) / (n - 1))


def num_frames(n_samples: int) -> int:
    return 0 if n_samples < FRAME_LEN else 1 + (n_samples - FRAME_LEN) // FRAME_SHIFT   # snip_edges=True


def kaldi_fbank(wave: np.ndarray) -> np.ndarray:
    """wave: [L] float (already DC-removed by the caller like dataloader.py:287). Returns [num_frames, 128] float32."""
    wave = np.asarray(wave, dtype=np.float64)
    nf = num_frames(wave.shape[0])
    idx = np.arange(FRAME_LEN)[None, :] + FRAME_SHIFT * np.arange(nf)[:, None]
    fr = wave[idx]                                            # [nf, 400]
    fr = fr - fr.mean(axis=1, keepdims=True)                  # remove_dc_offset
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)    # replicate-pad on the left
    fr = fr - PREEMPH * prev                                  # pre-emphasis
    fr = fr * povey_free_hann()[None, :]
    spec = np.fft.rfft(fr, n=NFFT, axis=1)
    power = spec.real ** 2 + spec.imag ** 2                   # use_power=True
    mel = power @ mel_banks().T
    return np.log(np.maximum(mel, EPS)).astype(np.float32)    # use_log_fbank


def wav2fbank(wave: np.ndarray, target_length: int = 1024, norm_mean: float = -5.081, norm_std: float = 4.4849,
              remove_mean: bool = True) -> np.ndarray:
    """dataloader.py:287,323-339,506: waveform mean removal, fbank, zero-pad / crop to target_length, normalise.
    (Zero padding happens BEFORE normalisation, so padded frames become -norm_mean / norm_std.)"""
    wave = np.asarray(wave, dtype=np.float64)
    if remove_mean:
        wave = wave - wave.mean()
    fb = kaldi_fbank(wave)
    out = np.zeros((target_length, NUM_MEL), dtype=np.float32)
    n = min(target_length, fb.shape[0])
    out[:n] = fb[:n]
    return ((out - norm_mean) / norm_std).astype(np.float32)
