"""GPU audio front end (reference: AudiosetDataset._wav2fbank + normalisation, src/dataloader.py:268-345,505-506).

    fbank = avsiam_b200.wav2fbank(waveforms)        # [B, L] fp32 CUDA, 16 kHz mono  ->  [B, 1024, 128] fp32 CUDA

One launch for the whole batch instead of one torchaudio.compliance.kaldi.fbank call per clip on the loader's CPU
workers.  The mel filterbank (kaldi get_mel_banks: 128 triangular filters on the mel axis between 20 Hz and Nyquist,
512-point FFT) is computed here in float64 and uploaded once per device; everything else runs in csrc/fbank.cu.
The loader's random augmentations (mixup, SpecAugment masks, noise / roll, dataloader.py:491-516) stay host-side
decisions and are not part of this call.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib

SAMPLE_RATE, NFFT, NUM_MEL = 16000, 512, 128
_TABLES: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}


def mel_filterbank(num_bins: int = NUM_MEL, nfft: int = NFFT, sample_freq: float = SAMPLE_RATE, low: float = 20.0,
                   high: float = 0.0) -> np.ndarray:
    """kaldi get_mel_banks (no VTLN): [num_bins, nfft/2 + 1]; the Nyquist bin carries zero weight."""
    nyq = 0.5 * sample_freq
    if high <= 0.0:
        high += nyq
    mel = lambda f: 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)  # noqa: E731
    mlo, mhi = mel(low), mel(high)
    delta = (mhi - mlo) / (num_bins + 1)
    b = np.arange(num_bins, dtype=np.float64)[:, None]
    left, center, right = mlo + b * delta, mlo + (b + 1) * delta, mlo + (b + 2) * delta
    m = mel((sample_freq / nfft) * np.arange(nfft // 2, dtype=np.float64))[None, :]
    w = np.maximum(0.0, np.minimum((m - left) / (center - left), (right - m) / (right - center)))
    return np.concatenate([w, np.zeros((num_bins, 1))], axis=1)


def _tables(device: torch.device):
    key = str(device)
    if key not in _TABLES:
        w = mel_filterbank()
        rng = np.zeros((NUM_MEL, 2), dtype=np.int16)
        for i in range(NUM_MEL):
            nz = np.nonzero(w[i])[0]
            rng[i] = (nz[0], nz[-1]) if len(nz) else (1, 0)
        _TABLES[key] = (torch.from_numpy(w.astype(np.float32)).contiguous().to(device),
                        torch.from_numpy(rng).contiguous().to(device))
    return _TABLES[key]


def num_frames(n_samples: int) -> int:
    return 0 if n_samples < 400 else 1 + (n_samples - 400) // 160


def wav2fbank(waveform: torch.Tensor, target_length: int = 1024, norm_mean: float = -5.081, norm_std: float = 4.4849,
              remove_mean: bool = True) -> torch.Tensor:
    """waveform: [B, L] (or [L]) fp32 CUDA tensor at 16 kHz. Returns the normalised log-mel spectrogram
    [B, target_length, 128] the model consumes (`a_input`)."""
    if not waveform.is_cuda:
        raise RuntimeError("avsiam_b200.wav2fbank runs on CUDA only — there is no CPU path")
    squeeze = waveform.dim() == 1
    wav = (waveform.unsqueeze(0) if squeeze else waveform).contiguous().float()
    B, L = wav.shape
    melw, rng = _tables(wav.device)
    out = torch.empty(B, target_length, NUM_MEL, dtype=torch.float32, device=wav.device)
    scratch = torch.empty(max(B, 1), dtype=torch.float32, device=wav.device)
    rc = _lib.lib().avs_fbank(wav.data_ptr(), wav.stride(0), B, L, 1 if remove_mean else 0, melw.data_ptr(),
                              rng.data_ptr(), scratch.data_ptr(), out.data_ptr(), target_length, float(norm_mean),
                              float(norm_std), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "avs_fbank")
    return out[0] if squeeze else out
