"""GPU input pipeline, second half (reference: AudiosetDataset.__getitem__, src/dataloader.py:148-155,371-516).

What the loader's CPU workers do to ONE decoded sample, done here for a whole batch on the device:

    frames = preprocess_frames(u8)            # uint8 [N,3,H,W] -> /255 -> Resize([224,224], BICUBIC, antialias) ->
                                              # Normalize(IMAGENET mean/std)              (dataloader.py:152-155,455-456)
    frames = mix_frames(frames, frames2, w)   # weight*image + (1-weight)*image2          (:419-420)
    draws  = draw_augment_params(B, ...)      # the loader's random draws, in its order   (:491-516)
    fbank  = augment_fbank(fbank, draws, ...) # SpecAugment masks, normalise, noise, roll (:491-516)

Video DEMUX / DECODE (torchvision.io.VideoReader, :452-454) is not here: it needs NVDEC bindings that this image does
not ship; frames enter as uint8 tensors.  Random draws are made on the host with the same generators and in the same
order as the reference (torch.rand for the SpecAugment bands, numpy for the noise scale and the roll), and are passed
to the kernels as data, so a run is reproducible and testable against torchaudio / torchvision bit for bit (audio) or
to fp32 rounding (bicubic).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib

F32 = torch.float32
IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)     # timm.data.constants, used at dataloader.py:154
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


# ------------------------------------------------------------------------------------------------ frames
def _aa_cubic(x: np.ndarray) -> np.ndarray:
    """PIL / ATen antialias bicubic kernel, a = -0.5, evaluated in float32 like ATen's opmath for float tensors."""
    a = np.float32(-0.5)
    x = np.abs(x).astype(np.float32)
    near = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + np.float32(1)
    far = ((a * x - np.float32(5) * a) * x + np.float32(8) * a) * x - np.float32(4) * a
    return np.where(x < 1, near, np.where(x < 2, far, np.float32(0))).astype(np.float32)


def aa_resize_weights(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per output index: first input tap, number of taps, normalised weights [out, max_taps] (zero padded) of
    torch.nn.functional.interpolate(mode='bicubic', antialias=True, align_corners=False) along one axis."""
    f = np.float32
    scale = f(in_size) / f(out_size)
    support = f(2.0) * scale if scale >= 1.0 else f(2.0)
    invscale = f(1.0) / scale if scale >= 1.0 else f(1.0)
    max_taps = int(math.ceil(float(support))) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    xsize = np.zeros(out_size, np.int32)
    w = np.zeros((out_size, max_taps), np.float32)
    for i in range(out_size):
        center = scale * f(i + 0.5)
        lo = max(int(f(center - support + f(0.5))), 0)
        n = min(int(f(center + support + f(0.5))), in_size) - lo
        n = min(max(n, 0), max_taps)
        j = np.arange(n, dtype=np.float32)
        wj = _aa_cubic((j + f(lo) - center + f(0.5)) * invscale)
        total = f(0)
        for v in wj:                                    # sequential float32 sum, as ATen accumulates it
            total = f(total + v)
        if total != 0:
            wj = wj * (f(1.0) / total)
        xmin[i], xsize[i] = lo, n
        w[i, :n] = wj
    return xmin, xsize, w


_RESIZE_TABLES: Dict[Tuple[str, int, int], tuple] = {}
_NORM_TABLES: Dict[Tuple[str, tuple, tuple], Tuple[torch.Tensor, torch.Tensor]] = {}
_SMEM_BUDGET = 160 * 1024                      # bytes of shared memory a resize CTA may use


def _resize_tables(device, in_size: int, out_size: int):
    key = (str(device), in_size, out_size)
    if key not in _RESIZE_TABLES:
        xmin, xsize, w = aa_resize_weights(in_size, out_size)
        _RESIZE_TABLES[key] = (torch.from_numpy(w).contiguous().to(device), torch.from_numpy(xmin).to(device),
                               torch.from_numpy(xsize).to(device), w.shape[1], xmin, xsize)
    return _RESIZE_TABLES[key]


def _tile_plan(ymin: np.ndarray, ysize: np.ndarray, out_w: int, taps_x: int, W: int) -> Tuple[int, int]:
    """Output rows per CTA and the widest input-row span any tile needs, so that the CTA's shared memory
    (table + transposed column weights + the fp32 tile + 4 staged input rows) stays within the budget."""
    oh = len(ymin)
    for ty in (16, 8, 4, 2, 1):
        span = max(int(ymin[min(y0 + ty, oh) - 1] + ysize[min(y0 + ty, oh) - 1] - ymin[y0]) for y0 in range(0, oh, ty))
        if 4 * (256 + taps_x * out_w + span * out_w + 4 * W) <= _SMEM_BUDGET:
            return ty, span
    raise ValueError("preprocess_frames: frame too large for the resize tile (shared memory)")


def preprocess_frames(frames: torch.Tensor, size: int = 224, mean=IMAGENET_DEFAULT_MEAN,
                      std=IMAGENET_DEFAULT_STD) -> torch.Tensor:
    """frames: uint8 CUDA tensor [N, C, H, W] (decoded video frames) -> fp32 [N, C, size, size], i.e.
    `my_normalize(frames / 255)` of dataloader.py:152-155,455-456."""
    if not frames.is_cuda:
        raise RuntimeError("avsiam_b200.augment.preprocess_frames runs on CUDA only — there is no CPU path")
    if frames.dtype != torch.uint8 or frames.dim() != 4:
        raise ValueError("preprocess_frames: expected a uint8 tensor [N, C, H, W]")
    frames = frames.contiguous()
    N, C, H, W = frames.shape
    if len(mean) != C or len(std) != C:
        raise ValueError("preprocess_frames: mean / std need one entry per channel")
    dev = frames.device
    wx, xmin, xsize, tx, _, _ = _resize_tables(dev, W, size)
    wy, ymin, ysize, ty, ymin_h, ysize_h = _resize_tables(dev, H, size)
    tile_rows, span = _tile_plan(ymin_h, ysize_h, size, tx, W)
    nkey = (str(dev), tuple(mean), tuple(std))
    if nkey not in _NORM_TABLES:
        _NORM_TABLES[nkey] = (torch.tensor(mean, dtype=F32, device=dev), torch.tensor(std, dtype=F32, device=dev))
    m, s = _NORM_TABLES[nkey]
    out = torch.empty(N, C, size, size, dtype=F32, device=dev)
    step = max(1, 65535 // C)                                  # planes per launch (grid.y limit)
    for n0 in range(0, N, step):
        n1 = min(N, n0 + step)
        rc = _lib.lib().avs_frames_preprocess(frames[n0:n1].data_ptr(), n1 - n0, C, H, W, wx.data_ptr(),
                                              xmin.data_ptr(), xsize.data_ptr(), tx, wy.data_ptr(), ymin.data_ptr(),
                                              ysize.data_ptr(), ty, size, size, tile_rows, span, m.data_ptr(),
                                              s.data_ptr(), out[n0:n1].data_ptr(),
                                              torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "avs_frames_preprocess")
    return out


def mix_frames(image: torch.Tensor, image2: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """In place: image[n] = weight[n] * image[n] + (1 - weight[n]) * image2[n]   (dataloader.py:419-420)."""
    if not (image.is_cuda and image2.is_cuda):
        raise RuntimeError("avsiam_b200.augment.mix_frames runs on CUDA only — there is no CPU path")
    assert image.shape == image2.shape and image.dtype == F32 and image2.dtype == F32 and image.is_contiguous()
    N = image.shape[0]
    weight = weight.to(device=image.device, dtype=F32).contiguous()
    assert weight.numel() == N
    rc = _lib.lib().avs_mix_frames(image.data_ptr(), image2.contiguous().data_ptr(), weight.data_ptr(), N,
                                   image[0].numel() if N else 1, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "avs_mix_frames")
    return image


# ------------------------------------------------------------------------------------------------ audio
@dataclass
class AugmentDraws:
    """Per-sample random draws of dataloader.py:491-516.  params int32 [B, 6] = f0, f1, t0, t1, shift, 0 (bands are
    [start, end)); scale fp32 [B] = the loader's np.random.rand() (the kernel applies `noise * scale / 10`); noise fp32
    [B, T, F] uniform [0, 1) or None."""
    params: torch.Tensor
    scale: Optional[torch.Tensor]
    noise: Optional[torch.Tensor]


def _band(mask_param: int, size: int, generator) -> Tuple[int, int]:
    """torchaudio.functional.mask_along_axis (p = 1): two torch.rand(1) draws -> [start, end)."""
    if mask_param < 1:
        return 0, 0
    value = torch.rand(1, generator=generator) * mask_param
    min_value = torch.rand(1, generator=generator) * (size - value)
    start = int(min_value.long())
    return start, start + int(value.long())


def draw_augment_params(B: int, T: int = 1024, F: int = 128, freqm: int = 0, timem: int = 0, noise: bool = False,
                        generator: Optional[torch.Generator] = None, np_rng=None, device="cuda",
                        host_noise: bool = False) -> AugmentDraws:
    """Draws, sample after sample in the loader's order: frequency band (2 x torch.rand), time band (2 x torch.rand),
    then — with `noise` — the uniform noise field torch.rand(T, F), the noise scale np.random.rand() and the roll
    np.random.randint(-T, T).  The noise field is one device-side torch.rand for the batch; `host_noise=True` draws it
    per sample from `generator` on the host exactly where the loader does (slow — for reproducing a loader run)."""
    np_rng = np_rng if np_rng is not None else np.random
    p = np.zeros((B, 6), np.int32)
    scale = np.zeros(B, np.float32)
    fields = []
    for b in range(B):
        p[b, 0], p[b, 1] = _band(freqm, F, generator)
        p[b, 2], p[b, 3] = _band(timem, T, generator)
        if noise:
            if host_noise:
                fields.append(torch.rand(T, F, generator=generator))
            scale[b] = np.float32(np_rng.rand())
            p[b, 4] = np_rng.randint(-T, T)
    params = torch.from_numpy(p).to(device)
    if not noise:
        return AugmentDraws(params, None, None)
    field = torch.stack(fields).to(device) if host_noise else torch.rand(B, T, F, dtype=F32, device=device)
    return AugmentDraws(params, torch.from_numpy(scale).to(device), field)


def augment_fbank(fbank: torch.Tensor, draws: AugmentDraws, norm_mean: float = -5.081, norm_std: float = 4.4849,
                  skip_norm: bool = False) -> torch.Tensor:
    """fbank fp32 CUDA [B, T, F] UN-normalised log-mel (wav2fbank(..., normalise off) or the loader's tensor) ->
    masked, normalised, noised, rolled copy (dataloader.py:491-516)."""
    if not fbank.is_cuda:
        raise RuntimeError("avsiam_b200.augment.augment_fbank runs on CUDA only — there is no CPU path")
    fbank = fbank.contiguous().to(F32)
    B, T, F = fbank.shape
    assert draws.params.shape == (B, 6) and draws.params.dtype == torch.int32
    out = torch.empty_like(fbank)
    noise = draws.noise.contiguous() if draws.noise is not None else None
    if noise is not None:
        assert noise.shape == (B, T, F) and draws.scale is not None
    rc = _lib.lib().avs_fbank_augment(fbank.data_ptr(), draws.params.contiguous().data_ptr(),
                                      draws.scale.data_ptr() if noise is not None else None,
                                      noise.data_ptr() if noise is not None else None, out.data_ptr(), B, T, F,
                                      float(norm_mean), float(norm_std), 1 if skip_norm else 0,
                                      torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "avs_fbank_augment")
    return out
