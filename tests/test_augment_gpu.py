"""Loader transforms on the GPU against golden vectors from the library calls the reference makes
(tests/golden/augment.pt: torchvision Resize(BICUBIC, antialias)+Normalize, torchaudio Frequency/TimeMasking + noise +
roll — src/dataloader.py:152-155,491-516) and against the CPU oracle.
Audio: bit-exact given the same draws.  Frames: 2e-5 absolute (fp32 separable filter, ~13 taps per axis)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from avsiam_b200 import augment as A  # noqa: E402
from oracle import augment_oracle as AO  # noqa: E402
from oracle.make_golden_augment import MEAN, STD, SUB, synth_fbank, synth_frames  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "augment.pt"), weights_only=False)


def test_frames_match_torchvision_golden(golden):
    for c in golden["frames"]:
        u8 = torch.from_numpy(synth_frames(c["seed"], c["n"], c["h"], c["w"])).to(DEV)
        y = A.preprocess_frames(u8, 224, MEAN, STD)
        assert y.shape == (c["n"], 3, 224, 224) and y.dtype == torch.float32
        assert (y[SUB].cpu() - c["sub"]).abs().max() < 2e-5
        assert abs(float(y.double().sum()) - c["sum"]) < 1e-6 * c["abs_sum"]


def test_frames_batch_against_oracle_and_mixup():
    u8 = synth_frames(7, 10, 240, 426)                               # the loader's 10 frames of one clip
    y = A.preprocess_frames(torch.from_numpy(u8).to(DEV))
    ref = AO.resize_normalize(u8, 224, MEAN, STD)
    assert np.abs(y.cpu().numpy() - ref).max() < 2e-5
    y2 = A.preprocess_frames(torch.from_numpy(synth_frames(8, 10, 240, 426)).to(DEV))
    w = torch.rand(10, generator=torch.Generator().manual_seed(0))
    want = AO.mix_frames(y.cpu().numpy(), y2.cpu().numpy(), w.numpy())
    got = A.mix_frames(y, y2, w)
    assert got.data_ptr() == y.data_ptr()                            # in place
    assert np.array_equal(got.cpu().numpy(), want)
    with pytest.raises(ValueError):
        A.preprocess_frames(torch.zeros(1, 3, 8, 8, device=DEV))      # float input is not a decoded frame
    with pytest.raises(RuntimeError):
        A.preprocess_frames(torch.zeros(1, 3, 8, 8, dtype=torch.uint8))


def test_fbank_augment_bit_exact_vs_torchaudio_golden(golden):
    for c in golden["audio"]:
        T, F = c["T"], c["F"]
        gen = torch.Generator().manual_seed(c["seed"])
        d = A.draw_augment_params(1, T, F, c["freqm"], c["timem"], c["noise"], generator=gen,
                                  np_rng=np.random.RandomState(c["seed"]), device=DEV, host_noise=True)
        x = torch.from_numpy(synth_fbank(c["seed"], T, F)).to(DEV).unsqueeze(0)
        y = A.augment_fbank(x, d)
        assert torch.equal(y[0].cpu(), c["out"])


def test_fbank_augment_batch_against_oracle():
    B, T, F = 9, 1024, 128
    gen, rs = torch.Generator().manual_seed(5), np.random.RandomState(5)
    d = A.draw_augment_params(B, T, F, freqm=48, timem=192, noise=True, generator=gen, np_rng=rs, device=DEV)
    assert d.noise.shape == (B, T, F) and d.noise.is_cuda
    x = np.stack([synth_fbank(100 + b, T, F) for b in range(B)])
    y = A.augment_fbank(torch.from_numpy(x).to(DEV), d).cpu().numpy()
    p, sc, nz = d.params.cpu().numpy(), d.scale.cpu().numpy(), d.noise.cpu().numpy()
    assert (p[:, 1] - p[:, 0]).max() < 48 and (p[:, 3] - p[:, 2]).max() < 192 and (p[:, 3] - p[:, 2]).max() > 0
    for b in range(B):
        ref = AO.augment_fbank(x[b], p[b, 0:2], p[b, 2:4], -5.081, 4.4849, noise=nz[b], r=sc[b], shift=int(p[b, 4]))
        assert np.array_equal(y[b], ref)
    # eval configuration: no masks, no noise -> plain normalisation; skip_norm -> identity
    d0 = A.draw_augment_params(B, T, F, device=DEV)
    y0 = A.augment_fbank(torch.from_numpy(x).to(DEV), d0)
    assert torch.equal(y0.cpu(), (torch.from_numpy(x) - (-5.081)) / 4.4849)
    assert torch.equal(A.augment_fbank(torch.from_numpy(x).to(DEV), d0, skip_norm=True).cpu(), torch.from_numpy(x))
