"""GPU audio front end vs torchaudio's own kaldi.fbank outputs (tests/golden/fbank_kaldi.pt, the
call at src/dataloader.py:323) and vs the float64 oracle restatement. Tolerances: 2e-2 absolute in the log domain
(fp32 FFT on near-silent bins), 2e-3 on bins above the noise floor; padding / normalisation exact to fp32 rounding."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import avsiam_b200  # noqa: E402
from oracle import fbank_oracle as FB  # noqa: E402
from oracle.make_golden_fbank import synth_wave  # noqa: E402


def test_fbank_matches_torchaudio_golden(golden_dir):
    for c in torch.load(os.path.join(golden_dir, "fbank_kaldi.pt"), weights_only=False):
        w = synth_wave(c["seed"], c["n"])
        nf = FB.num_frames(c["n"])
        out = avsiam_b200.wav2fbank(w.cuda(), target_length=nf + 3, norm_mean=0.0, norm_std=1.0).cpu()
        ref = c["fbank"]
        assert out.shape == (nf + 3, 128)
        assert torch.all(out[nf:] == 0)                                   # zero padding, un-normalised
        err = (out[:nf] - ref).abs()
        assert float(err.max()) < 2e-2
        assert float(err[ref > -6.0].max()) < 2e-3


def test_fbank_batch_pad_crop_normalise_vs_oracle():
    waves = [synth_wave(11, 40000), synth_wave(12, 40000), synth_wave(13, 40000)]
    batch = torch.stack(waves).cuda()
    for target in (64, 300):                                              # crop (248 frames available) and pad
        out = avsiam_b200.wav2fbank(batch, target_length=target).cpu().numpy()
        assert out.shape == (3, target, 128)
        for b, w in enumerate(waves):
            ref = FB.wav2fbank(w.numpy(), target_length=target)
            assert float(np.abs(out[b] - ref).max()) < 2e-2 / 4.4849 + 1e-5
    # without the loader's waveform mean removal the per-frame DC removal still makes the result identical
    a = avsiam_b200.wav2fbank(batch, target_length=64, remove_mean=False)
    b = avsiam_b200.wav2fbank(batch, target_length=64, remove_mean=True)
    assert float((a - b).abs().max()) < 5e-3
    with pytest.raises(RuntimeError):
        avsiam_b200.wav2fbank(batch.cpu())
