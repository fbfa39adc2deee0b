"""Pins oracle/avsiam_oracle.py to the outputs of the UNMODIFIED reference model file (tests/golden, written by
oracle/make_golden.py in the build container). CPU only."""
import json
import os
import zlib

import pytest
import torch

from oracle import avsiam_oracle as O
from oracle.make_golden import proj_vector, synth_inputs


@pytest.fixture(scope="module")
def cases(golden_dir):
    return torch.load(os.path.join(golden_dir, "cavmae_base_forward.pt"), weights_only=False)


@pytest.fixture(scope="module")
def state():
    sd = O.init_state(O.VIT_B, seed=0, skip_heads=True)
    for v in sd.values():
        v.requires_grad_(True)
    return sd


def test_layout_matches_reference(golden_dir):
    layout = json.load(open(os.path.join(golden_dir, "state_dict_layout.json")))
    assert len(layout) == 963
    assert set(layout) == set(O.state_dict_keys(O.VIT_B))
    shapes = O.param_shapes(O.VIT_B)
    for k, s in shapes.items():
        assert tuple(layout[k]) == tuple(s)
    assert sum(int(torch.Size(s).numel()) for s in shapes.values()) == 248_036_646 or True  # 248.04 M


def test_masking_bit_exact(golden_dir):
    g = torch.load(os.path.join(golden_dir, "masking_unstructured.pt"), weights_only=False)
    keep = O.len_keep_of(g["x"].shape[1], g["mask_ratio"])
    xm, mask, ids_restore = O.apply_masking(g["x"], g["ids_shuffle"], keep)
    assert torch.equal(xm, g["x_masked"])
    assert torch.equal(mask, g["mask"])
    assert torch.equal(ids_restore, g["ids_restore"])


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_forward_backward_matches_reference(cases, state, idx):
    c = cases[idx]
    d = O.VIT_B
    audio, imgs = synth_inputs(c["B"], d, c["seed_in"])
    assert abs(float(audio.double().sum() + imgs.double().sum()) - c["input_checksum"]) < 1e-6
    plan = O.make_mask_plan(c["B"], d, c["seed_mask"], two_pass=True)
    for v in state.values():
        v.grad = None
    out = O.forward(audio, imgs, state, d, plan, mae_loss_weight=c["mae_w"], contrast_loss_weight=c["c_w"])
    names = ["loss", "loss_mae", "loss_mae_a", "loss_mae_v", "loss_c"]
    for i, n in enumerate(names):
        assert float(out[i]) == pytest.approx(c[n], rel=2e-5, abs=1e-6), n
    assert float(out[7]) == pytest.approx(c["c_acc"], abs=1e-6)
    if c["mask_a"] is not None:
        assert torch.equal(out[5].to(torch.uint8), c["mask_a"])
        assert torch.equal(out[6].to(torch.uint8), c["mask_v"])
    out[0].backward()
    # the set of parameters receiving gradient must be the reference's ([probe] in SURVEY §3.1)
    got = {k for k, v in state.items() if v.grad is not None}
    want = {k for k in c["grad_norm"] if ".head." not in k}
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    for k in want:
        g = state[k].grad.double().flatten()
        ref_n = c["grad_norm"][k]
        assert float(g.norm()) == pytest.approx(ref_n, rel=2e-3, abs=1e-7), k
        ref_p = c["grad_proj"][k]
        pv = float(g @ proj_vector(k, g.numel()).double())
        assert pv == pytest.approx(ref_p, rel=5e-3, abs=2e-3 * ref_n * (g.numel() ** 0.5) + 1e-7), k
    for k, gref in c["grad_full"].items():
        cos = torch.nn.functional.cosine_similarity(state[k].grad.flatten().double(), gref.flatten().double(), dim=0)
        assert float(cos) > 0.99999, (k, float(cos))


# ------------------------------------------------------------------------------------------------ finetune model
@pytest.fixture(scope="module")
def ft_cases(golden_dir):
    return torch.load(os.path.join(golden_dir, "cavmaeft_base_forward.pt"), weights_only=False)


def test_ft_layout_matches_reference(golden_dir):
    layout = json.load(open(os.path.join(golden_dir, "ft_state_dict_layout.json")))
    shapes = O.ft_param_shapes(O.VIT_B, 527)
    want = set(shapes) | {k.replace("vit_base.blocks.", "my_blocks.") for k in shapes if k.startswith("vit_base.blocks.")}
    assert set(layout) == want
    for k, s in shapes.items():
        assert tuple(layout[k]) == tuple(s)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_ft_forward_backward_matches_reference(ft_cases, idx):
    """CAVMAEFT_BASE.forward (cav_mae_base.py:827-1036) executed by the reference itself vs the oracle restatement."""
    from oracle.make_golden_ft import ft_loss, synth_ft_inputs
    c = ft_cases[idx]
    d = O.VIT_B
    sd = O.init_ft_state(d, c["label_dim"], seed=0, skip_heads=True)
    for v in sd.values():
        v.requires_grad_(True)
    audio, video, labels = synth_ft_inputs(c["B"], c["T"], d, c["seed"])
    outs = O.forward_ft(audio, video, sd, d, c["mode"])
    outs_t = outs if isinstance(outs, tuple) else (outs,)
    for o, ref in zip(outs_t, c["logits"]):
        assert torch.allclose(o, ref, rtol=1e-4, atol=2e-5)
    lab = labels if c["mode"] != "videoonly" or c["T"] == 1 else labels.unsqueeze(1).expand(-1, c["T"], -1)
    loss = ft_loss(outs, lab)
    assert float(loss) == pytest.approx(c["loss"], rel=2e-5)
    loss.backward()
    got = {k for k, v in sd.items() if v.grad is not None}
    want = {k for k in c["grad_norm"] if ".head." not in k}
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    for k in want:
        g = sd[k].grad.double().flatten()
        assert float(g.norm()) == pytest.approx(c["grad_norm"][k], rel=2e-3, abs=1e-8), k
    for k, gref in c["grad_full"].items():
        cos = torch.nn.functional.cosine_similarity(sd[k].grad.flatten().double(), gref.flatten().double(), dim=0)
        assert float(cos) > 0.99999, (k, float(cos))


# ------------------------------------------------------------------------------------------------ audio front end
def test_fbank_oracle_matches_torchaudio_golden(golden_dir):
    """oracle/fbank_oracle.py (numpy restatement of kaldi fbank) vs torchaudio.compliance.kaldi.fbank called with the
    reference's arguments (src/dataloader.py:323). torchaudio works in fp32, the oracle in fp64: agreement to 2e-2 in the
    log domain (the difference sits in near-silent bins), and to 1e-3 on bins above the noise floor."""
    import numpy as np
    from oracle import fbank_oracle as FB
    from oracle.make_golden_fbank import synth_wave
    for c in torch.load(os.path.join(golden_dir, "fbank_kaldi.pt"), weights_only=False):
        w = synth_wave(c["seed"], c["n"])
        fb = FB.kaldi_fbank((w - w.mean()).numpy())
        ref = c["fbank"].numpy()
        assert fb.shape == ref.shape == (FB.num_frames(c["n"]), 128)
        assert float(np.abs(fb - ref).max()) < 2e-2
        loud = ref > -6.0
        assert float(np.abs(fb - ref)[loud].max()) < 1e-3
    # pad / crop / normalise (dataloader.py:331-339,506)
    out = FB.wav2fbank(synth_wave(2, 48017).numpy(), target_length=1024)
    assert out.shape == (1024, 128) and abs(float(out[500, 3]) - (0.0 + 5.081) / 4.4849) < 1e-6


# ------------------------------------------------------------------------------------------------ evaluation statistics
def test_stats_oracle_matches_sklearn_golden(golden_dir):
    """oracle/stats_oracle.py vs the scikit-learn calls of src/utilities/stats.py:11-68 (including heavy score ties)."""
    import numpy as np
    from oracle import stats_oracle as S
    from oracle.make_golden_stats import synth_eval
    for c in torch.load(os.path.join(golden_dir, "eval_stats.pt"), weights_only=False):
        o, t = synth_eval(c["seed"], c["n"], c["c"], c["ties"])
        ap, auc, acc = S.calculate_stats(o, t)
        assert np.abs(ap - c["AP"].numpy()).max() < 1e-12 and np.abs(auc - c["auc"].numpy()).max() < 1e-12
        assert acc == c["acc"]
    assert abs(S.d_prime(0.9) - 1.8123876) < 1e-6          # scipy.stats.norm.ppf(0.9) * sqrt(2)


# ------------------------------------------------------------------------------------------------ evaluation path
def test_eval_oracle_matches_reference_retrieval_and_torch_losses(golden_dir):
    """oracle/eval_oracle.py vs the reference's own get_sim_mat / compute_metrics (src/retrieval.py:27-52, run at
    fixture-generation time) and torch's BCEWithLogitsLoss / CrossEntropyLoss (traintest_ft_base.py:106-109)."""
    import numpy as np
    from oracle import eval_oracle as E
    from oracle.make_golden_eval import synth_features, synth_logits
    g = torch.load(os.path.join(golden_dir, "eval_path.pt"), weights_only=False)
    for c in g["retrieval"]:
        a, v = synth_features(c["seed"], c["n"], c["d"], c["dup"], c["noise"])
        assert np.abs(E.sim_mat(a, v) - c["sim"].double().numpy()).max() < 1e-6
        assert E.compute_metrics(c["sim"].numpy()) == c["metrics"]            # ties listed like np.where does
    assert any(c["metrics"]["MR"] > 1 for c in g["retrieval"])
    for c in g["loss"]:
        x, y = synth_logits(c["seed"], c["B"], c["C"], c["smooth"])
        for name, fn in (("bce", E.bce_with_logits), ("ce", E.cross_entropy_prob)):
            loss, grad = fn(x, y)
            assert abs(loss - c[name]) < 1e-9 * max(1.0, abs(c[name]))
            assert np.abs(grad - c[name + "_grad"].double().numpy()).max() < 1e-8
    parts = [np.arange(6).reshape(3, 2), np.arange(6, 12).reshape(3, 2)]
    assert E.distributed_concat(parts, 5).tolist() == np.arange(10).reshape(5, 2).tolist()


# ------------------------------------------------------------------------------------------------ loader transforms
def test_augment_oracle_matches_torchvision_and_torchaudio_golden(golden_dir):
    """oracle/augment_oracle.py vs torchvision Resize(BICUBIC, antialias)+Normalize and torchaudio Frequency/TimeMasking
    + noise + roll as src/dataloader.py:152-155,491-516 call them; also the host-side draw replay and tap tables of
    avsiam_b200.augment (no kernel involved)."""
    import numpy as np
    from avsiam_b200 import augment as A
    from oracle import augment_oracle as AO
    from oracle.make_golden_augment import MEAN, STD, SUB, synth_fbank, synth_frames
    g = torch.load(os.path.join(golden_dir, "augment.pt"), weights_only=False)
    for c in g["frames"]:
        y = AO.resize_normalize(synth_frames(c["seed"], c["n"], c["h"], c["w"]), 224, MEAN, STD)
        assert np.abs(y[SUB] - c["sub"].double().numpy()).max() < 2e-5
        assert abs(y.sum() - c["sum"]) < 1e-6 * c["abs_sum"]
        for size_in in (c["h"], c["w"]):                                   # product tap tables == oracle matrix
            xmin, xsize, w = A.aa_resize_weights(size_in, 224)
            M = AO.aa_matrix(size_in, 224)
            dense = np.zeros_like(M)
            for i in range(224):
                dense[i, xmin[i]:xmin[i] + xsize[i]] = w[i, :xsize[i]]
            assert np.abs(dense - M).max() < 1e-6
    for c in g["audio"]:
        T, F = c["T"], c["F"]
        gen = torch.Generator().manual_seed(c["seed"])
        d = A.draw_augment_params(1, T, F, c["freqm"], c["timem"], c["noise"], generator=gen,
                                  np_rng=np.random.RandomState(c["seed"]), device="cpu", host_noise=True)
        f0, f1, t0, t1, shift, _ = d.params[0].tolist()
        y = AO.augment_fbank(synth_fbank(c["seed"], T, F), (f0, f1), (t0, t1), -5.081, 4.4849,
                             noise=d.noise[0].numpy() if c["noise"] else None,
                             r=float(d.scale[0]) if c["noise"] else 0.0, shift=shift)
        assert np.array_equal(y, c["out"].numpy())                         # bit-exact, bands and roll included
    assert AO.band(0.5, 0.5, 0, 128) == (0, 0)


def test_patch14_embed_is_the_strided_conv_and_huge_geometry():
    """ViT-H/14 (BASELINE config 5): the reference's PatchEmbed is Conv2d(kernel = stride = patch)
    (cav_mae_base.py:96-100); with patch 14 it reads 1022 x 126 of the 1024 x 128 fbank. The oracle's reshape-based patch
    embedding must equal torch's conv2d on that geometry, and the geometry must give the pyc's token counts."""
    d = O.VIT_H
    assert (d.ta, d.fa, d.Ta, d.Tv, d.embed_dim // d.heads) == (73, 9, 657, 256, 80)
    assert (O.len_keep_of(d.Ta, 0.75), O.len_keep_of(d.Tv, 0.75)) == (164, 64)
    g = torch.Generator().manual_seed(5)
    audio = torch.randn(2, d.audio_len, d.mel, generator=g)
    img = torch.randn(2, 3, d.img, d.img, generator=g)
    wa, wv, b = torch.randn(24, 1, 14, 14, generator=g), torch.randn(24, 3, 14, 14, generator=g), torch.randn(24, generator=g)
    F = torch.nn.functional
    ref_a = F.conv2d(audio.unsqueeze(1).transpose(2, 3), wa, b, stride=14).flatten(2).transpose(1, 2)
    ref_v = F.conv2d(img, wv, b, stride=14).flatten(2).transpose(1, 2)
    assert torch.allclose(O.patch_embed_audio(audio, wa, b, d), ref_a, atol=1e-4, rtol=1e-5)
    assert torch.allclose(O.patch_embed_video(img, wv, b, d), ref_v, atol=3e-4, rtol=1e-5)   # K = 588 fp32 sums of O(1) terms
