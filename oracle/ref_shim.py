"""TEST INFRASTRUCTURE ONLY — loads the REAL reference model file for golden-vector generation.

Imports /root/reference/src/models/cav_mae_base.py *unmodified*, by file path, with the minimum shims it needs
to execute in this container: the un-vendored dependencies `timm==0.9.5`, `tome`, `ipdb`
and the missing `models.yb_tome` are replaced by stubs that restate only what the model file touches
(Appendix A of SURVEY.md), and the hard-coded `torch.load('/mnt/opr/...')` at cav_mae_base.py:240 returns {}.

/root/reference does not exist on the GPU box; nothing outside oracle/make_golden.py may import this module.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_SRC = os.environ.get("AVSIAM_REFERENCE", "/root/reference") + "/src"


# ---------------------------------------------------------------------------------------------------------
# timm 0.9.5 restated: only the pieces cav_mae_base.py:7-9,25-26,236,259 touch.
# ---------------------------------------------------------------------------------------------------------
class _Mlp(nn.Module):
    """timm.layers.mlp.Mlp: fc1 -> GELU(exact erf) -> drop(0) -> fc2 -> drop(0)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, **_):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _TimmAttn(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _TimmBlock(nn.Module):
    """Parameter container with timm's names: norm1, attn.{qkv,proj}, norm2, mlp.{fc1,fc2}."""

    def __init__(self, dim, num_heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _TimmAttn(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, dim * 4)


class _TimmPatchEmbed(nn.Module):
    def __init__(self):
        super().__init__()
        self.proj = nn.Conv2d(3, 768, kernel_size=16, stride=16)


class _TimmViT(nn.Module):
    """Shape-identical stand-in for timm 'vit_base_patch16_224.augreg_in21k' (weights: torch default init)."""

    def __init__(self):
        super().__init__()
        self.patch_embed = _TimmPatchEmbed()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, 768))
        self.pos_embed = nn.Parameter(torch.randn(1, 197, 768) * 0.02)
        self.norm_pre = nn.Identity()
        self.blocks = nn.Sequential(*[_TimmBlock(768, 12) for _ in range(12)])
        self.norm = nn.LayerNorm(768, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head = nn.Linear(768, 21843)


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def to_2tuple(x):
        return x if isinstance(x, tuple) else (x, x)

    class _Dummy(nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()

    def _noop(*a, **k):
        return None

    timm = mod("timm", create_model=lambda name, pretrained=False, **k: _TimmViT())
    layers = mod("timm.layers", PatchEmbed=_Dummy, Mlp=_Mlp, DropPath=_Dummy, trunc_normal_=_noop,
                 lecun_normal_=_noop, resample_patch_embed=_noop, resample_abs_pos_embed=_noop, RmsNorm=_Dummy,
                 PatchDropout=_Dummy, use_fused_attn=lambda: True, SwiGLUPacked=_Dummy)
    layers_mlp = mod("timm.layers.mlp", Mlp=_Mlp)
    layers.mlp = layers_mlp
    models = mod("timm.models")
    models_layers = mod("timm.models.layers", to_2tuple=to_2tuple, trunc_normal_=_noop, DropPath=_Dummy)
    vt = mod("timm.models.vision_transformer", Attention=_Dummy, Mlp=_Mlp, PatchEmbed=_Dummy, Block=_Dummy)
    timm.layers, timm.models = layers, models
    models.layers, models.vision_transformer = models_layers, vt
    mod("ipdb", set_trace=_noop)
    tome = mod("tome")
    tome.merge = mod("tome.merge", bipartite_soft_matching=_noop, merge_source=_noop, merge_wavg=_noop)
    # `models` package rooted at the reference's directory, without executing its broken __init__.py
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF_SRC, "models")]
    sys.modules["models"] = pkg
    mod("models.yb_tome", yb_bipartite_soft_matching=_noop)


_REF = None


def load_reference():
    """Returns the reference's own `models.cav_mae_base` module object."""
    global _REF
    if _REF is not None:
        return _REF
    if not os.path.isdir(REF_SRC):
        raise RuntimeError(f"reference tree not found at {REF_SRC} (it only exists in the build container)")
    _install_stubs()
    orig_load = torch.load

    def patched_load(f, *a, **k):
        if isinstance(f, str) and f.startswith("/mnt/opr"):
            return {}
        return orig_load(f, *a, **k)

    torch.load = patched_load
    cwd = os.getcwd()
    try:
        _REF = importlib.import_module("models.cav_mae_base")
    finally:
        os.chdir(cwd)
    return _REF


def ensure_process_group():
    """GatherLayer (gather_layer.py:29) needs an initialised process group even at world size 1."""
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        dist.init_process_group("gloo", rank=0, world_size=1)


def build_reference_model(norm_pix_loss=False):
    ref = load_reference()
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.CAVMAE_BASE(audio_length=1024, norm_pix_loss=norm_pix_loss, modality_specific_depth=23, tr_pos=False)
    return m


# ---------------------------------------------------------------------------------------------------------
# The reference's training loop (src/traintest_cavmae_base.py), loaded UNMODIFIED by file path.
# ---------------------------------------------------------------------------------------------------------
class SyntheticAVDataset(torch.utils.data.Dataset):
    """Stands in for dataloader.AudiosetDataset (the loop builds its own DataLoader from it at
    traintest_cavmae_base.py:95-97): seeded (fbank [1024,128], frame [3,224,224], label) samples."""

    n_samples = 4
    seed = 4242

    def __init__(self, dataset_json_file=None, audio_conf=None, label_csv=None, **_):
        g = torch.Generator().manual_seed(self.seed)
        self.a = torch.randn(self.n_samples, 1024, 128, generator=g)
        self.v = torch.randn(self.n_samples, 3, 224, 224, generator=g)

    def __len__(self):
        return self.n_samples

    def __getitem__(self, i):
        return self.a[i], self.v[i], torch.zeros(1)


_TRAINTEST = None


def load_traintest():
    """Returns the reference's own `traintest_cavmae_base` module object. Its imports that cannot be satisfied here
    are stubbed from the OUTSIDE: `utilities` (the shipped utilities/__init__.py is not Python source; the package is
    rebuilt from utilities/util.py and utilities/stats.py, which are), `dataloader` (AudiosetDataset -> synthetic
    samples), `wandb`, `ipdb`, `deepspeed` (imported, never used on the path)."""
    global _TRAINTEST
    if _TRAINTEST is not None:
        return _TRAINTEST
    load_reference()                       # installs timm / tome / ipdb / models stubs
    import importlib.util as iu

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    util_spec = iu.spec_from_file_location("_ref_utilities_util", os.path.join(REF_SRC, "utilities", "util.py"))
    util = iu.module_from_spec(util_spec)
    util_spec.loader.exec_module(util)
    public = {k: v for k, v in util.__dict__.items() if not k.startswith("_")}
    utilities = mod("utilities", **public)
    utilities.__path__ = []
    mod("dataloader", AudiosetDataset=SyntheticAVDataset)
    mod("wandb", log=lambda *a, **k: None, init=lambda *a, **k: None)
    ds = mod("deepspeed")
    ds.profiling = mod("deepspeed.profiling")
    ds.profiling.flops_profiler = mod("deepspeed.profiling.flops_profiler")
    ds.profiling.flops_profiler.profiler = mod("deepspeed.profiling.flops_profiler.profiler", FlopsProfiler=object)
    spec = iu.spec_from_file_location("traintest_cavmae_base", os.path.join(REF_SRC, "traintest_cavmae_base.py"))
    m = iu.module_from_spec(spec)
    argv = sys.argv
    try:
        sys.argv = [os.path.join(REF_SRC, "traintest_cavmae_base.py")]   # the file appends to sys.path from sys.path[0]
        spec.loader.exec_module(m)
    finally:
        sys.argv = argv
    _TRAINTEST = m
    return m
