"""TEST INFRASTRUCTURE ONLY — CPU oracle: a plain-PyTorch fp32 restatement of the AVSiam pretraining hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file;
the product (avsiam_b200/) never does.

Every function cites the reference lines it follows (paths relative to /root/reference/). The restatement is
*functional* (it works on a flat {key: tensor} state in the reference's checkpoint key layout) and is pinned
to the reference's own model file by tests/golden/*.pt, produced by oracle/make_golden.py, which executes the
unmodified reference here and records its outputs for the same seeded weights, inputs and supplied mask indices.
The reference ships no tests or golden vectors of its own, so those fixtures are the pin.

Third-party arithmetic restated here (absent from /root/reference): timm==0.9.5 (requirements.txt:101) —
`Mlp` = fc2(GELU_erf(fc1 x)), final LayerNorm eps 1e-6, `norm_pre` = Identity.
"""
from __future__ import annotations

import math
import random as _pyrandom
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]


@dataclass(frozen=True)
class Dims:
    """Model geometry. Defaults = CAVMAE_BASE literals (cav_mae_base.py:219-222,249-261,316-329)."""
    embed_dim: int = 768
    depth: int = 12
    heads: int = 12
    dec_dim: int = 512
    dec_depth: int = 8
    dec_heads: int = 16
    patch: int = 16
    audio_len: int = 1024
    mel: int = 128
    img: int = 224
    in_chans: int = 3
    head_classes: int = 21843  # timm in21k head (unused on the path, present in the checkpoint)

    @property
    def fa(self):  # mel-bin blocks (f) — cav_mae_base.py:403 "audio patch is in shape [f,t]"
        return self.mel // self.patch

    @property
    def ta(self):
        return self.audio_len // self.patch

    @property
    def Ta(self):
        return self.fa * self.ta

    @property
    def Tv(self):
        return (self.img // self.patch) ** 2


VIT_B = Dims()
VIT_L = Dims(embed_dim=1024, depth=24, heads=16, dec_depth=6)   # SURVEY.md Appendix C (cav_mae_large.cpython-39.pyc)
VIT_H = Dims(embed_dim=1280, depth=32, heads=16, patch=14)      # Appendix C (cav_mae_huge.cpython-39.pyc): head_dim 80,
#                                                                 657 + 256 tokens; contrastive-only as shipped
TINY = Dims(embed_dim=128, depth=2, heads=2, dec_dim=64, dec_depth=2, dec_heads=2, patch=16, audio_len=256, mel=32,
            img=64, head_classes=16)

LN_EPS_BLOCK = 1e-5   # nn.LayerNorm default — cav_mae_base.py:116,331
LN_EPS_FINAL = 1e-6   # timm vit norm / norm_a — SURVEY.md Appendix A


# ---------------------------------------------------------------------------------------------------------
# Parameter set (checkpoint layout; cav_mae_base.py:236-337)
# ---------------------------------------------------------------------------------------------------------
def _block_shapes(D: int, hidden: int, with_modality_norms: bool = True):
    sh = OrderedDict()
    sh["norm1.weight"] = (D,); sh["norm1.bias"] = (D,)
    if with_modality_norms:
        sh["norm1_a.weight"] = (D,); sh["norm1_a.bias"] = (D,)
        sh["norm1_v.weight"] = (D,); sh["norm1_v.bias"] = (D,)
    sh["attn.qkv.weight"] = (3 * D, D); sh["attn.qkv.bias"] = (3 * D,)
    sh["attn.proj.weight"] = (D, D); sh["attn.proj.bias"] = (D,)
    sh["norm2.weight"] = (D,); sh["norm2.bias"] = (D,)
    if with_modality_norms:
        sh["norm2_a.weight"] = (D,); sh["norm2_a.bias"] = (D,)
        sh["norm2_v.weight"] = (D,); sh["norm2_v.bias"] = (D,)
    sh["mlp.fc1.weight"] = (hidden, D); sh["mlp.fc1.bias"] = (hidden,)
    sh["mlp.fc2.weight"] = (D, hidden); sh["mlp.fc2.bias"] = (D,)
    return sh


def _vit_shapes(d: Dims):
    D, p = d.embed_dim, d.patch
    sh = OrderedDict()
    sh["cls_token"] = (1, 1, D)
    sh["pos_embed"] = (1, d.Tv + 1, D)
    sh["pos_embed_a"] = (1, d.Ta, D)
    sh["patch_embed.proj.weight"] = (D, d.in_chans, p, p); sh["patch_embed.proj.bias"] = (D,)
    for i in range(d.depth):
        for k, s in _block_shapes(D, 4 * D).items():
            sh[f"blocks.{i}.{k}"] = s
    sh["norm.weight"] = (D,); sh["norm.bias"] = (D,)
    sh["head.weight"] = (d.head_classes, D); sh["head.bias"] = (d.head_classes,)
    sh["patch_embed_a.proj.weight"] = (D, 1, p, p); sh["patch_embed_a.proj.bias"] = (D,)
    sh["norm_a.weight"] = (D,); sh["norm_a.bias"] = (D,)
    return sh


def param_shapes(d: Dims = VIT_B) -> "OrderedDict[str, tuple]":
    """Unique parameters (no aliases) of CAVMAE_BASE, keyed as in its state_dict."""
    D, Dd, p = d.embed_dim, d.dec_dim, d.patch
    sh = OrderedDict()
    for k, s in _vit_shapes(d).items():
        sh[f"vit_base.{k}"] = s
    sh["my_patch_embed.proj.weight"] = (D, d.in_chans, p, p); sh["my_patch_embed.proj.bias"] = (D,)
    sh["my_patch_embed_a.proj.weight"] = (D, 1, p, p); sh["my_patch_embed_a.proj.bias"] = (D,)
    for k, s in _vit_shapes(d).items():
        sh[f"ast_base.{k}"] = s
    for name in ("mm_layer_1", "mm_layer_2"):
        for k, s in _block_shapes(D, 4 * D).items():
            sh[f"{name}.{k}"] = s
    sh["decoder_embed.weight"] = (Dd, D); sh["decoder_embed.bias"] = (Dd,)
    sh["decoder_pos_embed_a"] = (1, d.Ta, Dd)
    sh["decoder_pos_embed_v"] = (1, d.Tv, Dd)
    sh["mask_token"] = (1, 1, Dd)
    for i in range(d.dec_depth):
        for k, s in _block_shapes(Dd, 4 * Dd).items():
            sh[f"decoder_blocks.{i}.{k}"] = s
    sh["decoder_norm.weight"] = (Dd,); sh["decoder_norm.bias"] = (Dd,)
    sh["decoder_pred_a.weight"] = (p * p, Dd); sh["decoder_pred_a.bias"] = (p * p,)
    sh["decoder_pred_v.weight"] = (p * p * d.in_chans, Dd); sh["decoder_pred_v.bias"] = (p * p * d.in_chans,)
    sh["decoder_modality_a"] = (1, 1, Dd); sh["decoder_modality_v"] = (1, 1, Dd)
    return sh


def state_dict_keys(d: Dims = VIT_B) -> List[str]:
    """All state_dict keys including the `my_blocks.*` aliases of `vit_base.blocks.*` (cav_mae_base.py:278)."""
    keys = list(param_shapes(d).keys())
    keys += [k.replace("vit_base.blocks.", "my_blocks.") for k in keys if k.startswith("vit_base.blocks.")]
    return keys


def init_state(d: Dims = VIT_B, seed: int = 0, skip_heads: bool = False) -> State:
    """Seeded random weights (there are no pretrained weights offline). Deterministic in (dims, seed): every
    tensor is drawn from its own generator keyed by (seed, crc(key)) so subsets can be regenerated alone.
    Non-degenerate on purpose: biases, pos-embeds, mask token and decoder embeddings are non-zero so every add
    on the path is exercised (the reference zero-inits some of them, cav_mae_base.py:312-314,336-337)."""
    import zlib

    sd: State = OrderedDict()
    for k, shape in param_shapes(d).items():
        if skip_heads and ".head." in k:
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2**31))
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if "norm" in k and k.endswith(".weight"):
            t = 1.0 + 0.1 * t
        elif "norm" in k and k.endswith(".bias"):
            t = 0.05 * t
        elif k.endswith(".bias"):
            t = 0.02 * t
        elif "pos_embed" in k or "token" in k or "modality" in k:
            t = 0.02 * t
        elif "patch_embed" in k:
            t = t * (1.0 / math.sqrt(shape[1] * shape[2] * shape[3]))
        else:  # Linear weights: fan-in scaled so activations stay O(1) through 22 blocks
            t = t * (0.7 / math.sqrt(shape[-1]))
        sd[k] = t
    return sd


def with_aliases(sd: State) -> State:
    out = OrderedDict(sd)
    for k in list(sd.keys()):
        if k.startswith("vit_base.blocks."):
            out[k.replace("vit_base.blocks.", "my_blocks.")] = sd[k]
    return out


# ---------------------------------------------------------------------------------------------------------
# Layers
# ---------------------------------------------------------------------------------------------------------
def patch_embed_audio(audio: torch.Tensor, w: torch.Tensor, b: torch.Tensor, d: Dims) -> torch.Tensor:
    """cav_mae_base.py:444-448 + PatchEmbed.forward :98-100 for the audio branch.
    audio [B, T, F] -> unsqueeze(1).transpose(2,3) = [B,1,F,T]; Conv2d(1,D,16,16); token = f*ta + t;
    the 256-vector of a patch is the (freq, time) tile (SURVEY.md §8a identity 3)."""
    B = audio.shape[0]
    p = d.patch
    # Conv2d(kernel = stride = p) never reads past the last whole patch: with ViT-H's patch 14 the 1024 x 128 fbank
    # contributes its first 1022 x 126 samples (73 x 9 = 657 tokens); a no-op when p divides both
    audio = audio[:, :d.ta * p, :d.fa * p]
    x = audio.transpose(1, 2).reshape(B, d.fa, p, d.ta, p)           # [B, f, pf, t, pt]
    x = x.permute(0, 1, 3, 2, 4).reshape(B, d.Ta, p * p)             # token (f,t); vec (pf,pt)
    return x @ w.reshape(w.shape[0], -1).t() + b


def patch_embed_video(img: torch.Tensor, w: torch.Tensor, b: torch.Tensor, d: Dims) -> torch.Tensor:
    """cav_mae_base.py:453 + :98-100. token = h*14+w; K order (c, p, q) (identity 4)."""
    B, C = img.shape[0], img.shape[1]
    p = d.patch
    g = d.img // p
    img = img[:, :, :g * p, :g * p]
    x = img.reshape(B, C, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, C * p * p)
    return x @ w.reshape(w.shape[0], -1).t() + b


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def attention(x, sd: State, pfx: str, heads: int):
    """Attention.forward, cav_mae_base.py:58-83 (fused SDPA branch: scale = head_dim**-0.5, no mask, p=0)."""
    B, N, C = x.shape
    hd = C // heads
    qkv = F.linear(x, sd[pfx + "qkv.weight"], sd[pfx + "qkv.bias"]).reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    att = att.softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])


def mlp(x, sd: State, pfx: str):
    """timm.layers.mlp.Mlp (0.9.5): fc2(GELU_erf(fc1(x))); used at cav_mae_base.py:138-143."""
    h = F.gelu(F.linear(x, sd[pfx + "fc1.weight"], sd[pfx + "fc1.bias"]))
    return F.linear(h, sd[pfx + "fc2.weight"], sd[pfx + "fc2.bias"])


CHECKPOINT_BLOCKS = False   # True: recompute each block in backward (same arithmetic; lets the fp32 oracle run the
#                             benchmark's B = 256 on one GPU as the checker — the S x S attention matrices of the
#                             decoder alone are 8 GB per block in fp32)


def block(x, sd: State, pfx: str, heads: int, modality: Optional[str]):
    """Block.forward, cav_mae_base.py:149-193: pre-LN, LayerNorm set chosen by modality in {None,'a','v'}."""
    if CHECKPOINT_BLOCKS and torch.is_grad_enabled():
        from torch.utils.checkpoint import checkpoint
        return checkpoint(_block, x, sd, pfx, heads, modality, use_reentrant=False)
    return _block(x, sd, pfx, heads, modality)


def _block(x, sd: State, pfx: str, heads: int, modality: Optional[str]):
    sfx = "" if modality is None else "_" + modality
    x = x + attention(layer_norm(x, sd[f"{pfx}norm1{sfx}.weight"], sd[f"{pfx}norm1{sfx}.bias"], LN_EPS_BLOCK), sd,
                      pfx + "attn.", heads)
    x = x + mlp(layer_norm(x, sd[f"{pfx}norm2{sfx}.weight"], sd[f"{pfx}norm2{sfx}.bias"], LN_EPS_BLOCK), sd,
                pfx + "mlp.")
    return x


# ---------------------------------------------------------------------------------------------------------
# Masking (int path — bit exact)
# ---------------------------------------------------------------------------------------------------------
def len_keep_of(L: int, ratio: float) -> int:
    return int(L * (1 - ratio))  # cav_mae_base.py:372,399 (python float arithmetic, e.g. int(512*(1-0.6000000000000001)))


def apply_masking(x: torch.Tensor, ids_shuffle: torch.Tensor, len_keep: int):
    """Everything in random_masking_* after the argsort (cav_mae_base.py:378-390 / :427-439)."""
    N, L, D = x.shape
    ids_restore = torch.argsort(ids_shuffle, dim=1)
    ids_keep = ids_shuffle[:, :len_keep]
    x_masked = torch.gather(x, 1, ids_keep.unsqueeze(-1).expand(-1, -1, D))
    mask = torch.ones(N, L, dtype=x.dtype, device=x.device)
    mask[:, :len_keep] = 0
    mask = torch.gather(mask, 1, ids_restore)
    return x_masked, mask, ids_restore


def unstructured_noise(N: int, L: int, gen: torch.Generator) -> torch.Tensor:
    return torch.rand(N, L, generator=gen)  # cav_mae_base.py:374


def structured_noise_tf(N: int, L: int, ratio: float, t: int, f: int, gen: torch.Generator,
                        rng: _pyrandom.Random) -> torch.Tensor:
    """random_masking_structured mode 'tf' noise construction, cav_mae_base.py:401-423."""
    noise = torch.rand(N, L, generator=gen).reshape(N, f, t)
    for i in range(N):
        for k in rng.sample(range(t), int(t * ratio * 0.7)):
            noise[i, :, k] = 1.1
    for i in range(N):
        for k in rng.sample(range(f), int(f * ratio * 0.7)):
            noise[i, k, :] = 1.1
    return noise.reshape(N, L)


def stable_argsort(noise: torch.Tensor) -> torch.Tensor:
    """argsort with index tie-break (the reference's torch.argsort leaves ties implementation-defined; the CUDA
    kernel and this oracle both define them by ascending index)."""
    return torch.argsort(noise, dim=1, stable=True)


@dataclass
class MaskPlan:
    """The random draws of one CAVMAE_BASE.forward, made explicit so oracle, reference and CUDA path consume
    identical indices (SURVEY.md §8c 'identical supplied')."""
    ids_shuffle_a: Optional[torch.Tensor] = None        # [B, Ta]  pass 2 (forward_encoder :476)
    ids_shuffle_v: Optional[torch.Tensor] = None        # [B, Tv]  pass 2 (:477)
    perm_a: Optional[torch.Tensor] = None               # [B] chunk permutation, mmixed :533
    perm_v: Optional[torch.Tensor] = None               # [B] :537
    chunk_ids_a: List[torch.Tensor] = field(default_factory=list)  # 5 x [n_i, Ta]  (:546)
    chunk_ids_v: List[torch.Tensor] = field(default_factory=list)  # 5 x [n_i, Tv]  (:549)

    def to(self, device) -> "MaskPlan":
        mv = lambda t: None if t is None else t.to(device)
        return MaskPlan(mv(self.ids_shuffle_a), mv(self.ids_shuffle_v), mv(self.perm_a), mv(self.perm_v),
                        [t.to(device) for t in self.chunk_ids_a], [t.to(device) for t in self.chunk_ids_v])


N_CHUNKS = 5  # cav_mae_base.py:534,538


def chunk_sizes(B: int, n: int = N_CHUNKS) -> List[int]:
    """torch.chunk(perm, 5) sizes (cav_mae_base.py:534): ceil-sized chunks, possibly fewer than 5."""
    cs = -(-B // n)
    out = []
    left = B
    while left > 0:
        out.append(min(cs, left))
        left -= out[-1]
    return out


def make_mask_plan(B: int, d: Dims, seed: int, two_pass: bool = True, ratio: float = 0.75) -> MaskPlan:
    g = torch.Generator().manual_seed(seed)
    rng = _pyrandom.Random(seed)
    plan = MaskPlan()
    plan.ids_shuffle_a = stable_argsort(unstructured_noise(B, d.Ta, g))
    plan.ids_shuffle_v = stable_argsort(unstructured_noise(B, d.Tv, g))
    if two_pass:
        plan.perm_a = torch.randperm(B, generator=g)
        plan.perm_v = torch.randperm(B, generator=g)
        for i, n in enumerate(chunk_sizes(B)):
            r = 0 + 0.2 * i
            plan.chunk_ids_a.append(stable_argsort(structured_noise_tf(n, d.Ta, r, d.ta, d.fa, g, rng)))
            plan.chunk_ids_v.append(stable_argsort(unstructured_noise(n, d.Tv, g)))
    return plan


# ---------------------------------------------------------------------------------------------------------
# Encoders / decoder / losses
# ---------------------------------------------------------------------------------------------------------
def embed_tokens(audio, imgs, sd: State, d: Dims, vit: str = "vit_base."):
    """cav_mae_base.py:444-455 / :511-522: x = x + pos; x = x + norm_pre(x) with norm_pre = Identity => 2(x+pos)."""
    a = patch_embed_audio(audio, sd[vit + "patch_embed_a.proj.weight"], sd[vit + "patch_embed_a.proj.bias"], d)
    a = a + sd[vit + "pos_embed_a"]
    a = a + a
    v = patch_embed_video(imgs, sd[vit + "patch_embed.proj.weight"], sd[vit + "patch_embed.proj.bias"], d)
    v = v + sd[vit + "pos_embed"][:, 1:]
    v = v + v
    return a, v


def forward_encoder(audio, imgs, sd: State, d: Dims, plan: MaskPlan, ratio_a=0.75, ratio_v=0.75):
    """forward_encoder, cav_mae_base.py:441-504: video through vit_base.blocks[i](v,'v'); audio through the
    deep copy ast_base.blocks[i](a) with modality=None norms; final vit_base.norm / ast_base.norm_a."""
    a, v = embed_tokens(audio, imgs, sd, d)
    a, mask_a, ids_restore_a = apply_masking(a, plan.ids_shuffle_a, len_keep_of(d.Ta, ratio_a))
    v, mask_v, ids_restore_v = apply_masking(v, plan.ids_shuffle_v, len_keep_of(d.Tv, ratio_v))
    for i in range(d.depth):
        v = block(v, sd, f"vit_base.blocks.{i}.", d.heads, "v")
        a = block(a, sd, f"ast_base.blocks.{i}.", d.heads, None)
    cv = layer_norm(v, sd["vit_base.norm.weight"], sd["vit_base.norm.bias"], LN_EPS_FINAL)
    ca = layer_norm(a, sd["ast_base.norm_a.weight"], sd["ast_base.norm_a.bias"], LN_EPS_FINAL)
    x = torch.cat((ca, cv), dim=1)
    return x, mask_a, ids_restore_a, mask_v, ids_restore_v, ca, cv


def forward_encoder_mmixed(audio, imgs, sd: State, d: Dims, plan: MaskPlan):
    """forward_encoder_mmixed, cav_mae_base.py:508-594: 5 chunks at mask ratio 0.2*i (audio structured 'tf',
    video unstructured), all through the SHARED vit_base.blocks with 'a' / 'v' norms; final LN + token mean;
    results restored to batch order."""
    a, v = embed_tokens(audio, imgs, sd, d)
    B = a.shape[0]
    sizes = chunk_sizes(B)
    idx_a = torch.split(plan.perm_a, sizes)
    idx_v = torch.split(plan.perm_v, sizes)
    outs_a, outs_v = [], []
    for i in range(len(sizes)):
        r = 0 + 0.2 * i
        xa, _, _ = apply_masking(a[idx_a[i]], plan.chunk_ids_a[i], len_keep_of(d.Ta, r))
        xv, _, _ = apply_masking(v[idx_v[i]], plan.chunk_ids_v[i], len_keep_of(d.Tv, r))
        for l in range(d.depth):
            xa = block(xa, sd, f"vit_base.blocks.{l}.", d.heads, "a")
            xv = block(xv, sd, f"vit_base.blocks.{l}.", d.heads, "v")
        outs_v.append(layer_norm(xv, sd["vit_base.norm.weight"], sd["vit_base.norm.bias"], LN_EPS_FINAL)
                      .mean(dim=1, keepdim=True))
        outs_a.append(layer_norm(xa, sd["vit_base.norm_a.weight"], sd["vit_base.norm_a.bias"], LN_EPS_FINAL)
                      .mean(dim=1, keepdim=True))
    cv = torch.cat(outs_v, 0)
    ca = torch.cat(outs_a, 0)
    inv_a = torch.empty_like(plan.perm_a); inv_a[plan.perm_a] = torch.arange(B, device=plan.perm_a.device)   # :584-586
    inv_v = torch.empty_like(plan.perm_v); inv_v[plan.perm_v] = torch.arange(B, device=plan.perm_v.device)
    return ca[inv_a], cv[inv_v]


def forward_decoder(x, ids_restore_a, ids_restore_v, keep_a: int, keep_v: int, sd: State, d: Dims):
    """forward_decoder, cav_mae_base.py:597-638."""
    x = F.linear(x, sd["decoder_embed.weight"], sd["decoder_embed.bias"])
    B, _, Dd = x.shape
    mt = sd["mask_token"]
    a_ = torch.cat([x[:, :keep_a], mt.expand(B, d.Ta - keep_a, Dd)], 1)
    a_ = torch.gather(a_, 1, ids_restore_a.unsqueeze(-1).expand(-1, -1, Dd))
    v_ = torch.cat([x[:, keep_a:], mt.expand(B, d.Tv - keep_v, Dd)], 1)
    v_ = torch.gather(v_, 1, ids_restore_v.unsqueeze(-1).expand(-1, -1, Dd))
    a_ = a_ + sd["decoder_pos_embed_a"] + sd["decoder_modality_a"]
    v_ = v_ + sd["decoder_pos_embed_v"] + sd["decoder_modality_v"]
    x = torch.cat([a_, v_], 1)
    for i in range(d.dec_depth):
        x = block(x, sd, f"decoder_blocks.{i}.", d.dec_heads, None)
    x = layer_norm(x, sd["decoder_norm.weight"], sd["decoder_norm.bias"], LN_EPS_BLOCK)
    pa = F.linear(x[:, :d.Ta], sd["decoder_pred_a.weight"], sd["decoder_pred_a.bias"])
    pv = F.linear(x[:, d.Ta:], sd["decoder_pred_v.weight"], sd["decoder_pred_v.bias"])
    return pa, pv


def patchify_target_audio(audio, d: Dims):
    """patchify on [B,1,F,T] (cav_mae_base.py:343-351,666-668): target vec order (p,q,c) with c=1 -> (pf,pt)."""
    B, p = audio.shape[0], d.patch
    x = audio.transpose(1, 2).reshape(B, d.fa, p, d.ta, p).permute(0, 1, 3, 2, 4)
    return x.reshape(B, d.Ta, p * p)


def patchify_target_video(img, d: Dims):
    """patchify 'nchpwq->nhwpqc' (cav_mae_base.py:349): target vec order (p,q,c) — NOT the embed's (c,p,q)."""
    B, C, p = img.shape[0], img.shape[1], d.patch
    g = d.img // p
    x = img.reshape(B, C, g, p, g, p).permute(0, 2, 4, 3, 5, 1)
    return x.reshape(B, g * g, p * p * C)


def mae_loss(target, pred, mask):
    """forward_mae_loss, cav_mae_base.py:679-683 (norm_pix_loss branch is commented out in the reference)."""
    loss = ((pred - target) ** 2).mean(dim=-1)
    return (loss * mask).sum() / mask.sum()


def contrastive(audio_rep, video_rep, bidirect: bool = True):
    """forward_contrastive, cav_mae_base.py:641-661: logits = a_hat v_hat^T / 0.05, log_softmax over dim 0."""
    a = F.normalize(audio_rep, dim=-1)
    v = F.normalize(video_rep, dim=-1)
    total = a @ v.t() / 0.05
    n = total.shape[0]
    ar = torch.arange(n, device=total.device)
    nce_1 = -torch.mean(torch.diag(F.log_softmax(total, dim=0)))
    acc_1 = (torch.argmax(total, dim=0) == ar).sum() / n
    if not bidirect:
        return nce_1, acc_1
    nce_2 = -torch.mean(torch.diag(F.log_softmax(total.t(), dim=0)))
    acc_2 = (torch.argmax(total.t(), dim=0) == ar).sum() / n
    return (nce_1 + nce_2) / 2, (acc_1 + acc_2) / 2


def forward(audio, imgs, sd: State, d: Dims, plan: MaskPlan, mae_loss_weight=1.0, contrast_loss_weight=0.01,
            gather=None):
    """CAVMAE_BASE.forward, cav_mae_base.py:685-741 (two-pass arrangement). `gather(x)->x_global` stands in for
    GatherLayer (gather_layer.py:21-37); None = world size 1."""
    zero = torch.tensor(0.0, device=audio.device)
    mask_a = mask_v = None
    if mae_loss_weight != 0:
        x, mask_a, ira, mask_v, irv, _, _ = forward_encoder(audio, imgs, sd, d, plan, 0.75, 0.75)  # :696 hard-coded
        x = block(x, sd, "mm_layer_1.", d.heads, "a")
        x = block(x, sd, "mm_layer_2.", d.heads, "a")
        pa, pv = forward_decoder(x, ira, irv, len_keep_of(d.Ta, 0.75), len_keep_of(d.Tv, 0.75), sd, d)
        loss_mae_a = mae_loss(patchify_target_audio(audio, d), pa, mask_a)
        loss_mae_v = mae_loss(patchify_target_video(imgs, d), pv, mask_v)
        loss_mae = loss_mae_a + loss_mae_v                      # NOT multiplied by the weight (:707)
    else:
        loss_mae_a = loss_mae_v = loss_mae = zero
    if contrast_loss_weight != 0:
        ca, cv = forward_encoder_mmixed(audio, imgs, sd, d, plan)
        if gather is not None:
            ca, cv = gather(ca), gather(cv)
        loss_c, c_acc = contrastive(ca.mean(dim=1), cv.mean(dim=1), bidirect=True)
        loss_c = contrast_loss_weight * loss_c
    else:
        loss_c, c_acc = zero, zero
    loss = loss_c + loss_mae
    return loss, loss_mae, loss_mae_a, loss_mae_v, loss_c, mask_a, mask_v, c_acc


def forward_single_pass(audio, imgs, sd: State, d: Dims, plan: MaskPlan, mae_loss_weight=1.0,
                        contrast_loss_weight=0.01, ratio_a=0.75, ratio_v=0.75, gather=None, bidirect=True):
    """Single-pass arrangement (recovered from cav_mae_huge.cpython-39.pyc): ONE shared
    vit_base.blocks stack over both modalities ('a'/'v' norms), one mask ratio; the same encoder output feeds the
    MAE branch (fusion blocks + decoder) and the InfoNCE branch (final-LN token means)."""
    a, v = embed_tokens(audio, imgs, sd, d)
    keep_a, keep_v = len_keep_of(d.Ta, ratio_a), len_keep_of(d.Tv, ratio_v)
    a, mask_a, ira = apply_masking(a, plan.ids_shuffle_a, keep_a)
    v, mask_v, irv = apply_masking(v, plan.ids_shuffle_v, keep_v)
    for i in range(d.depth):
        v = block(v, sd, f"vit_base.blocks.{i}.", d.heads, "v")
        a = block(a, sd, f"vit_base.blocks.{i}.", d.heads, "a")
    cv = layer_norm(v, sd["vit_base.norm.weight"], sd["vit_base.norm.bias"], LN_EPS_FINAL)
    ca = layer_norm(a, sd["vit_base.norm_a.weight"], sd["vit_base.norm_a.bias"], LN_EPS_FINAL)
    zero = torch.tensor(0.0, device=audio.device)
    if mae_loss_weight != 0:
        x = torch.cat((ca, cv), 1)
        x = block(x, sd, "mm_layer_1.", d.heads, "a")
        x = block(x, sd, "mm_layer_2.", d.heads, "a")
        pa, pv = forward_decoder(x, ira, irv, keep_a, keep_v, sd, d)
        loss_mae_a = mae_loss(patchify_target_audio(audio, d), pa, mask_a)
        loss_mae_v = mae_loss(patchify_target_video(imgs, d), pv, mask_v)
        loss_mae = mae_loss_weight * (loss_mae_a + loss_mae_v)
    else:
        loss_mae_a = loss_mae_v = loss_mae = zero
    if contrast_loss_weight != 0:
        ea, ev = ca.mean(dim=1), cv.mean(dim=1)
        if gather is not None:
            ea, ev = gather(ea), gather(ev)
        loss_c, c_acc = contrastive(ea, ev, bidirect=bidirect)
        loss_c = contrast_loss_weight * loss_c
    else:
        loss_c, c_acc = zero, zero
    return loss_c + loss_mae, loss_mae, loss_mae_a, loss_mae_v, loss_c, mask_a, mask_v, c_acc


# ---------------------------------------------------------------------------------------------------------
# Optimizer step (traintest_cavmae_base.py:64-66,139,151): torch.optim.Adam, coupled L2 weight decay
# ---------------------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step: int, lr=2e-4, beta1=0.95, beta2=0.999, eps=1e-8, weight_decay=5e-7):
    """One torch.optim.Adam update (non-amsgrad, coupled L2), restated; in-place on p, m, v."""
    g = g + weight_decay * p
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# ---------------------------------------------------------------------------------------------------------
# Finetune model CAVMAEFT_BASE (cav_mae_base.py:745-1036) — SURVEY.md §8(f) rank 1
# ---------------------------------------------------------------------------------------------------------
def ft_param_shapes(d: Dims = VIT_B, label_dim: int = 527) -> "OrderedDict[str, tuple]":
    """Unique parameters of CAVMAEFT_BASE keyed as in its state_dict (:750-820): the shared ViT with the per-modality
    LayerNorm copies, the unused my_patch_embed* copies, four LayerNorm+Linear heads and the two fusion blocks."""
    D, p = d.embed_dim, d.patch
    sh = OrderedDict()
    for k, s in _vit_shapes(d).items():
        sh[f"vit_base.{k}"] = s
    sh["my_patch_embed.proj.weight"] = (D, d.in_chans, p, p); sh["my_patch_embed.proj.bias"] = (D,)
    sh["my_patch_embed_a.proj.weight"] = (D, 1, p, p); sh["my_patch_embed_a.proj.bias"] = (D,)
    for name, din in (("mlp_head", D), ("mlp_head_a", D), ("mlp_head_mm", 2 * D), ("mlp_head_mm_v2", D)):
        sh[f"{name}.0.weight"] = (din,); sh[f"{name}.0.bias"] = (din,)
        sh[f"{name}.1.weight"] = (label_dim, din); sh[f"{name}.1.bias"] = (label_dim,)
    for name in ("mm_layer_1", "mm_layer_2"):
        for k, s in _block_shapes(D, 4 * D).items():
            sh[f"{name}.{k}"] = s
    return sh


def init_ft_state(d: Dims = VIT_B, label_dim: int = 527, seed: int = 0, skip_heads: bool = False) -> State:
    """Seeded weights for the finetune model, same scheme as init_state (every tensor from its own generator)."""
    import zlib

    sd: State = OrderedDict()
    for k, shape in ft_param_shapes(d, label_dim).items():
        if skip_heads and ".head." in k:
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(("ft:" + k).encode())) % (2**31))
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        is_ln = ("norm" in k) or (k.startswith("mlp_head") and ".0." in k)
        if is_ln and k.endswith(".weight"):
            t = 1.0 + 0.1 * t
        elif is_ln and k.endswith(".bias"):
            t = 0.05 * t
        elif k.endswith(".bias"):
            t = 0.02 * t
        elif "pos_embed" in k or "token" in k:
            t = 0.02 * t
        elif "patch_embed" in k:
            t = t * (1.0 / math.sqrt(shape[1] * shape[2] * shape[3]))
        else:
            t = t * (0.7 / math.sqrt(shape[-1]))
        sd[k] = t
    return sd


def ft_head(x, sd: State, name: str):
    """nn.Sequential(nn.LayerNorm(D), nn.Linear(D, label_dim)) — cav_mae_base.py:813-816 (LayerNorm eps 1e-5)."""
    h = layer_norm(x, sd[name + ".0.weight"], sd[name + ".0.bias"], LN_EPS_BLOCK)
    return F.linear(h, sd[name + ".1.weight"], sd[name + ".1.bias"])


def ft_encode_audio(audio, sd: State, d: Dims):
    """:830-840 — patch embed, + pos, + norm_pre_a (Identity) doubling, 12 shared blocks with the 'a' norms, norm_a."""
    a = patch_embed_audio(audio, sd["vit_base.patch_embed_a.proj.weight"], sd["vit_base.patch_embed_a.proj.bias"], d)
    a = a + sd["vit_base.pos_embed_a"]
    a = a + a
    for i in range(d.depth):
        a = block(a, sd, f"vit_base.blocks.{i}.", d.heads, "a")
    return layer_norm(a, sd["vit_base.norm_a.weight"], sd["vit_base.norm_a.bias"], LN_EPS_FINAL)


def ft_encode_video(video, sd: State, d: Dims):
    """:855-871 — video [B, T, C, H, W] -> (B T) frames, shared blocks with the 'v' norms, vit_base.norm."""
    v = video.reshape(-1, *video.shape[2:])
    v = patch_embed_video(v, sd["vit_base.patch_embed.proj.weight"], sd["vit_base.patch_embed.proj.bias"], d)
    v = v + sd["vit_base.pos_embed"][:, 1:]
    v = v + v
    for i in range(d.depth):
        v = block(v, sd, f"vit_base.blocks.{i}.", d.heads, "v")
    return layer_norm(v, sd["vit_base.norm.weight"], sd["vit_base.norm.bias"], LN_EPS_FINAL)


def forward_ft(audio, video, sd: State, d: Dims, mode: str):
    """CAVMAEFT_BASE.forward training paths (is_eval=False), cav_mae_base.py:827-1036.
       'audioonly' -> out_a [B, C];  'videoonly' -> [B, T, C] squeezed at T == 1;  'mm_grad' -> (out, out_a, out_v)
       (the reference's torch.cat((a, v), dim=1) at :1019 requires one frame per sample)."""
    if mode == "audioonly":
        a = ft_encode_audio(audio, sd, d)
        return ft_head(a.mean(dim=1), sd, "mlp_head_a")
    if mode == "videoonly":
        bs, t = video.shape[0], video.shape[1]
        v = ft_encode_video(video, sd, d)
        x = ft_head(v.mean(dim=1), sd, "mlp_head")
        return x.reshape(bs, t, -1).squeeze(1)
    if mode == "mm_grad":
        a = ft_encode_audio(audio, sd, d)
        v = ft_encode_video(video, sd, d)
        out_a = ft_head(a.mean(dim=1), sd, "mlp_head_a")
        out_v = ft_head(v.mean(dim=1), sd, "mlp_head")
        av = torch.cat((a, v), dim=1)
        av = block(av, sd, "mm_layer_1.", d.heads, "a")
        av = block(av, sd, "mm_layer_2.", d.heads, "a")
        av = torch.cat((av[:, :d.Ta].mean(dim=1), av[:, d.Ta:].mean(dim=1)), dim=-1)
        return ft_head(av, sd, "mlp_head_mm"), out_a, out_v
    raise ValueError(f"unsupported mode {mode!r}")
