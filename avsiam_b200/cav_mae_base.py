"""Drop-in `CAVMAE_BASE` (reference: src/models/cav_mae_base.py:216-741) whose forward/backward run on the
hand-written sm_100a kernels of libavsiam_b200.so.

Kept byte-for-byte from the reference: the constructor signature (cav_mae_base.py:219-222), the
`forward(audio, imgs, mask_ratio_a, mask_ratio_v, mae_loss_weight, contrast_loss_weight, mask_mode)` -> 8-tuple
contract (:685-741), and the 963-key checkpoint layout including the `my_blocks.*` aliases (:278) and the unused
copies (`my_patch_embed*`, `ast_base.patch_embed*`, heads ...).  The nn.Module tree below is ONLY a parameter
container with the reference's names — none of its submodules is ever called; there is no torch fallback.

Two arrangements of the same kernels:
  * "two_pass" (default, literal CAVMAE_BASE): MAE branch = forward_encoder (:441-504; audio through the
    `ast_base` copy, video through `vit_base` 'v' norms, ratios hard-coded 0.75 :696); contrastive branch =
    forward_encoder_mmixed (:508-594; five chunks at ratios 0.2*i through the shared `vit_base.blocks`).
  * "single_pass" (the north-star step): one shared-encoder pass feeds both losses.
"""
from __future__ import annotations

import copy
import random
from collections import OrderedDict
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .engine import Act, EmbedSpec, Engine, Group, ParamArena
from .gather_layer import PendingGather, all_gather_embeddings

I32, F32 = torch.int32, torch.float32


# ------------------------------------------------------------------------------------------------------------
# Parameter containers (names = the reference's / timm 0.9.5's; never called)
# ------------------------------------------------------------------------------------------------------------
class _Attn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    """Parameter set of Block (cav_mae_base.py:104-145): three LayerNorm sets per sub-layer."""

    def __init__(self, dim, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.norm1_a = nn.LayerNorm(dim)
        self.norm1_v = nn.LayerNorm(dim)
        self.attn = _Attn(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm2_a = nn.LayerNorm(dim)
        self.norm2_v = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _PatchEmbed(nn.Module):
    def __init__(self, in_chans, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)


class _ViT(nn.Module):
    """Parameter set of the surgically modified timm ViT (cav_mae_base.py:236-300)."""

    def __init__(self, d):
        super().__init__()
        D = d.embed_dim
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        self.pos_embed = nn.Parameter(torch.randn(1, d.Tv + 1, D) * 0.02)
        self.patch_embed = _PatchEmbed(d.in_chans, D, d.patch)
        self.blocks = nn.Sequential(*[_Block(D) for _ in range(d.depth)])
        self.norm = nn.LayerNorm(D, eps=1e-6)
        self.head = nn.Linear(D, d.head_classes)


class _Dims:
    """Geometry; defaults are CAVMAE_BASE's literals (cav_mae_base.py:249-261,316-329)."""

    def __init__(self, embed_dim=768, depth=12, heads=12, dec_dim=512, dec_depth=8, dec_heads=16, patch=16,
                 audio_len=1024, mel=128, img=224, in_chans=3, head_classes=21843):
        self.embed_dim, self.depth, self.heads = embed_dim, depth, heads
        self.dec_dim, self.dec_depth, self.dec_heads = dec_dim, dec_depth, dec_heads
        self.patch, self.audio_len, self.mel, self.img, self.in_chans = patch, audio_len, mel, img, in_chans
        self.head_classes = head_classes
        self.fa, self.ta = mel // patch, audio_len // patch
        self.Ta, self.Tv = self.fa * self.ta, (img // patch) ** 2


# BASELINE config 5 geometries (recovered from src/models/__pycache__/cav_mae_{large,huge}.*.pyc;
# models/__init__.py:8-17 imports both). ViT-H/14: head_dim 80, 73 x 9 = 657 audio tokens (the stride-14 conv reads
# 1022 x 126 of the 1024 x 128 fbank), 16 x 16 = 256 video tokens; as shipped it builds no decoder — contrastive-only,
# single pass, single-direction InfoNCE (construct with arrangement="single_pass", bidirect_contrast=False and call
# with mae_loss_weight=0).
VIT_L_DIMS = dict(embed_dim=1024, depth=24, heads=16, dec_depth=6)
VIT_H_DIMS = dict(embed_dim=1280, depth=32, heads=16, patch=14)


def len_keep_of(L: int, ratio: float) -> int:
    return int(L * (1 - ratio))  # cav_mae_base.py:372,399


def chunk_sizes(B: int, n: int = 5) -> List[int]:
    """Sizes produced by torch.chunk(perm, 5) (cav_mae_base.py:534)."""
    cs = -(-B // n)
    out, left = [], B
    while left > 0:
        out.append(min(cs, left))
        left -= out[-1]
    return out


class _TapeFn(torch.autograd.Function):
    """Connects the engine's hand-written reverse pass to torch autograd: forward re-emits the already computed
    scalars; backward seeds the upstream-gradient scalars, runs the tape, and hands back per-parameter grads."""

    @staticmethod
    def forward(ctx, holder, la, lv, lc, *params):
        ctx.holder = holder
        return la.clone(), lv.clone(), lc.clone()

    @staticmethod
    def backward(ctx, ga, gv, gc):
        h = ctx.holder
        mod = h["module"]
        up = h["up"]
        zero = torch.zeros((), device=up.device)
        up[0:1].copy_((ga if ga is not None else zero).reshape(1).to(F32))
        up[1:2].copy_((gv if gv is not None else zero).reshape(1).to(F32))
        up[2:3].copy_((gc if gc is not None else zero).reshape(1).to(F32))
        arena = mod._arena
        if not h["tape"]:
            raise RuntimeError("avsiam_b200: backward through this forward a second time — the activation tape is freed "
                               "after its first backward (retain_graph is not supported); call forward again")
        if not mod.accumulate_into_arena:
            arena.zero_grads()
        order = list(reversed(h["tape"]))
        sync = mod.grad_sync
        if sync is not None and sync.world > 1:
            sync.begin(h["key"], arena, h["used"], [getattr(fn, "touch", ()) for fn in order])
            for i, fn in enumerate(order):
                fn()
                sync.after_closure(i)
            sync.finish()
        else:
            for fn in order:
                fn()
        h["tape"].clear()
        mod._last_active = h["active"]
        if mod.direct_grads:
            return (None, None, None, None) + tuple(None for _ in h["used"])
        return (None, None, None, None) + tuple(arena.grad(n).clone() for n in h["used"])


class CAVMAE_BASE(nn.Module):
    """CAV-MAE / AVSiam pretraining model, B200-native. See module docstring."""

    def __init__(self, img_size=224, audio_length=1024, patch_size=16, in_chans=3,
                 embed_dim=768, modality_specific_depth=23, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16,
                 mlp_ratio=4., norm_layer=nn.LayerNorm, norm_pix_loss=False, tr_pos=False, opt=None,
                 *, dims: Optional[_Dims] = None, arrangement: str = "two_pass", bidirect_contrast: bool = True):
        super().__init__()
        # The reference ignores embed_dim / num_heads / decoder_* / modality_specific_depth / tr_pos / norm_pix_loss
        # (dims are literals at cav_mae_base.py:249-261,316-329; norm_pix_loss is commented out at :673-676).
        d = dims if dims is not None else _Dims(audio_len=audio_length, img=img_size, patch=patch_size,
                                                in_chans=in_chans)
        assert arrangement in ("two_pass", "single_pass")
        self.dims = d
        self.opt = opt
        self.arrangement = arrangement
        self.bidirect_contrast = bidirect_contrast
        D, Dd, p = d.embed_dim, d.dec_dim, d.patch

        self.vit_base = _ViT(d)
        # cav_mae_base.py:264-269 norm{1,2}_{a,v} start as copies of norm{1,2}
        for blk in self.vit_base.blocks:
            for n in ("norm1", "norm2"):
                getattr(blk, n + "_a").load_state_dict(getattr(blk, n).state_dict())
                getattr(blk, n + "_v").load_state_dict(getattr(blk, n).state_dict())
        self.my_blocks = self.vit_base.blocks                                   # alias (:278)
        self.my_patch_embed = _PatchEmbed(d.in_chans, D, p)                      # :285-288
        self.my_patch_embed_a = _PatchEmbed(1, D, p)
        self.my_patch_embed.load_state_dict(self.vit_base.patch_embed.state_dict())
        with torch.no_grad():                                                    # :291-294 channel-mean init
            self.my_patch_embed_a.proj.weight.copy_(self.vit_base.patch_embed.proj.weight.mean(dim=1, keepdim=True))
            self.my_patch_embed_a.proj.bias.copy_(self.vit_base.patch_embed.proj.bias)
        self.vit_base.patch_embed_a = copy.deepcopy(self.my_patch_embed_a)       # :297
        self.vit_base.pos_embed_a = nn.Parameter(                                # :298 nearest interpolation
            F.interpolate(self.vit_base.pos_embed[:, 1:].detach().permute(0, 2, 1), size=[d.Ta]).permute(0, 2, 1)
            .contiguous())
        self.vit_base.norm_a = copy.deepcopy(self.vit_base.norm)                 # :299
        self.ast_base = copy.deepcopy(self.vit_base)                             # :303
        self.mm_layer_1 = copy.deepcopy(self.vit_base.blocks[d.depth - 1])       # :306-307
        self.mm_layer_2 = copy.deepcopy(self.vit_base.blocks[d.depth - 1])
        self.decoder_embed = nn.Linear(D, Dd, bias=True)                         # :311-314
        self.decoder_pos_embed_a = nn.Parameter(torch.zeros(1, d.Ta, Dd))
        self.decoder_pos_embed_v = nn.Parameter(torch.zeros(1, d.Tv, Dd))
        self.mask_token = nn.Parameter(torch.zeros(1, 1, Dd))
        self.decoder_blocks = nn.Sequential(*[_Block(Dd) for _ in range(d.dec_depth)])  # :316-329
        self.decoder_norm = nn.LayerNorm(Dd)
        self.decoder_pred_a = nn.Linear(Dd, p * p, bias=True)                    # :334-337
        self.decoder_pred_v = nn.Linear(Dd, p * p * d.in_chans, bias=True)
        self.decoder_modality_a = nn.Parameter(torch.zeros(1, 1, Dd))
        self.decoder_modality_v = nn.Parameter(torch.zeros(1, 1, Dd))

        # runtime state (not part of the checkpoint)
        self._arena: Optional[ParamArena] = None
        self._engine: Optional[Engine] = None
        self.mask_plan = None            # inject supplied indices (oracle.MaskPlan-like) for parity runs
        self.last_mask_plan = None
        self.direct_grads = False        # True: gradients stay in the arena (FusedAdam / B200DDP route)
        self.accumulate_into_arena = False
        self.process_group = None        # None => default group when torch.distributed is initialised
        self.grad_sync = None            # ddp.GradSync, installed by B200DDP
        self._used_cache = {}            # (arrangement, do_mae, do_c) -> (names, active-chunk bitmap)
        self._last_active = None         # bitmap of the parameters the last backward wrote (FusedAdam skips the rest)
        self.register_load_state_dict_post_hook(CAVMAE_BASE._after_load_state_dict)
        from .optim import register_model
        register_model(self)   # FusedAdam(model.parameters(), ...) finds the arena owner from a bare parameter list

    @staticmethod
    def _after_load_state_dict(module, incompatible_keys):
        if module._arena is not None:      # weights changed under the bf16 shadow
            module._arena.shadow_fresh = False

    # -------------------------------------------------------------------------------------------- arena plumbing
    def _unique_named_params(self) -> "OrderedDict[str, nn.Parameter]":
        out = OrderedDict()
        for n, p in self.named_parameters():  # de-duplicated: `vit_base.blocks.*` wins over the `my_blocks.*` alias
            out[n] = p
        return out

    def _ensure_engine(self, device: torch.device) -> Engine:
        if self._arena is None or self._arena.device != device:
            self._arena = ParamArena(self._unique_named_params(), device)
            self._engine = Engine(self._arena, self.dims)
            self._used_cache = {}
            from .optim import register_model
            register_model(self)
        elif not self._arena.is_bound():
            self._arena.bind()
        return self._engine

    @property
    def arena(self) -> ParamArena:
        if self._arena is None:
            dev = next(self.parameters()).device
            self._ensure_engine(dev)
        return self._arena

    # -------------------------------------------------------------------------------------------- masking (host part)
    def _ids_from_noise(self, noise: torch.Tensor, keep: int):
        return ops.mask_argsort(noise.contiguous(), keep)

    def _unstructured_ids(self, n: int, L: int, ratio: float, dev, supplied=None):
        """random_masking_unstructured (cav_mae_base.py:365-390): noise = rand; argsort on device."""
        keep = len_keep_of(L, ratio)
        if supplied is not None:
            return ops.mask_from_ids(supplied.to(dev, I32).contiguous(), keep) + (keep,)
        return self._ids_from_noise(torch.rand(n, L, device=dev), keep) + (keep,)

    def _structured_ids(self, n: int, L: int, ratio: float, dev, supplied=None, mode: str = "tf", noise=None):
        """random_masking_structured (cav_mae_base.py:392-439), modes 'time' / 'freq' / 'tf': random.sample picks
        int(t*r) time columns ('time'), int(f*r) frequency rows ('freq') or int(t*r*0.7) columns then int(f*r*0.7) rows
        ('tf') of the [f, t] patch grid per sample; their noise is forced to 1.1 ("large value will be removed").
        The draws use Python's `random` in the reference's exact order (all samples' columns, then all samples' rows), so
        a run seeded with random.seed(s) removes the same columns / rows as the reference; the N x k slice writes become
        one launch (ops.mask_force_noise), then the stable device argsort. `noise` (fp32 [n, L]) may be supplied."""
        keep = len_keep_of(L, ratio)
        if supplied is not None:
            return ops.mask_from_ids(supplied.to(dev, I32).contiguous(), keep) + (keep,)
        d = self.dims
        t, f = d.ta, d.fa
        cols = rows = None
        if mode == "time":
            cols = [random.sample(range(t), int(t * ratio)) for _ in range(n)]
        elif mode == "freq":
            rows = [random.sample(range(f), int(f * ratio)) for _ in range(n)]
        elif mode == "tf":
            cols = [random.sample(range(t), int(t * ratio * 0.7)) for _ in range(n)]
            rows = [random.sample(range(f), int(f * ratio * 0.7)) for _ in range(n)]
        else:
            raise ValueError(f"mask_mode {mode!r}: expected 'unstructured', 'time', 'freq' or 'tf'")
        noise = torch.rand(n, L, device=dev) if noise is None else noise.to(dev, F32).contiguous().clone()
        as_dev = lambda lst: torch.tensor(lst, dtype=I32).to(dev) if lst and len(lst[0]) else None   # [n, k]
        ops.mask_force_noise(noise, f, t, as_dev(cols), as_dev(rows), 1.1)
        return self._ids_from_noise(noise, keep) + (keep,)

    # -------------------------------------------------------------------------------------------- forward
    def forward(self, audio, imgs, mask_ratio_a=0.75, mask_ratio_v=0.75, mae_loss_weight=1., contrast_loss_weight=0.01,
                mask_mode='unstructured'):
        if not audio.is_cuda:
            raise RuntimeError("avsiam_b200.CAVMAE_BASE runs on CUDA (sm_100a) only — there is no CPU path")
        dev = audio.device
        eng = self._ensure_engine(dev)
        arena = self._arena
        audio = audio.contiguous().float()
        imgs = imgs.contiguous().float()
        B = audio.shape[0]
        d = self.dims
        want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        tape: Optional[list] = [] if want_grad else None
        arena.refresh_shadow()
        plan = self.mask_plan
        dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
        world = torch.distributed.get_world_size(self.process_group) if dist_on else 1
        rank = torch.distributed.get_rank(self.process_group) if dist_on else 0
        gather = (lambda a, v: all_gather_embeddings(a, v, self.process_group)) if world > 1 else None

        losses = torch.zeros(4, dtype=F32, device=dev)   # [mae_a, mae_v, nce, acc]
        up = torch.zeros(3, dtype=F32, device=dev)       # upstream dL/d{mae_a, mae_v, nce}, filled at backward
        mask_a = mask_v = None
        used_before = set()
        do_mae, do_c = mae_loss_weight != 0, contrast_loss_weight != 0
        if do_mae and (d.audio_len % d.patch or d.mel % d.patch or d.img % d.patch or (d.patch * d.patch) % 8):
            # the reference's patchify is a reshape into whole patches (cav_mae_base.py:343-351): it has no MAE target for
            # a patch size that does not divide the input, and CAVMAE_HUGE (patch 14) builds no decoder at all
            raise RuntimeError(f"avsiam_b200: the MAE branch needs a patch size dividing the inputs (patch {d.patch}, fbank "
                               f"{d.audio_len}x{d.mel}, frame {d.img}); ViT-H/14 runs contrastive-only: mae_loss_weight=0")

        def run_mae(xcat, ira, irv, ma, mv, ka, kv):
            eng.mae_branch(tape, xcat, audio, imgs, B, ka, kv, ira, irv, ma, mv, up[0:1], up[1:2], losses)

        if self.arrangement == "single_pass":
            ids_a, ira, mask_a, ka = self._unstructured_ids(B, d.Ta, mask_ratio_a, dev, getattr(plan, "ids_shuffle_a", None)) \
                if mask_mode == 'unstructured' else self._structured_ids(B, d.Ta, mask_ratio_a, dev, getattr(plan, "ids_shuffle_a", None), mode=mask_mode)
            ids_v, irv, mask_v, kv = self._unstructured_ids(B, d.Tv, mask_ratio_v, dev, getattr(plan, "ids_shuffle_v", None))
            x, groups = eng.embed(tape, audio, imgs, [EmbedSpec("a", ids_a, ka), EmbedSpec("v", ids_v, kv)])
            for i in range(d.depth):
                x = eng.block(tape, x, groups, f"vit_base.blocks.{i}.", d.heads)
            xcat, pooled = eng.final_norm(tape, x, groups, {"a": "vit_base.norm_a", "v": "vit_base.norm"}, cat=True,
                                          pool=do_c)
            if do_c and world > 1:
                # GatherLayer's all-gather (cav_mae_base.py:724-725) starts here, under the whole MAE branch
                pending = PendingGather(pooled[0].t, pooled[1].t, self.process_group)
                gather = lambda a, v: pending.wait()
            if do_mae:
                run_mae(xcat, ira, irv, mask_a, mask_v, ka, kv)
            if do_c:
                eng.contrastive(tape, pooled[0], pooled[1], 1.0, self.bidirect_contrast, up[2:3], losses[2:4], gather,
                                rank, world)
        else:
            if do_mae:
                # forward_encoder (:441-504): ratios hard-coded to 0.75 by the caller at :696
                ids_a, ira, mask_a, ka = self._unstructured_ids(B, d.Ta, 0.75, dev, getattr(plan, "ids_shuffle_a", None))
                ids_v, irv, mask_v, kv = self._unstructured_ids(B, d.Tv, 0.75, dev, getattr(plan, "ids_shuffle_v", None))
                xa, ga = eng.embed(tape, audio, imgs, [EmbedSpec("a", ids_a, ka)])
                xv, gv = eng.embed(tape, audio, imgs, [EmbedSpec("v", ids_v, kv)])
                ga_none = [Group(0, B, ka, None)]       # ast_base.blocks[i](a): modality=None norms (:489)
                for i in range(d.depth):
                    xv = eng.block(tape, xv, gv, f"vit_base.blocks.{i}.", d.heads)
                    xa = eng.block(tape, xa, ga_none, f"ast_base.blocks.{i}.", d.heads)
                # concatenate [ca | cv] per sample through the LN row map: two launches into one buffer
                xcat = self._final_norm_two_inputs(eng, tape, xa, xv, B, ka, kv)
                run_mae(xcat, ira, irv, mask_a, mask_v, ka, kv)
            if do_c:
                ea, ev = self._mmixed(eng, tape, audio, imgs, B, dev, plan)
                eng.contrastive(tape, ea, ev, 1.0, True, up[2:3], losses[2:4], gather, rank, world)

        la, lv, lc_raw, acc = losses[0], losses[1], losses[2], losses[3]
        if want_grad:
            # the set of gradient-receiving parameters also depends on requires_grad (freeze / unfreeze schedules)
            rg = hash(tuple(p.requires_grad for p in arena.params.values()))
            key = (self.arrangement, do_mae, do_c, rg)
            if key not in self._used_cache:
                names = self._used_param_names(do_mae, do_c)
                self._used_cache[key] = (names, arena.active_bitmap(names))
            used, active = self._used_cache[key]
            self._last_used = used
            holder = {"module": self, "tape": tape, "up": up, "used": used, "active": active, "key": key}
            la, lv, lc_raw = _TapeFn.apply(holder, la, lv, lc_raw, *[arena.params[n] for n in used])
        acc = acc.detach().clone()
        zero = torch.zeros((), device=dev)
        if do_mae:
            loss_mae_a, loss_mae_v = la, lv
            loss_mae = loss_mae_a + loss_mae_v
            if self.arrangement == "single_pass":
                loss_mae = mae_loss_weight * loss_mae       # [pyc] CAVMAE_HUGE.forward
        else:
            loss_mae_a, loss_mae_v, loss_mae = zero, zero.clone(), zero.clone()
        if do_c:
            loss_c = contrast_loss_weight * lc_raw          # cav_mae_base.py:735
            c_acc = acc
        else:
            loss_c, c_acc = zero.clone(), zero.clone()
        loss = loss_c + loss_mae                            # :739 (loss_mae NOT weighted in the two-pass model, :707)
        return loss, loss_mae, loss_mae_a, loss_mae_v, loss_c, mask_a, mask_v, c_acc

    # -------------------------------------------------------------------------------------------- pieces
    def _final_norm_two_inputs(self, eng: Engine, tape, xa: Act, xv: Act, B, ka, kv) -> Act:
        """cav_mae_base.py:492-503 for the two-pass MAE branch: cv = vit_base.norm(v), ca = ast_base.norm_a(a),
        x = cat((ca, cv), dim=1) — the audio and video tokens live in two separate buffers here."""
        P = eng.P
        D = self.dims.embed_dim
        S = ka + kv
        y = Act(torch.empty(B * S, D, dtype=torch.bfloat16, device=xa.t.device))
        stats = []
        for (xin, nm, keep, off) in ((xa, "ast_base.norm_a", ka, 0), (xv, "vit_base.norm", kv, ka)):
            rows = B * keep
            mean = torch.empty(rows, dtype=F32, device=y.t.device)
            rstd = torch.empty_like(mean)
            ops.layernorm_fwd(xin.t, P.f32(nm + ".weight"), P.f32(nm + ".bias"), 1e-6, y.t, mean, rstd, rows, D,
                              seq_len=keep, y_seq_stride=S, y_off=off)
            stats.append((xin, nm, keep, off, mean, rstd))
        if tape is not None:
            def bwd():
                for (xin, nm, keep, off, mean, rstd) in stats:
                    rows = B * keep
                    dx = torch.empty_like(xin.t)
                    ops.layernorm_bwd(y.g, xin.t, mean, rstd, P.f32(nm + ".weight"), dx, P.grad(nm + ".weight"),
                                      P.grad(nm + ".bias"), rows, D, seq_len=keep, y_seq_stride=S, y_off=off,
                                      dbias=xin.take_sink())
                    xin.g = dx
            bwd.touch = ("ast_base.norm_a.", "vit_base.norm.")
            tape.append(bwd)
        return y

    def _mmixed(self, eng: Engine, tape, audio, imgs, B, dev, plan):
        """forward_encoder_mmixed (cav_mae_base.py:508-594): 5 chunks, mask ratio 0.2*i, audio structured 'tf', video
        unstructured, all through the shared vit_base.blocks; final LN + token mean; restore batch order. All chunks of
        both modalities are ONE token batch, so each GEMM of a layer is a single launch."""
        d = self.dims
        sizes = chunk_sizes(B)
        perm_a = (plan.perm_a if plan is not None and plan.perm_a is not None else torch.randperm(B)).to(dev)
        perm_v = (plan.perm_v if plan is not None and plan.perm_v is not None else torch.randperm(B)).to(dev)
        idx_a = torch.split(perm_a.to(I32), sizes)
        idx_v = torch.split(perm_v.to(I32), sizes)
        specs = []
        for i, n in enumerate(sizes):
            r = 0 + 0.2 * i
            sup = plan.chunk_ids_a[i] if plan is not None and plan.chunk_ids_a else None
            ids, _, _, keep = self._structured_ids(n, d.Ta, r, dev, sup)
            specs.append(EmbedSpec("a", ids, keep, idx_a[i].contiguous()))
        for i, n in enumerate(sizes):
            r = 0 + 0.2 * i
            sup = plan.chunk_ids_v[i] if plan is not None and plan.chunk_ids_v else None
            ids, _, _, keep = self._unstructured_ids(n, d.Tv, r, dev, sup)
            specs.append(EmbedSpec("v", ids, keep, idx_v[i].contiguous()))
        x, groups = eng.embed(tape, audio, imgs, specs)
        for l in range(d.depth):
            x = eng.block(tape, x, groups, f"vit_base.blocks.{l}.", d.heads)
        _, pooled = eng.final_norm(tape, x, groups, {"a": "vit_base.norm_a", "v": "vit_base.norm"}, cat=False,
                                   pool=True)
        nch = len(sizes)
        D = d.embed_dim
        outs = []
        for pl, perm in ((pooled[:nch], perm_a), (pooled[nch:], perm_v)):
            chunked = torch.cat([p.t for p in pl], 0)                               # chunk order
            inv = torch.empty(B, dtype=I32, device=dev)
            inv[perm.long()] = torch.arange(B, dtype=I32, device=dev)               # :584-586 without the Python loop
            e = Act(ops.gather_rows(chunked.view(1, B, D), inv.view(1, B), B).view(B, D))   # :589-590
            outs.append((e, pl, perm.to(I32).contiguous()))
        if tape is not None:
            def bwd():
                for e, pl, perm32 in outs:
                    dch = ops.gather_rows(e.g.view(1, B, D), perm32.view(1, B), B).view(B, D)
                    r0 = 0
                    for p in pl:
                        n = p.t.shape[0]
                        p.g = dch[r0:r0 + n]
                        r0 += n
            bwd.touch = ()
            tape.append(bwd)
        return outs[0][0], outs[1][0]

    def last_used_names(self) -> List[str]:
        """Names of the parameters that received a gradient in the most recent forward that recorded a tape."""
        return list(self._last_used) if getattr(self, "_last_used", None) is not None else []

    def _used_param_names(self, do_mae: bool, do_c: bool) -> List[str]:
        """Parameters that receive gradient in this call — the reference's sets, SURVEY.md §3.1 [probe]."""
        names = []
        single = self.arrangement == "single_pass"
        for n in self._arena.slots:
            if n.startswith("vit_base.patch_embed") or n.startswith("vit_base.pos_embed"):
                ok = True
            elif n.startswith("vit_base.blocks."):
                rest = n.split(".", 3)[3]
                if rest.startswith("norm"):
                    nm = rest.split(".")[0]
                    if single:
                        ok = nm.endswith("_a") or nm.endswith("_v")
                    else:
                        ok = (do_c and (nm.endswith("_a") or nm.endswith("_v"))) or (do_mae and nm.endswith("_v"))
                else:
                    ok = True
            elif n.startswith("vit_base.norm_a."):
                ok = single or do_c
            elif n.startswith("vit_base.norm."):
                ok = True
            elif n.startswith("ast_base.blocks."):
                rest = n.split(".", 3)[3]
                nm = rest.split(".")[0]
                ok = (not single) and do_mae and nm in ("norm1", "norm2", "attn", "mlp")
            elif n.startswith("ast_base.norm_a."):
                ok = (not single) and do_mae
            elif n.startswith("mm_layer_"):
                nm = n.split(".")[1]
                ok = do_mae and nm in ("norm1_a", "norm2_a", "attn", "mlp")
            elif n.startswith("decoder_blocks."):
                nm = n.split(".", 3)[2]
                ok = do_mae and nm in ("norm1", "norm2", "attn", "mlp")   # decoder blocks run with modality=None
            elif n.startswith("decoder_") or n == "mask_token":
                ok = do_mae
            else:
                ok = False
            if ok and self._arena.params[n].requires_grad:
                names.append(n)
        return names
