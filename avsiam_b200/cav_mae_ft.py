"""Drop-in `CAVMAEFT_BASE` (reference: src/models/cav_mae_base.py:745-1036) — the finetune model of
run_cavmae_ft_base.py / traintest_ft_base.py — on the same hand-written sm_100a kernels as the pretraining model.

Kept from the reference: constructor signature (:746-747), parameter names / checkpoint layout (vit_base.*,
my_blocks.* aliases, my_patch_embed*, mlp_head{,_a,_mm,_mm_v2}, mm_layer_{1,2}; `strict=False` loading of a
pretraining checkpoint works because the shared names coincide), `__create_fusion__` (:823-825) and
`forward(a, v, mode, is_eval=False)` for the modes the training / evaluation loops use:
  'audioonly' (:828-849), 'videoonly' (:852-880), 'retrieval' (:883-917), 'mm_grad' training (:983-1036) and
  evaluation (:936-980).
No masking, no decoder: full sequences (512 audio tokens, 196 tokens per frame) through the 12 shared blocks with
the per-modality LayerNorms, final norms, token means, LayerNorm+Linear heads; 'mm_grad' adds the two fusion
blocks over the concatenated 708-token sequence and the [mean_a | mean_v] head.  The returned logits are fp32 and
autograd-connected (the loops apply BCEWithLogits / CE themselves, traintest_ft_base.py:78-83).
The nn.Module tree is only a parameter container; there is no torch fallback.
"""
from __future__ import annotations

import copy
from collections import OrderedDict
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .cav_mae_base import _Dims, _PatchEmbed, _ViT
from .engine import Act, EmbedSpec, Engine, Group, ParamArena

I32, F32 = torch.int32, torch.float32


class _FtTapeFn(torch.autograd.Function):
    """forward re-emits the logits the kernels already produced; backward seeds their gradients, runs the
    hand-written reverse pass and hands per-parameter gradients back to autograd (or leaves them in the arena)."""

    @staticmethod
    def forward(ctx, holder, n_out, *tensors):
        ctx.holder = holder
        ctx.n_out = n_out
        return tuple(t.clone() for t in tensors[:n_out])

    @staticmethod
    def backward(ctx, *grads):
        h = ctx.holder
        mod = h["module"]
        arena = mod._arena
        if not mod.accumulate_into_arena:
            arena.zero_grads()
        for act, g in zip(h["outs"], grads):
            act.g = None if g is None else g.contiguous().to(F32)
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
            return (None, None) + tuple(None for _ in range(ctx.n_out + len(h["used"])))
        return (None, None) + tuple(None for _ in range(ctx.n_out)) + tuple(arena.grad(n).clone() for n in h["used"])


def _head(dim: int, label_dim: int) -> nn.Sequential:
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, label_dim))


class CAVMAEFT_BASE(nn.Module):
    """AVSiam / CAV-MAE finetune model, B200-native. See module docstring."""

    def __init__(self, label_dim, img_size=224, audio_length=1024, patch_size=16, in_chans=3,
                 embed_dim=768, modality_specific_depth=23, num_heads=16, mlp_ratio=4., norm_layer=nn.LayerNorm,
                 norm_pix_loss=False, tr_pos=True, *, dims: Optional[_Dims] = None):
        super().__init__()
        # like the reference, embed_dim / num_heads / modality_specific_depth / tr_pos are ignored (:750-767 literals)
        d = dims if dims is not None else _Dims(audio_len=audio_length, img=img_size, patch=patch_size,
                                                in_chans=in_chans)
        self.dims = d
        self.label_dim = label_dim
        D, p = d.embed_dim, d.patch
        self.vit_base = _ViT(d)
        for blk in self.vit_base.blocks:                                          # :769-774
            for n in ("norm1", "norm2"):
                getattr(blk, n + "_a").load_state_dict(getattr(blk, n).state_dict())
                getattr(blk, n + "_v").load_state_dict(getattr(blk, n).state_dict())
        self.my_blocks = self.vit_base.blocks                                     # alias (:783)
        self.my_patch_embed = _PatchEmbed(d.in_chans, D, p)                       # :792-801
        self.my_patch_embed_a = _PatchEmbed(1, D, p)
        self.my_patch_embed.load_state_dict(self.vit_base.patch_embed.state_dict())
        with torch.no_grad():
            self.my_patch_embed_a.proj.weight.copy_(self.vit_base.patch_embed.proj.weight.mean(dim=1, keepdim=True))
            self.my_patch_embed_a.proj.bias.copy_(self.vit_base.patch_embed.proj.bias)
        self.vit_base.patch_embed_a = copy.deepcopy(self.my_patch_embed_a)        # :804
        self.vit_base.pos_embed_a = nn.Parameter(                                 # :805
            F.interpolate(self.vit_base.pos_embed[:, 1:].detach().permute(0, 2, 1), size=[d.Ta]).permute(0, 2, 1)
            .contiguous())
        self.vit_base.norm_a = copy.deepcopy(self.vit_base.norm)                  # :806
        self.mlp_head = _head(D, label_dim)                                       # :813-816
        self.mlp_head_a = _head(D, label_dim)
        self.mlp_head_mm = _head(2 * D, label_dim)
        self.mlp_head_mm_v2 = _head(D, label_dim)
        self.mm_layer_1 = copy.deepcopy(self.vit_base.blocks[d.depth - 2])        # :819-820
        self.mm_layer_2 = copy.deepcopy(self.vit_base.blocks[d.depth - 1])

        self._arena: Optional[ParamArena] = None
        self._engine: Optional[Engine] = None
        self.direct_grads = False
        self.accumulate_into_arena = False
        self.process_group = None
        self.grad_sync = None
        self._used_cache = {}
        self._last_active = None
        self.register_load_state_dict_post_hook(CAVMAEFT_BASE._after_load_state_dict)
        from .optim import register_model
        register_model(self)   # FusedAdam(model.parameters(), ...) finds the arena owner from a bare parameter list

    def __create_fusion__(self):
        """Run after loading pretraining weights (:823-825): the fusion blocks restart from blocks 10 / 11."""
        d = self.dims
        self.mm_layer_1.load_state_dict(self.vit_base.blocks[d.depth - 2].state_dict())
        self.mm_layer_2.load_state_dict(self.vit_base.blocks[d.depth - 1].state_dict())
        if self._arena is not None:
            self._arena.shadow_fresh = False

    @staticmethod
    def _after_load_state_dict(module, incompatible_keys):
        if module._arena is not None:
            module._arena.shadow_fresh = False

    def _unique_named_params(self) -> "OrderedDict[str, nn.Parameter]":
        out = OrderedDict()
        for n, p in self.named_parameters():
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
            self._ensure_engine(next(self.parameters()).device)
        return self._arena

    # ------------------------------------------------------------------------------------------ parameter sets
    def _used_param_names(self, mode: str) -> List[str]:
        """Parameters that receive gradient in `mode` (matches the reference's sets in tests/golden)."""
        use_a = mode in ("audioonly", "mm_grad")
        use_v = mode in ("videoonly", "mm_grad")
        names = []
        for n in self._arena.slots:
            if n.startswith("vit_base.patch_embed_a.") or n == "vit_base.pos_embed_a" or n.startswith("vit_base.norm_a."):
                ok = use_a
            elif n.startswith("vit_base.patch_embed.") or n == "vit_base.pos_embed" or n.startswith("vit_base.norm."):
                ok = use_v
            elif n.startswith("vit_base.blocks."):
                nm = n.split(".", 3)[3].split(".")[0]
                if nm.startswith("norm"):
                    ok = (use_a and nm.endswith("_a")) or (use_v and nm.endswith("_v"))
                else:
                    ok = True
            elif n.startswith("mlp_head_a."):
                ok = use_a
            elif n.startswith("mlp_head_mm."):
                ok = mode == "mm_grad"
            elif n.startswith("mlp_head."):
                ok = use_v
            elif n.startswith("mm_layer_"):
                ok = mode == "mm_grad" and n.split(".")[1] in ("norm1_a", "norm2_a", "attn", "mlp")
            else:
                ok = False          # cls_token, head, my_patch_embed*, mlp_head_mm_v2, unused norms
            if ok and self._arena.params[n].requires_grad:
                names.append(n)
        return names

    # ------------------------------------------------------------------------------------------ forward
    @staticmethod
    def _identity_ids(n: int, T: int, dev) -> torch.Tensor:
        return torch.arange(T, dtype=I32, device=dev).unsqueeze(0).expand(n, T).contiguous()

    def forward(self, a, v, mode, is_eval=False):
        ref = a if a is not None else v
        if not ref.is_cuda:
            raise RuntimeError("avsiam_b200.CAVMAEFT_BASE runs on CUDA (sm_100a) only — there is no CPU path")
        if mode not in ("audioonly", "videoonly", "mm_grad", "retrieval"):
            raise ValueError(f"CAVMAEFT_BASE.forward: unknown mode {mode!r} (audioonly | videoonly | mm_grad | retrieval)")
        dev = ref.device
        eng = self._ensure_engine(dev)
        arena = self._arena
        d = self.dims
        want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        tape: Optional[list] = [] if want_grad else None
        arena.refresh_shadow()
        if mode == "retrieval":
            tape = None                      # feature extraction (retrieval.py runs it under no_grad)
        use_a = mode in ("audioonly", "mm_grad", "retrieval")
        use_v = mode in ("videoonly", "mm_grad", "retrieval")
        B = ref.shape[0]
        audio = a.contiguous().float() if use_a else None
        frames = None
        T = 1
        if use_v:
            T = v.shape[1]
            frames = v.reshape(B * T, *v.shape[2:]).contiguous().float()      # 'b t c w h -> (b t) c w h'
        if mode == "mm_grad" and not is_eval and T != 1:
            raise ValueError("mode 'mm_grad' (training) takes one frame per sample, like the reference's "
                             "torch.cat((a, v), dim=1) at cav_mae_base.py:1019")
        specs = []
        if use_a:
            specs.append(EmbedSpec("a", self._identity_ids(B, d.Ta, dev), d.Ta))
        if use_v:
            specs.append(EmbedSpec("v", self._identity_ids(B * T, d.Tv, dev), d.Tv))
        x, groups = eng.embed(tape, audio, frames, specs)
        for i in range(d.depth):
            x = eng.block(tape, x, groups, f"vit_base.blocks.{i}.", d.heads)
        norm_of = {"a": "vit_base.norm_a", "v": "vit_base.norm"}
        outs: List[Act] = []
        if mode == "audioonly":
            _, pooled = eng.final_norm(tape, x, groups, norm_of, cat=False, pool=True)
            outs.append(eng.head(tape, pooled[0], "mlp_head_a"))
        elif mode == "videoonly":
            _, pooled = eng.final_norm(tape, x, groups, norm_of, cat=False, pool=True)
            outs.append(eng.head(tape, pooled[0], "mlp_head"))
        elif mode == "retrieval":
            # :883-917 — normalised token features of the audio clip and of frame 5 (`return a, v[:, 5]`)
            y, _ = eng.final_norm(None, x, groups, norm_of, cat=False, pool=False)
            if T <= 5:
                raise ValueError("mode 'retrieval' returns v[:, 5]: it needs at least 6 frames per sample")
            ya = y.t[:B * d.Ta].view(B, d.Ta, -1).float()
            yv = y.t[B * d.Ta:].view(B, T, d.Tv, -1)[:, 5].float()
            return ya, yv
        elif not is_eval:
            y, pooled = eng.final_norm(tape, x, groups, norm_of, cat=True, pool=True)
            out_a = eng.head(tape, pooled[0], "mlp_head_a")
            out_v = eng.head(tape, pooled[1], "mlp_head")
            fg = [Group(0, B, d.Ta + d.Tv, "a")]                               # mm_layer_k(av, 'a') (:1020-1021)
            av = eng.block(tape, y, fg, "mm_layer_1.", d.heads)
            av = eng.block(tape, av, fg, "mm_layer_2.", d.heads)
            feats = eng.segment_means(tape, av, B, (d.Ta, d.Tv))
            outs += [eng.head(tape, feats, "mlp_head_mm"), out_a, out_v]
        else:
            # evaluation (:936-980): the audio tokens are fused with every frame in turn
            y, _ = eng.final_norm(None, x, groups, norm_of, cat=False, pool=False)
            ya = y.t[:B * d.Ta].view(B, d.Ta, -1)
            yv = y.t[B * d.Ta:].view(B, T, d.Tv, -1)
            fg = [Group(0, B, d.Ta + d.Tv, "a")]
            per_frame = []
            for t_idx in range(T):
                av = Act(torch.cat((ya, yv[:, t_idx]), dim=1).reshape(B * (d.Ta + d.Tv), -1).contiguous())
                av = eng.block(None, av, fg, "mm_layer_1.", d.heads)
                av = eng.block(None, av, fg, "mm_layer_2.", d.heads)
                feats = eng.segment_means(None, av, B, (d.Ta, d.Tv))
                per_frame.append(eng.head(None, feats, "mlp_head_mm").t.unsqueeze(1))
            return torch.hstack(per_frame)

        tensors = [o.t for o in outs]
        if want_grad:
            rg = hash(tuple(p.requires_grad for p in arena.params.values()))
            if (mode, rg) not in self._used_cache:
                names = self._used_param_names(mode)
                self._used_cache[(mode, rg)] = (names, arena.active_bitmap(names))
            used, active = self._used_cache[(mode, rg)]
            holder = {"module": self, "tape": tape, "outs": outs, "used": used, "active": active, "key": ("ft", mode)}
            tensors = list(_FtTapeFn.apply(holder, len(outs), *tensors, *[arena.params[n] for n in used]))
        if mode == "audioonly":
            out = tensors[0]
            return out.unsqueeze(1) if is_eval else out                          # :845-849
        if mode == "videoonly":
            return tensors[0].view(B, T, -1).squeeze(1)                          # :877
        return tensors[0], tensors[1], tensors[2]
