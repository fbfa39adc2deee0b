"""Step orchestration for the AVSiam pretraining hot path on top of the C-ABI kernels.

The engine owns no math: every FLOP and every byte moved on the device goes through `avsiam_b200.ops` (the C-ABI of
libavsiam_b200.so). What lives here is the order of launches, the activation buffers kept for backward, and a
hand-written reverse pass (a tape of closures) — the replacement for the autograd graph torch builds under
CAVMAE_BASE.forward in the reference (src/models/cav_mae_base.py:441-741).

Layout: activations are 2-D bf16 [tokens, D]; a token batch is a list of `Group`s (contiguous rows holding n_seq
sequences of S tokens of one modality), so audio and video tokens — and, in the mixed-ratio pass, all five chunks —
share ONE launch of every GEMM (the weights are shared); LayerNorm picks its affine set per group and attention runs
per group.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import os

import torch

from . import ops
from .ops import MAJOR_MN

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32
LN_EPS_BLOCK = 1e-5  # nn.LayerNorm default, cav_mae_base.py:116,331
LN_EPS_FINAL = 1e-6  # timm vit norm
TEMPERATURE = 0.05   # cav_mae_base.py:647


# ------------------------------------------------------------------------------------------------------------
# Parameter arena
# ------------------------------------------------------------------------------------------------------------
def _is_cold(name: str) -> bool:
    """Parameters the pretraining path never touches (SURVEY.md §3.1 [probe]); placed last in the arena."""
    if ".head." in name or "cls_token" in name or name.startswith("my_patch_embed"):
        return True
    if name.startswith("ast_base.") and not (name.startswith("ast_base.blocks.") or name.startswith("ast_base.norm_a.")):
        return True
    return False


class ParamArena:
    """All parameters in ONE contiguous fp32 buffer (+ same-shaped fp32 gradient buffer and bf16 shadow).

    `bind()` re-points every nn.Parameter's storage into the arena, so torch optimizers / state_dict / DDP keep
    working on the same memory while the kernels see flat, 256-byte-aligned pointers: the bf16 weight shadow is one
    cast launch, gradient zeroing one memset, the fused Adam one launch, the gradient all-reduce a few large buckets.
    """
    ALIGN = 64  # elements

    def __init__(self, named_params: "OrderedDict[str, torch.nn.Parameter]", device: torch.device):
        self.device = device
        names = [n for n in named_params if not _is_cold(n)] + [n for n in named_params if _is_cold(n)]
        self.slots: Dict[str, Tuple[int, int, torch.Size]] = OrderedDict()
        off = 0
        self.n_hot = 0
        for n in names:
            p = named_params[n]
            self.slots[n] = (off, p.numel(), p.shape)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            if not _is_cold(n):
                self.n_hot = off
        self.total = off
        self.flat = torch.zeros(self.total, dtype=F32, device=device)
        self.grads = torch.zeros(self.total, dtype=F32, device=device)
        self.shadow = torch.zeros(self.total, dtype=BF16, device=device)
        self.params = named_params
        self.shadow_fresh = False
        self.bind()

    def bind(self) -> None:
        with torch.no_grad():
            for n, (off, numel, shape) in self.slots.items():
                p = self.params[n]
                view = self.flat[off:off + numel].view(shape)
                if p.data.data_ptr() != view.data_ptr():
                    view.copy_(p.data)
                    p.data = view
        self.shadow_fresh = False

    def is_bound(self) -> bool:
        for n in (next(iter(self.slots)), next(reversed(self.slots))):
            off = self.slots[n][0]
            if self.params[n].data.data_ptr() != self.flat.data_ptr() + 4 * off:
                return False
        return True

    def f32(self, name: str) -> torch.Tensor:
        off, numel, shape = self.slots[name]
        return self.flat[off:off + numel].view(shape)

    def bf16(self, name: str) -> torch.Tensor:
        off, numel, shape = self.slots[name]
        return self.shadow[off:off + numel].view(shape)

    def grad(self, name: str) -> torch.Tensor:
        off, numel, shape = self.slots[name]
        return self.grads[off:off + numel].view(shape)

    def refresh_shadow(self) -> None:
        """fp32 master -> bf16 shadow for the hot prefix (skipped when the fused Adam already wrote it)."""
        if not self.shadow_fresh:
            ops.cast_f32_to_bf16(self.flat[:self.n_hot], self.shadow[:self.n_hot])

    def zero_grads(self) -> None:
        self.grads[:self.n_hot].zero_()

    def active_bitmap(self, names) -> torch.Tensor:
        """uint8 per ALIGN-element chunk of the hot prefix: 1 where the chunk belongs to one of `names`."""
        bm = torch.zeros(self.n_hot // self.ALIGN, dtype=torch.uint8)
        for n in names:
            off, numel, _ = self.slots[n]
            bm[off // self.ALIGN:(off + numel + self.ALIGN - 1) // self.ALIGN] = 1
        return bm.to(self.device)


# ------------------------------------------------------------------------------------------------------------
# Engine
# ------------------------------------------------------------------------------------------------------------
@dataclass
class Group:
    row0: int
    n_seq: int
    S: int
    mod: Optional[str]  # 'a' | 'v' | None (selects norm1{_a,_v,''})

    @property
    def rows(self) -> int:
        return self.n_seq * self.S


class Act:
    """An activation and (during backward) its gradient.

    `bias_sink`: fp32 gradient of the bias that was added when this activation was produced; a consumer whose
    backward kernel can emit column sums of the gradient it writes (LayerNorm backward) accumulates them there and
    sets `sink_done`, sparing the producer a separate column-sum pass over the gradient."""
    __slots__ = ("t", "g", "bias_sink", "sink_done")

    def __init__(self, t: torch.Tensor):
        self.t = t
        self.g: Optional[torch.Tensor] = None
        self.bias_sink: Optional[torch.Tensor] = None
        self.sink_done = False

    def take_sink(self) -> Optional[torch.Tensor]:
        if self.bias_sink is not None:
            self.sink_done = True
        return self.bias_sink


@dataclass
class EmbedSpec:
    mod: str                       # 'a' | 'v'
    ids_shuffle: torch.Tensor      # int32 [n, T] (first `keep` columns are the kept tokens, in order)
    keep: int
    sample_idx: Optional[torch.Tensor] = None  # int32 [n] input-sample index per output sample (chunk permutation)

    @property
    def n(self) -> int:
        return self.ids_shuffle.shape[0]


class Engine:
    def __init__(self, arena: ParamArena, dims):
        self.P = arena
        self.d = dims
        self.dev = arena.device
        self.fuse_fc1_bias = os.environ.get("AVS_FUSE_FC1_BIAS", "1") != "0"   # A/B switch
        # fc1's epilogue stores gelu'(pre) instead of pre, fc2's dgrad epilogue multiplies by it (A/B switch)
        self.store_gelu_grad = os.environ.get("AVS_STORE_GELU_GRAD", "1") != "0"

    # -------------------------------------------------------------------------------------------- helpers
    def _empty(self, *shape, dtype=BF16) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.dev)

    def _w2d(self, name: str) -> torch.Tensor:
        w = self.P.bf16(name)
        return w.view(w.shape[0], -1)

    def _g2d(self, name: str) -> torch.Tensor:
        g = self.P.grad(name)
        return g.view(g.shape[0], -1)

    def _linear_bwd(self, dy: torch.Tensor, x_in: torch.Tensor, wname: str, bname: str, M: int, n_out: int, k_in: int,
                    dx_out: Optional[torch.Tensor], dgelu_aux: Optional[torch.Tensor] = None, alpha: float = 1.0,
                    skip_bias: bool = False, dx_colsum: Optional[torch.Tensor] = None,
                    mul_aux: Optional[torch.Tensor] = None):
        """dgrad (optional), wgrad and bias-grad of y = x W^T + b. No transposes: see gemm_sm100.cuh.
        `dx_colsum` (fp32 [k_in]): += column sums of dx, accumulated in the dgrad GEMM's epilogue — the bias gradient of
        the Linear whose output gradient dx is (fc1, when this is fc2's dgrad with the dGELU epilogue)."""
        if dx_out is not None:
            ops.gemm(dy, self._w2d(wname), dx_out, M, k_in, n_out, b_major=MAJOR_MN, dgelu_aux=dgelu_aux,
                     colsum=dx_colsum, mul_aux=mul_aux)
        ops.gemm(dy, x_in, self._g2d(wname), n_out, k_in, M, a_major=MAJOR_MN, b_major=MAJOR_MN, accumulate=True,
                 split_k=0, alpha=alpha)
        if not skip_bias:
            ops.colsum(dy, self.P.grad(bname), M, n_out, alpha)

    # -------------------------------------------------------------------------------------------- patch embed
    def embed(self, tape: Optional[list], audio: torch.Tensor, imgs: torch.Tensor, specs: Sequence[EmbedSpec],
              vit: str = "vit_base.") -> Tuple[Act, List[Group]]:
        """PatchEmbed + pos-embed + the norm_pre doubling, fused with the kept-token gather: only kept patches are
        extracted, projected and written: x = 2*(patch W^T + b + pos[token])  (cav_mae_base.py:444-455,476-477)."""
        d = self.d
        D = d.embed_dim
        M = sum(s.n * s.keep for s in specs)
        x = Act(self._empty(M, D))
        groups: List[Group] = []
        r0 = 0
        for s in specs:
            rows = s.n * s.keep
            if s.mod == "a":
                K = d.patch * d.patch
                Kp = (K + 7) // 8 * 8   # TMA row pitch: 16-byte multiples (patch 14: 196 -> 200; pad columns are zeros)
                Ap = self._empty(rows, Kp)
                ops.patchify_audio(audio, s.ids_shuffle, s.keep, d.patch, Ap, s.sample_idx)
                wname, bname, pos = vit + "patch_embed_a.proj.weight", vit + "patch_embed_a.proj.bias", vit + "pos_embed_a"
                pos_t, gpos = self.P.f32(pos)[0], self.P.grad(pos)[0]
            else:
                K = d.patch * d.patch * d.in_chans
                Kp = (K + 7) // 8 * 8   # patch 14: 588 -> 592
                Ap = self._empty(rows, Kp)
                ops.patchify_video(imgs, s.ids_shuffle, s.keep, d.patch, Ap, s.sample_idx)
                wname, bname, pos = vit + "patch_embed.proj.weight", vit + "patch_embed.proj.bias", vit + "pos_embed"
                pos_t, gpos = self.P.f32(pos)[0, 1:], self.P.grad(pos)[0, 1:]
            rowidx = s.ids_shuffle[:, :s.keep].contiguous().view(-1)
            xs = x.t[r0:r0 + rows]
            w2d = self._w2d(wname)
            if Kp != K:   # projection weight re-pitched to the padded K (a [D, K] view has no 16-byte row pitch)
                wp = torch.zeros(D, Kp, dtype=BF16, device=self.dev)
                wp[:, :K].copy_(w2d)
                w2d = wp
            ops.gemm(Ap, w2d, xs, rows, D, Kp, bias=self.P.f32(bname), rowadd=pos_t, rowidx=rowidx, alpha=2.0)
            groups.append(Group(r0, s.n, s.keep, s.mod))
            if tape is not None:
                def bwd(r0=r0, rows=rows, Ap=Ap, wname=wname, bname=bname, gpos=gpos, rowidx=rowidx, K=K, Kp=Kp):
                    dx = x.g[r0:r0 + rows]
                    if Kp == K:
                        self._linear_bwd(dx, Ap, wname, bname, rows, D, K, None, alpha=2.0)
                    else:   # wgrad into a K-padded fp32 tile, its first K columns added to the [D, K] gradient
                        gw = torch.zeros(D, Kp, dtype=F32, device=self.dev)
                        ops.gemm(dx, Ap, gw, D, Kp, rows, a_major=MAJOR_MN, b_major=MAJOR_MN, accumulate=True,
                                 split_k=0, alpha=2.0)
                        self._g2d(wname).add_(gw[:, :K])
                        ops.colsum(dx, self.P.grad(bname), rows, D, 2.0)
                    ops.scatter_add_rows(dx, rowidx, gpos, 2.0)
                bwd.touch = (wname, bname, pos)
                tape.append(bwd)
            r0 += rows
        return x, groups

    # -------------------------------------------------------------------------------------------- transformer block
    @staticmethod
    def _ln_segments(groups):
        """Contiguous row ranges that share one affine set: consecutive groups of the same modality merge (the five
        audio chunks and the five video chunks of the mixed-ratio pass are two ranges)."""
        segs = []
        for g in groups:
            sfx = "" if g.mod is None else "_" + g.mod
            if segs and segs[-1][2] == sfx and segs[-1][1] == g.row0:
                segs[-1][1] = g.row0 + g.rows
            else:
                segs.append([g.row0, g.row0 + g.rows, sfx])
        return segs

    def _ln_fwd_groups(self, x, y, mean, rstd, groups, pfx, which, D):
        segs = self._ln_segments(groups)
        P = self.P
        if len(segs) == 2 and segs[0][0] == 0 and segs[0][1] == segs[1][0]:   # audio | video: ONE launch, two affine sets
            (_, split, s0), (_, end, s1) = segs
            ops.layernorm_fwd2(x[:end], P.f32(f"{pfx}{which}{s0}.weight"), P.f32(f"{pfx}{which}{s0}.bias"), split,
                               P.f32(f"{pfx}{which}{s1}.weight"), P.f32(f"{pfx}{which}{s1}.bias"), LN_EPS_BLOCK, y[:end],
                               mean[:end], rstd[:end], end, D)
            return
        for r0, r1, sfx in segs:
            ops.layernorm_fwd(x[r0:r1], P.f32(f"{pfx}{which}{sfx}.weight"), P.f32(f"{pfx}{which}{sfx}.bias"),
                              LN_EPS_BLOCK, y[r0:r1], mean[r0:r1], rstd[r0:r1], r1 - r0, D)

    def _ln_bwd_groups(self, dy, x, mean, rstd, dx, resid, groups, pfx, which, D, dbias=None):
        segs = self._ln_segments(groups)
        P = self.P
        if len(segs) == 2 and segs[0][0] == 0 and segs[0][1] == segs[1][0]:
            (_, split, s0), (_, end, s1) = segs
            ops.layernorm_bwd2(dy[:end], x[:end], mean[:end], rstd[:end], P.f32(f"{pfx}{which}{s0}.weight"),
                               P.grad(f"{pfx}{which}{s0}.weight"), P.grad(f"{pfx}{which}{s0}.bias"), split,
                               P.f32(f"{pfx}{which}{s1}.weight"), P.grad(f"{pfx}{which}{s1}.weight"),
                               P.grad(f"{pfx}{which}{s1}.bias"), dx[:end], end, D, resid=resid[:end], dbias=dbias)
            return
        for r0, r1, sfx in segs:
            ops.layernorm_bwd(dy[r0:r1], x[r0:r1], mean[r0:r1], rstd[r0:r1], P.f32(f"{pfx}{which}{sfx}.weight"),
                              dx[r0:r1], P.grad(f"{pfx}{which}{sfx}.weight"), P.grad(f"{pfx}{which}{sfx}.bias"),
                              r1 - r0, D, resid=resid[r0:r1], dbias=dbias)

    def block(self, tape: Optional[list], xin: Act, groups: Sequence[Group], pfx: str, heads: int) -> Act:
        """Block.forward (cav_mae_base.py:149-193): x += proj(attn(LN1_m x)); x += fc2(gelu(fc1(LN2_m x)))."""
        x = xin.t
        M, D = x.shape
        hd = D // heads
        Hid = self.P.slots[pfx + "mlp.fc1.weight"][2][0]
        ln1, mean1, rstd1 = self._empty(M, D), self._empty(M, dtype=F32), self._empty(M, dtype=F32)
        self._ln_fwd_groups(x, ln1, mean1, rstd1, groups, pfx, "norm1", D)
        qkv = self._empty(M, 3 * D)
        ops.gemm(ln1, self._w2d(pfx + "attn.qkv.weight"), qkv, M, 3 * D, D, bias=self.P.f32(pfx + "attn.qkv.bias"))
        o = self._empty(M, D)
        lses = []
        for g in groups:
            lse = self._empty(g.n_seq, heads, g.S, dtype=F32)
            ops.attention_fwd(qkv[g.row0:g.row0 + g.rows], o[g.row0:g.row0 + g.rows], lse, g.n_seq, g.S, heads, hd)
            lses.append(lse)
        x1 = self._empty(M, D)
        ops.gemm(o, self._w2d(pfx + "attn.proj.weight"), x1, M, D, D, bias=self.P.f32(pfx + "attn.proj.bias"), resid=x)
        ln2, mean2, rstd2 = self._empty(M, D), self._empty(M, dtype=F32), self._empty(M, dtype=F32)
        self._ln_fwd_groups(x1, ln2, mean2, rstd2, groups, pfx, "norm2", D)
        hpre, hact = self._empty(M, Hid), self._empty(M, Hid)
        sg = self.store_gelu_grad    # hpre then holds gelu'(fc1 output), not the fc1 output itself
        ops.gemm(ln2, self._w2d(pfx + "mlp.fc1.weight"), hact, M, Hid, D, bias=self.P.f32(pfx + "mlp.fc1.bias"),
                 gelu=True, aux_out=hpre, aux_grad=sg)
        out = Act(self._empty(M, D))
        out.bias_sink = self.P.grad(pfx + "mlp.fc2.bias")
        ops.gemm(hact, self._w2d(pfx + "mlp.fc2.weight"), out.t, M, D, Hid, bias=self.P.f32(pfx + "mlp.fc2.bias"),
                 resid=x1)
        if tape is not None:
            def bwd():
                dx2 = out.g
                dh = self._empty(M, Hid)
                fuse = self.fuse_fc1_bias
                self._linear_bwd(dx2, hact, pfx + "mlp.fc2.weight", pfx + "mlp.fc2.bias", M, D, Hid, dh,
                                 dgelu_aux=None if sg else hpre, mul_aux=hpre if sg else None,
                                 skip_bias=out.sink_done,
                                 dx_colsum=self.P.grad(pfx + "mlp.fc1.bias") if fuse else None)
                dln2 = self._empty(M, D)
                # fc1.bias gradient = column sums of dh: taken in the dGELU epilogue above instead of a pass over dh
                self._linear_bwd(dh, ln2, pfx + "mlp.fc1.weight", pfx + "mlp.fc1.bias", M, Hid, D, dln2, skip_bias=fuse)
                del dh
                dx1 = self._empty(M, D)
                # LN2 backward writes dx1 = d(x1) and its column sums = gradient of proj.bias in the same pass
                self._ln_bwd_groups(dln2, x1, mean2, rstd2, dx1, dx2, groups, pfx, "norm2", D,
                                    dbias=self.P.grad(pfx + "attn.proj.bias"))
                do = dln2  # reuse
                self._linear_bwd(dx1, o, pfx + "attn.proj.weight", pfx + "attn.proj.bias", M, D, D, do, skip_bias=True)
                dqkv = self._empty(M, 3 * D)
                for g, lse in zip(groups, lses):
                    delta = self._empty(g.n_seq, heads, g.S, dtype=F32)
                    r0, r1 = g.row0, g.row0 + g.rows
                    ops.attention_bwd(qkv[r0:r1], o[r0:r1], do[r0:r1], lse, delta, dqkv[r0:r1], g.n_seq, g.S, heads, hd,
                                      dbias=self.P.grad(pfx + "attn.qkv.bias"))   # qkv-bias gradient fused in
                dln1 = do
                self._linear_bwd(dqkv, ln1, pfx + "attn.qkv.weight", pfx + "attn.qkv.bias", M, 3 * D, D, dln1,
                                 skip_bias=True)
                dx = self._empty(M, D)
                self._ln_bwd_groups(dln1, x, mean1, rstd1, dx, dx1, groups, pfx, "norm1", D, dbias=xin.take_sink())
                xin.g = dx
                out.g = None
            bwd.touch = (pfx,)
            tape.append(bwd)
        return out

    # -------------------------------------------------------------------------------------------- final norm (+ pool)
    def final_norm(self, tape: Optional[list], xin: Act, groups: Sequence[Group], norm_of: Dict[str, str],
                   cat: bool, pool: bool) -> Tuple[Optional[Act], List[Optional[Act]]]:
        """Final LayerNorm (eps 1e-6) per group. cat=True writes the per-sample concatenation [ca | cv]
        (cav_mae_base.py:492-503) — requires all groups to have the same n_seq. pool=True also returns each group's
        token mean [n_seq, D] fp32 (.mean(dim=1), :563-566,729)."""
        x = xin.t
        M, D = x.shape
        mean, rstd = self._empty(M, dtype=F32), self._empty(M, dtype=F32)
        if cat:
            n = groups[0].n_seq
            assert all(g.n_seq == n for g in groups)
            stride = sum(g.S for g in groups)
            y = Act(self._empty(n * stride, D))
        else:
            stride = 0
            y = Act(self._empty(M, D))
        pooled: List[Optional[Act]] = []
        maps = []
        off = 0
        for g in groups:
            r0, r1 = g.row0, g.row0 + g.rows
            nm = norm_of[g.mod]
            if cat:
                ysl, ystride, yoff = y.t, stride, off
            else:
                ysl, ystride, yoff = y.t[r0:r1], 0, 0
            ops.layernorm_fwd(x[r0:r1], self.P.f32(nm + ".weight"), self.P.f32(nm + ".bias"), LN_EPS_FINAL, ysl,
                              mean[r0:r1], rstd[r0:r1], g.rows, D, seq_len=g.S, y_seq_stride=ystride, y_off=yoff)
            if pool:
                pa = Act(self._empty(g.n_seq, D, dtype=F32))
                ops.seq_mean_fwd(ysl, pa.t, g.n_seq, g.S, D, y_seq_stride=(ystride if cat else g.S), y_off=yoff)
                pooled.append(pa)
            else:
                pooled.append(None)
            maps.append((ystride, yoff))
            off += g.S
        if tape is not None:
            def bwd():
                dx = self._empty(M, D)
                sink = xin.take_sink()
                for g, pa, (ystride, yoff) in zip(groups, pooled, maps):
                    r0, r1 = g.row0, g.row0 + g.rows
                    nm = norm_of[g.mod]
                    dy = None
                    if y.g is not None:
                        dy = y.g if cat else y.g[r0:r1]
                    dpool = pa.g if (pa is not None and pa.g is not None) else None
                    if dy is None and dpool is None:
                        dx[r0:r1].zero_()
                        continue
                    ops.layernorm_bwd(dy, x[r0:r1], mean[r0:r1], rstd[r0:r1], self.P.f32(nm + ".weight"), dx[r0:r1],
                                      self.P.grad(nm + ".weight"), self.P.grad(nm + ".bias"), g.rows, D, dpool=dpool,
                                      pool_scale=1.0 / g.S, seq_len=g.S, y_seq_stride=ystride, y_off=yoff, dbias=sink)
                xin.g = dx
            bwd.touch = tuple(norm_of[g.mod] + "." for g in groups)
            tape.append(bwd)
        return y, pooled

    # -------------------------------------------------------------------------------------------- finetune heads
    def head(self, tape: Optional[list], x: Act, name: str) -> Act:
        """nn.Sequential(nn.LayerNorm(D), nn.Linear(D, label_dim)) on pooled fp32 features [B, D]
        (CAVMAEFT_BASE.mlp_head / mlp_head_a / mlp_head_mm, cav_mae_base.py:813-815)."""
        P = self.P
        logits, saved = ops.head_fwd(x.t, P.f32(name + ".0.weight"), P.f32(name + ".0.bias"), LN_EPS_BLOCK,
                                     P.f32(name + ".1.weight"), P.f32(name + ".1.bias"))
        out = Act(logits)
        if tape is not None:
            def bwd():
                if out.g is None:
                    return
                x.g = ops.head_bwd(out.g.contiguous().float(), saved, P.f32(name + ".0.weight"),
                                   P.f32(name + ".1.weight"), P.grad(name + ".1.weight"), P.grad(name + ".1.bias"),
                                   P.grad(name + ".0.weight"), P.grad(name + ".0.bias"))
            bwd.touch = (name + ".",)
            tape.append(bwd)
        return out

    def segment_means(self, tape: Optional[list], x: Act, n_seq: int, seg_lens: Sequence[int]) -> Act:
        """torch.cat((av[:, :Ta].mean(1), av[:, Ta:].mean(1)), dim=-1)  (cav_mae_base.py:1025-1028): fp32 [n_seq, k*D]
        from bf16 tokens [n_seq * sum(seg_lens), D]."""
        D = x.t.shape[1]
        stride = sum(seg_lens)
        parts, off = [], 0
        for L in seg_lens:
            m = self._empty(n_seq, D, dtype=F32)
            ops.seq_mean_fwd(x.t, m, n_seq, L, D, y_seq_stride=stride, y_off=off)
            parts.append(m)
            off += L
        out = Act(torch.cat(parts, dim=1))
        if tape is not None:
            def bwd():
                dx = self._empty(n_seq * stride, D)
                off = 0
                for j, L in enumerate(seg_lens):
                    ops.seq_mean_bwd(out.g[:, j * D:(j + 1) * D].contiguous(), dx, n_seq, L, D, stride, off)
                    off += L
                x.g = dx
            bwd.touch = ()
            tape.append(bwd)
        return out

    # -------------------------------------------------------------------------------------------- MAE branch
    def mae_branch(self, tape: Optional[list], xcat: Act, audio, imgs, B: int, keep_a: int, keep_v: int,
                   ids_restore_a, ids_restore_v, mask_a, mask_v, up_a: Optional[torch.Tensor],
                   up_v: Optional[torch.Tensor], losses: torch.Tensor):
        """mm_layer_1/2 ('a' norms on the concatenation), decoder, masked-MSE (cav_mae_base.py:699-707).
        `losses` fp32 [>=2] zero-initialised: [0] += loss_mae_a, [1] += loss_mae_v. up_a/up_v: device scalars holding
        dL/dloss_mae_{a,v}, read by the backward kernels when the tape runs."""
        d = self.d
        D, Dd = d.embed_dim, d.dec_dim
        Ta, Tv = d.Ta, d.Tv
        Sk, St = keep_a + keep_v, Ta + Tv
        g_cat = [Group(0, B, Sk, "a")]
        x = self.block(tape, xcat, g_cat, "mm_layer_1.", d.heads)
        x = self.block(tape, x, g_cat, "mm_layer_2.", d.heads)
        # decoder_embed
        e = self._empty(B * Sk, Dd)
        ops.gemm(x.t, self._w2d("decoder_embed.weight"), e, B * Sk, Dd, D, bias=self.P.f32("decoder_embed.bias"))
        xd = Act(self._empty(B * St, Dd))
        P = self.P
        ops.decoder_restore_fwd(e, ids_restore_a, ids_restore_v, P.f32("mask_token").view(-1),
                                P.f32("decoder_pos_embed_a")[0], P.f32("decoder_pos_embed_v")[0],
                                P.f32("decoder_modality_a").view(-1), P.f32("decoder_modality_v").view(-1), xd.t, B, Ta,
                                Tv, keep_a, keep_v, Dd)
        x_embed_in = x
        if tape is not None:
            def bwd_embed():
                de = self._empty(B * Sk, Dd)
                ops.decoder_restore_bwd(xd.g, ids_restore_a, ids_restore_v, de, P.grad("mask_token").view(-1),
                                        P.grad("decoder_pos_embed_a")[0], P.grad("decoder_pos_embed_v")[0],
                                        P.grad("decoder_modality_a").view(-1), P.grad("decoder_modality_v").view(-1), B,
                                        Ta, Tv, keep_a, keep_v, Dd)
                dx = self._empty(B * Sk, D)
                self._linear_bwd(de, x_embed_in.t, "decoder_embed.weight", "decoder_embed.bias", B * Sk, Dd, D, dx)
                x_embed_in.g = dx
                xd.g = None
            bwd_embed.touch = ("decoder_embed.", "mask_token", "decoder_pos_embed_", "decoder_modality_")
            tape.append(bwd_embed)
        g_dec = [Group(0, B, St, None)]
        xcur = xd
        for i in range(d.dec_depth):
            xcur = self.block(tape, xcur, g_dec, f"decoder_blocks.{i}.", d.dec_heads)
        # decoder_norm, split into the audio / video row sets (x-side row map), then the two prediction heads
        xdec = xcur
        parts = []
        for (off, T, wname, bname, inp, kind, mask, up, li) in (
                (0, Ta, "decoder_pred_a.weight", "decoder_pred_a.bias", audio, 0, mask_a, up_a, 0),
                (Ta, Tv, "decoder_pred_v.weight", "decoder_pred_v.bias", imgs, 1, mask_v, up_v, 1)):
            rows = B * T
            Pn = P.slots[wname][2][0]
            y = self._empty(rows, Dd)
            mean, rstd = self._empty(rows, dtype=F32), self._empty(rows, dtype=F32)
            ops.layernorm_fwd(xdec.t, P.f32("decoder_norm.weight"), P.f32("decoder_norm.bias"), LN_EPS_BLOCK, y, mean,
                              rstd, rows, Dd, seq_len=T, x_seq_stride=St, x_off=off)
            pred = self._empty(rows, Pn)
            ops.gemm(y, self._w2d(wname), pred, rows, Pn, Dd, bias=P.f32(bname))
            n_masked = float(B * (T - (keep_a if kind == 0 else keep_v)))
            if kind == 0:
                geom = (d.patch, 1, d.audio_len, d.mel)
            else:
                geom = (d.patch, d.in_chans, d.img, d.img)
            ops.mae_loss_fwd(pred, inp, mask, kind, B, *geom, n_masked, losses[li:li + 1])
            parts.append((off, T, wname, bname, inp, kind, mask, up, y, mean, rstd, pred, Pn, n_masked, geom))
        if tape is not None:
            def bwd_heads():
                dxdec = self._empty(B * St, Dd)
                sink = xdec.take_sink()
                for (off, T, wname, bname, inp, kind, mask, up, y, mean, rstd, pred, Pn, n_masked, geom) in parts:
                    rows = B * T
                    dpred = self._empty(rows, Pn)
                    ops.mae_loss_bwd(pred, inp, mask, kind, B, *geom, n_masked, up, dpred)
                    dy = self._empty(rows, Dd)
                    self._linear_bwd(dpred, y, wname, bname, rows, Pn, Dd, dy)
                    ops.layernorm_bwd(dy, xdec.t, mean, rstd, P.f32("decoder_norm.weight"), dxdec,
                                      P.grad("decoder_norm.weight"), P.grad("decoder_norm.bias"), rows, Dd, seq_len=T,
                                      x_seq_stride=St, x_off=off, dbias=sink)
                xdec.g = dxdec
            bwd_heads.touch = ("decoder_pred_", "decoder_norm.")
            tape.append(bwd_heads)

    # -------------------------------------------------------------------------------------------- contrastive branch
    def contrastive(self, tape: Optional[list], ea: Act, ev: Act, weight: float, bidirect: bool,
                    up_c: Optional[torch.Tensor], out: torch.Tensor, gather: Optional[Callable] = None, rank: int = 0,
                    world: int = 1):
        """InfoNCE on the (optionally all-gathered) pooled embeddings; out[0] = un-weighted loss, out[1] = accuracy.
        Backward returns this rank's slice of dL/d(embeddings) x world (GatherLayer.backward all-reduces W identical
        copies, gather_layer.py:34-37; DDP's later 1/W average restores the global-loss gradient)."""
        B, D = ea.t.shape
        if world > 1:
            ga, gv = gather(ea.t, ev.t)
        else:
            ga, gv = ea.t, ev.t
        N = ga.shape[0]
        ws = ops.infonce_workspace(N, D, self.dev)
        ops.infonce_fwd(ga, gv, TEMPERATURE, bidirect, ws, out[0:1], out[1:2])
        if tape is not None:
            def bwd():
                dea, dev_ = self._empty(B, D, dtype=F32), self._empty(B, D, dtype=F32)
                scratch = self._empty(2 * B * D, dtype=F32)
                ops.infonce_bwd(N, D, TEMPERATURE, bidirect, weight * world, up_c, ws, rank * B, B, scratch, dea, dev_)
                ea.g, ev.g = dea, dev_
            bwd.touch = ()
            tape.append(bwd)
