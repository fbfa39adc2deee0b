"""CPU-only checks of the host side: checkpoint layout of the drop-in module, the C-ABI library's exports, the
gradient-bucket planner, and the N>1 paths (GradSync, GatherLayer, global-batch InfoNCE) on a world_size-2 gloo
group. No kernel is launched here."""
import ctypes
import dataclasses
import json
import os
import re
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avsiam_b200 import CAVMAE_BASE, Dims, _lib
from avsiam_b200.cav_mae_base import chunk_sizes, len_keep_of
from avsiam_b200.ddp import GradSync, plan_buckets
from oracle import avsiam_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tiny_model():
    return CAVMAE_BASE(dims=Dims(**dataclasses.asdict(O.TINY)))


# ------------------------------------------------------------------------------------------------ layout / boundary
def test_dropin_state_dict_layout_is_the_references(golden_dir):
    layout = json.load(open(os.path.join(golden_dir, "state_dict_layout.json")))
    model = CAVMAE_BASE(audio_length=1024, norm_pix_loss=False, modality_specific_depth=23, tr_pos=False, opt=None)
    sd = model.state_dict()
    assert set(sd) == set(layout) and len(sd) == 963
    for k, shape in layout.items():
        assert list(sd[k].shape) == shape, k
    # my_blocks.* are aliases of vit_base.blocks.* (cav_mae_base.py:278): same storage, one parameter
    assert sd["my_blocks.3.attn.qkv.weight"].data_ptr() == sd["vit_base.blocks.3.attn.qkv.weight"].data_ptr()
    assert sum(p.numel() for p in model.parameters()) == 248_035_494
    # a DDP-style checkpoint ("module." prefix, traintest_cavmae_base.py:229-234) loads through a wrapper
    wrapped = torch.nn.Module()
    wrapped.module = model
    assert all(k.startswith("module.") for k in wrapped.state_dict())


def test_ctor_matches_reference_call_site():
    # run_cavmae_pretrain_base.py:175
    m = CAVMAE_BASE(audio_length=1024, norm_pix_loss=True, modality_specific_depth=23, tr_pos=False, opt=object(),
                    dims=Dims(**dataclasses.asdict(O.TINY)))
    assert m.arrangement == "two_pass"
    a, v = torch.zeros(2, 256, 32), torch.zeros(2, 3, 64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(a, v, 0.75, 0.75, mae_loss_weight=1.0, contrast_loss_weight=0.01, mask_mode="unstructured")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "avsiam_b200.h")).read()
    declared = set(re.findall(r"\b(avs_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"avs_gemm_epilogue_t"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)          # loads without a GPU; no compute call is made
    for name in declared:
        assert hasattr(lib, name), name
    lib.avs_version.restype = ctypes.c_int
    assert lib.avs_version() >= 100


def test_host_index_math_matches_oracle():
    for B in (1, 2, 4, 5, 7, 52, 256):
        assert chunk_sizes(B) == O.chunk_sizes(B)
    for L in (512, 196, 657, 256, 32, 16):
        for i in range(5):
            assert len_keep_of(L, 0.2 * i) == O.len_keep_of(L, 0.2 * i)
        assert len_keep_of(L, 0.75) == O.len_keep_of(L, 0.75)
    assert len_keep_of(512, 0.2 * 3) == 204    # python float rounding of int(512*(1-0.6000000000000001))


def test_used_parameter_sets_match_reference_probe(golden_dir):
    """Which parameters receive gradient per pass = the reference's own sets recorded in the golden fixtures."""
    cases = torch.load(os.path.join(golden_dir, "cavmae_base_forward.pt"), weights_only=False)
    model = CAVMAE_BASE()

    class _FakeArena:
        slots = {n: None for n, _ in model.named_parameters()}
        params = dict(model.named_parameters())
    model._arena = _FakeArena()
    for c in cases:
        want = {k for k in c["grad_norm"] if ".head." not in k}
        got = set(model._used_param_names(c["mae_w"] != 0, c["c_w"] != 0))
        assert got == want, (c["name"], sorted(got ^ want)[:6])


# ------------------------------------------------------------------------------------------------ bucket planner
def test_plan_buckets_orders_and_skips_unused():
    slots, off = {}, 0
    for n, numel in (("emb.w", 1000), ("blk0.w", 5000), ("blk0.spare", 64), ("blk0.b", 64), ("big_unused", 100000),
                     ("blk1.w", 5000), ("dec.w", 3000)):
        slots[n] = (off, numel, None)
        off += (numel + 63) // 64 * 64
    used = ["emb.w", "blk0.w", "blk0.b", "blk1.w", "dec.w"]
    touches = [("dec.",), ("blk1.",), ("blk0.",), ("emb.",)]       # execution order of the reverse pass
    plan = plan_buckets(slots, used, touches, bucket_elems=6000)
    covered = set()
    for s, e, r in plan:
        assert s % 64 == 0 and e % 64 == 0 and e > s
        covered.update(range(s, e, 64))
    for n in used:
        o, numel, _ = slots[n]
        assert set(range(o, o + numel, 64)) <= covered
    o, numel, _ = slots["big_unused"]
    assert not (set(range(o, o + numel, 64)) & covered)           # never communicated
    by = {s: r for s, e, r in plan}
    assert by[slots["dec.w"][0]] == 0 and by[slots["blk1.w"][0]] == 1
    blk0 = [b for b in plan if b[0] <= slots["blk0.w"][0] < b[1]][0]
    assert blk0[1] >= slots["blk0.b"][0] + 64 and blk0[2] == 2      # the small spare tensor is bridged


# ------------------------------------------------------------------------------------------------ world_size 2 (gloo)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = globals()[fn_name](rank, world)
        # tensors go through the queue as numpy arrays (no shared-memory handles that die with the worker)
        q.put((rank, tuple(t.detach().numpy().copy() if torch.is_tensor(t) else t for t in res)))
    finally:
        dist.destroy_process_group()


def _run2(fn_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fn_name, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return {r: tuple(torch.from_numpy(t) if hasattr(t, "dtype") and hasattr(t, "shape") else t for t in res)
            for r, res in out.items()}


class _Arena:
    ALIGN = 64

    def __init__(self, rank):
        self.slots = {"a": (0, 640, None), "skip": (640, 6400000, None), "b": (6400640, 320, None)}
        g = torch.Generator().manual_seed(rank)
        self.grads = torch.randn(6400640 + 320, generator=g)


def _gradsync_case(rank, world):
    arena = _Arena(rank)
    before = arena.grads.clone()
    sync = GradSync()
    sync.begin(("k",), arena, ["a", "b"], [("b",), ("a",)])
    sync.after_closure(0)
    sync.after_closure(1)
    sync.finish()
    return before, arena.grads, sync.bytes_last_step


def test_gradsync_averages_used_ranges_only():
    out = _run2("_gradsync_case")
    (b0, g0, bytes0), (b1, g1, _) = out[0], out[1]
    avg = (b0 + b1) / 2
    for s, e in ((0, 640), (6400640, 6400960)):
        assert torch.allclose(g0[s:e], avg[s:e]) and torch.allclose(g1[s:e], avg[s:e])
    assert torch.equal(g0[640:6400640], b0[640:6400640])          # unused range untouched, not communicated
    assert bytes0 == (640 + 320) * 4


def _gather_case(rank, world):
    from avsiam_b200.gather_layer import GatherLayer, all_gather_embeddings
    g = torch.Generator().manual_seed(100 + rank)
    B, D = 3, 16
    ea = torch.randn(B, 1, D, generator=g).requires_grad_(True)
    ev = torch.randn(B, 1, D, generator=g).requires_grad_(True)
    # the reference's usage, cav_mae_base.py:724-729
    ga = torch.cat(GatherLayer.apply(ea), dim=0)
    gv = torch.cat(GatherLayer.apply(ev), dim=0)
    loss, acc = O.contrastive(ga.mean(dim=1), gv.mean(dim=1), bidirect=True)
    loss.backward()
    pa, pv = all_gather_embeddings(ea.detach()[:, 0], ev.detach()[:, 0])
    return (ea.detach(), ev.detach(), ea.grad, ev.grad, float(loss), ga.detach(), pa, pv)


def test_gather_layer_global_infonce_matches_single_process():
    out = _run2("_gather_case")
    ea = torch.cat([out[0][0], out[1][0]], 0).requires_grad_(True)
    ev = torch.cat([out[0][1], out[1][1]], 0).requires_grad_(True)
    loss, _ = O.contrastive(ea.mean(dim=1), ev.mean(dim=1), bidirect=True)
    loss.backward()
    for r in range(2):
        assert out[r][4] == pytest.approx(float(loss), rel=1e-6)
        assert torch.equal(out[r][5], ea.detach())                              # rank-major concatenation
        assert torch.equal(out[r][6], ea.detach()[:, 0]) and torch.equal(out[r][7], ev.detach()[:, 0])
        # GatherLayer.backward all-reduces W identical copies: each rank holds W x its slice of the global gradient
        # (DDP's 1/W average restores the global-loss gradient) — gather_layer.py:34-37
        assert torch.allclose(out[r][2], 2 * ea.grad[3 * r:3 * r + 3], rtol=1e-5, atol=1e-7)
        assert torch.allclose(out[r][3], 2 * ev.grad[3 * r:3 * r + 3], rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------ checkpoints
def _concat_case(rank, world):
    from avsiam_b200.evaluate import distributed_concat
    t = torch.arange(rank * 6, rank * 6 + 6, dtype=torch.float32).reshape(3, 2)
    return (distributed_concat(t, 5),)


def test_distributed_concat_gathers_in_rank_order_and_truncates():
    """traintest_cavmae_base.py:21-26: all ranks' rows, rank order, cut to the dataset size."""
    out = _run2("_concat_case")
    want = torch.arange(10, dtype=torch.float32).reshape(5, 2)
    assert torch.equal(out[0][0], want) and torch.equal(out[1][0], want)
    from avsiam_b200.evaluate import distributed_concat
    assert torch.equal(distributed_concat(want, 3), want[:3])          # no process group: single-rank identity


def test_checkpoint_roundtrip_pt_to_ft_and_weight_average(tmp_path):
    """`module.`-prefixed save/load, strict=False PT -> FT transfer (run_cavmae_ft_base.py:245-258) and
    wa_model (run_cavmae_ft_base.py:169-180) on the parameter containers (no GPU needed)."""
    import dataclasses
    from avsiam_b200 import CAVMAE_BASE, CAVMAEFT_BASE, Dims
    from avsiam_b200 import checkpoint as ck
    from oracle import avsiam_oracle as O
    d = Dims(**dataclasses.asdict(O.TINY))
    pt = CAVMAE_BASE(dims=d)
    pt.load_state_dict(O.with_aliases(O.init_state(O.TINY, seed=3)))
    path = str(tmp_path / "models" / "audio_model.1.pth")
    ck.save_model(pt, path)
    saved = torch.load(path)
    assert all(k.startswith("module.") for k in saved) and len(saved) == len(pt.state_dict())
    ft = CAVMAEFT_BASE(label_dim=7, dims=d)
    missing, unexpected = ck.load_pretrained(ft, path, create_fusion=True)
    assert all(k.startswith("module.mlp_head") for k in missing)            # only the new heads are missing
    assert any(k.startswith("module.decoder_") for k in unexpected) and any(k.startswith("module.ast_base.") for k in unexpected)
    assert torch.equal(ft.vit_base.blocks[1].attn.qkv.weight, pt.vit_base.blocks[1].attn.qkv.weight)
    assert torch.equal(ft.mm_layer_2.mlp.fc1.weight, ft.vit_base.blocks[O.TINY.depth - 1].mlp.fc1.weight)  # __create_fusion__
    # weight averaging over two epochs
    pt2 = CAVMAE_BASE(dims=d)
    pt2.load_state_dict(O.with_aliases(O.init_state(O.TINY, seed=4)))
    ck.save_model(pt2, str(tmp_path / "models" / "audio_model.2.pth"))
    avg = ck.wa_model(str(tmp_path), 1, 2)
    k = "module.vit_base.blocks.0.mlp.fc2.weight"
    want = (pt.state_dict()[k[7:]] + pt2.state_dict()[k[7:]]) / 2
    assert torch.allclose(avg[k], want, atol=1e-7)
    m3 = CAVMAE_BASE(dims=d)
    assert ck.load_model(m3, avg) == ([], [])


def test_patch_rebinds_the_names_the_reference_train_loop_resolves():
    """avsiam_b200.patch() on the UNMODIFIED reference modules (build container only): `DDP`, `torch.optim.Adam` and
    `models.CAVMAE_BASE` / `CAVMAEFT_BASE` are what train() / run_cavmae_*.py resolve at run time, everything else
    passes through, and undo() restores the originals."""
    import os
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not present (GPU box)")
    import sys
    import torch
    import avsiam_b200
    from oracle import ref_shim
    tt = ref_shim.load_traintest()
    models_pkg = sys.modules["models"]
    names = tt.train.__code__.co_names
    assert "DDP" in names and "torch" in names and "GradScaler" in names      # globals looked up by name at call time
    assert tt.train.__globals__ is tt.__dict__
    orig_ddp, orig_torch = tt.DDP, tt.torch
    undo = avsiam_b200.patch(tt, models_pkg)
    try:
        assert tt.DDP is avsiam_b200.B200DDP
        assert tt.torch.optim.Adam is avsiam_b200.FusedAdam
        assert tt.torch.optim.lr_scheduler.MultiStepLR is torch.optim.lr_scheduler.MultiStepLR
        assert tt.torch.device is torch.device and tt.torch.nn.SyncBatchNorm is torch.nn.SyncBatchNorm
        assert tt.torch.utils.data.DataLoader is torch.utils.data.DataLoader and tt.torch.save is torch.save
        assert models_pkg.CAVMAE_BASE is avsiam_b200.CAVMAE_BASE and models_pkg.CAVMAEFT_BASE is avsiam_b200.CAVMAEFT_BASE
        # the constructor call of run_cavmae_pretrain_base.py:175 is accepted as written
        m = models_pkg.CAVMAE_BASE(audio_length=1024, norm_pix_loss=False, modality_specific_depth=23, tr_pos=False, opt=None)
        assert len(m.state_dict()) == 963
    finally:
        undo()
    assert tt.DDP is orig_ddp and tt.torch is orig_torch


def test_bench_flop_model_matches_baseline_figure():
    """bench.py's per-sample training FLOP model (3 x forward: encoder over the kept tokens, two fusion blocks, decoder over
    all tokens, embeddings, prediction heads, attention 4 S^2 hd) against the figure BASELINE.md quotes for the ViT-B/16
    single-pass step (242.0 GFLOP), and its contrastive-only variant against a hand count for ViT-H/14."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    d = O.VIT_B
    got = bench.train_gflop_per_sample(d, int(d.Ta * 0.25), int(d.Tv * 0.25))
    assert abs(got - bench.TRAIN_GFLOP_PER_SAMPLE["single_pass"]) < 0.01 * got, got
    h = O.VIT_H
    ka, kv = int(h.Ta * 0.25), int(h.Tv * 0.25)
    assert (ka, kv) == (164, 64)
    D, hd = h.embed_dim, h.embed_dim // h.heads
    per_block = 2.0 * (ka + kv) * D * D * 12 + 4.0 * (ka * ka + kv * kv) * hd * h.heads
    embed = 2.0 * ka * D * 196 + 2.0 * kv * D * 588
    want = 3.0 * (h.depth * per_block + embed) / 1e9
    assert abs(bench.train_gflop_per_sample(h, ka, kv, with_mae=False) - want) < 1e-6 * want
