#!/usr/bin/env python
"""Benchmark of the AVSiam ViT-B/16 pretraining step (BASELINE.json config 2) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl avsiam_b200|reference] [--batch 256]
                    [--arrangement single_pass|two_pass]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one optimisation step over a per-GPU batch of synthetic AudioSet-shaped samples (randn fbank
[B,1024,128] + one randn frame [B,3,224,224], 75 % random mask): forward (shared ViT-B/16 encoder over both
modalities, fusion blocks + MAE decoder + masked-MSE, global-batch InfoNCE), hand-written reverse pass, bucketed
gradient all-reduce (N>1) and the fused Adam step. `value` = samples/s over all ranks with inputs resident in
HBM; `e2e` = the same through the public nn.Module call with HOST inputs (pinned H2D copy + loss read-back inside
the timed region). `--impl reference` times the reference algorithm's CPU restatement (oracle/, "port": the
reference itself is a pure-Python research dump that cannot travel to the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "AV pretrain samples/sec"
UNIT = "samples/s"
TRAIN_GFLOP_PER_SAMPLE = {"single_pass": 242.0, "two_pass": 474.0}   # BASELINE.md §4 (3 x forward), ViT-B/16
# BASELINE config 5 geometries. ViT-H/14 (head_dim 80, patch 14, 657 + 256 tokens) runs as the
# reference ships it: no decoder, contrastive-only, single-direction InfoNCE (cav_mae_huge.cpython-39.pyc).
MODEL_DIMS = {"vit_b": {}, "vit_l": dict(embed_dim=1024, depth=24, heads=16, dec_depth=6),
              "vit_h": dict(embed_dim=1280, depth=32, heads=16, patch=14, dec_depth=1)}
MODEL_NAME = {"vit_b": "ViT-B/16", "vit_l": "ViT-L/16", "vit_h": "ViT-H/14"}
# (mae_loss_weight, contrast_loss_weight, bidirectional InfoNCE) of the step each model runs
MODEL_LOSS = {"vit_b": (1.0, 0.01, True), "vit_l": (1.0, 0.01, True), "vit_h": (0.0, 0.01, False)}


def train_gflop_per_sample(dims, keep_a, keep_v, with_mae=True):
    """3 x forward FLOPs of the single-pass step (2 per multiply-add): encoder over the kept tokens, two fusion blocks,
    decoder over all tokens, patch embedding and prediction heads; attention 4 S^2 hd per head."""
    D, Dd, p = dims.embed_dim, dims.dec_dim, dims.patch

    def block(tokens, seqs, width, heads):
        gemm = 2.0 * tokens * width * width * 12            # qkv 3 + proj 1 + fc1 4 + fc2 4
        attn = sum(4.0 * S * S * (width // heads) * heads for S in seqs)
        return gemm + attn

    enc_tokens = keep_a + keep_v
    f = dims.depth * block(enc_tokens, [keep_a, keep_v], D, dims.heads)
    if not with_mae:      # contrastive-only (ViT-H/14): encoder + patch embedding
        return 3.0 * (f + 2.0 * keep_a * D * p * p + 2.0 * keep_v * D * p * p * dims.in_chans) / 1e9
    f += 2 * block(enc_tokens, [enc_tokens], D, dims.heads)
    f += dims.dec_depth * block(dims.Ta + dims.Tv, [dims.Ta + dims.Tv], Dd, dims.dec_heads)
    f += 2.0 * keep_a * D * p * p + 2.0 * keep_v * D * p * p * dims.in_chans + 2.0 * enc_tokens * D * Dd
    f += 2.0 * dims.Ta * Dd * p * p + 2.0 * dims.Tv * Dd * p * p * dims.in_chans
    return 3.0 * f / 1e9
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}  # B200_PROFILING.md


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {k: float(d[k]) for k in FALLBACK_PEAKS}, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step_fn(arrangement: str, B: int):
    """The reference algorithm's CPU restatement (oracle/avsiam_oracle.py, pinned to the unmodified reference by
    tests/golden): literal train-step body of traintest_cavmae_base.py:131-152 minus autocast/GradScaler, fp32."""
    from oracle import avsiam_oracle as O
    d = O.VIT_B
    sd = O.init_state(d, seed=0, skip_heads=True)
    params = {k: v.requires_grad_(True) for k, v in sd.items()}
    g = torch.Generator().manual_seed(87)
    audio = torch.randn(B, d.audio_len, d.mel, generator=g)
    imgs = torch.randn(B, d.in_chans, d.img, d.img, generator=g)
    plist = list(params.values())
    opt1 = torch.optim.Adam(plist, 2e-4, weight_decay=5e-7, betas=(0.95, 0.999))
    opt2 = torch.optim.Adam(plist, 2e-4, weight_decay=5e-7, betas=(0.95, 0.999))
    counter = [0]

    def step():
        counter[0] += 1
        plan = O.make_mask_plan(B, d, 1234 + counter[0], two_pass=(arrangement == "two_pass"))
        if arrangement == "single_pass":
            out = O.forward_single_pass(audio, imgs, params, d, plan, mae_loss_weight=1.0, contrast_loss_weight=0.01)
            opt1.zero_grad(set_to_none=True)
            out[0].backward()
            opt1.step()
        else:
            out = O.forward(audio, imgs, params, d, plan, mae_loss_weight=0, contrast_loss_weight=1)
            opt1.zero_grad(set_to_none=True)
            out[0].backward()
            opt1.step()
            out = O.forward(audio, imgs, params, d, plan, mae_loss_weight=1, contrast_loss_weight=0)
            opt2.zero_grad(set_to_none=True)
            out[0].backward()
            opt2.step()
        return float(out[0])

    return step


def time_cpu(arrangement: str, B: int, steps: int, warmup: int):
    step = cpu_reference_step_fn(arrangement, B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    dt = time.perf_counter() - t0
    assert loss == loss, "CPU reference produced NaN"
    return B * steps / dt, dt / steps


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch
    # all host cores, whatever the launcher exported (torch.distributed.run sets OMP_NUM_THREADS=1 for N > 1)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cores = torch.get_num_threads()
    sps, s_per_step = time_cpu(args.arrangement, B, args.steps, args.warmup)
    sample = (f"{args.steps} steps x batch {B} of the ViT-B/16 {args.arrangement} step (fp32, torch CPU kernels, "
              f"{cores} threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, B, 1),
        "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu_batch: int, world: int):
    model = getattr(args, "model", "vit_b")
    name = MODEL_NAME[model]
    what = "MAE + global-batch InfoNCE" if MODEL_LOSS[model][0] else "global-batch InfoNCE only (no decoder, as shipped)"
    return {"workload": f"{name} AVSiam pretrain step ({args.arrangement}), 1024x128 fbank + 1x224x224 frame, "
                        f"mask 0.75, {what}, Adam",
            "arrangement": args.arrangement, "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * world,
            "parallelism": f"dp{world}", "l2": "inputs (288 MB/step) and activations (>40 GB/step) exceed the 126 MB L2"}


# ----------------------------------------------------------------------------------------------------- checkers
def oracle_state_from_model(model, dev, dtype=torch.float32):
    """The model's current weights as the oracle's flat {key: tensor} state (test infrastructure: the oracle is used
    here only as the CHECKER of the step-0 loss and as the library-call comparator, never in the timed product path)."""
    return {k: v.detach().to(dev, dtype).clone() for k, v in model.state_dict().items()
            if not k.startswith("my_blocks.") and ".head." not in k}


def check_step0_loss(model, net, args, dev, rank, world, dist):
    """One untimed forward with supplied mask indices, compared with the fp32 oracle evaluated by the same rank on its
    own batch (InfoNCE over the all-gathered oracle embeddings when world > 1): the N-GPU bench refuses to report a
    number for a path whose loss is wrong. Returns (loss, oracle_loss)."""
    from oracle import avsiam_oracle as O
    d = O.Dims(**MODEL_DIMS[args.model])
    B = min(args.batch, 16)          # bounded: the checker runs the fp32 oracle
    g = torch.Generator().manual_seed(4321 + rank)
    audio = torch.randn(B, d.audio_len, d.mel, generator=g).to(dev)
    imgs = torch.randn(B, d.in_chans, d.img, d.img, generator=g).to(dev)
    plan = O.make_mask_plan(B, d, 99 + rank, two_pass=(args.arrangement == "two_pass"))
    model.mask_plan = plan
    mae_w, c_w, bidirect = MODEL_LOSS[args.model]
    with torch.no_grad():
        out = net(audio, imgs, 0.75, 0.75, mae_loss_weight=mae_w, contrast_loss_weight=c_w)
    model.mask_plan = None
    state = oracle_state_from_model(model, dev)

    def gather(x):
        if world == 1:
            return x
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous())
        return torch.cat(parts, 0)

    with torch.no_grad():
        if args.arrangement == "single_pass":
            ref = O.forward_single_pass(audio, imgs, state, d, plan.to(dev), mae_loss_weight=mae_w,
                                        contrast_loss_weight=c_w, gather=gather, bidirect=bidirect)
        else:
            ref = O.forward(audio, imgs, state, d, plan.to(dev), mae_loss_weight=1.0, contrast_loss_weight=0.01,
                            gather=lambda x: gather(x))
    got, want = float(out[0]), float(ref[0])
    if not abs(got - want) <= 1e-3 * abs(want) + 1e-4:
        raise SystemExit(f"bench.py: rank {rank}: step-0 loss {got} differs from the fp32 oracle {want}")
    return got, want


def time_library_baseline(args, dev, model, steps=3, warmup=2):
    """The comparator SURVEY.md 8(d) asks for: the same step written as plain torch library calls (the oracle module's
    F.linear / F.layer_norm / F.gelu -> cuBLAS and torch kernels, attention through F.scaled_dot_product_attention as in the
    reference, cav_mae_base.py:65-68) under torch.autocast(bfloat16) with torch autograd and torch.optim.Adam(fused=True),
    on THIS B200, same batch, nothing recomputed. A reported comparator, never part of the product path."""
    from oracle import avsiam_oracle as O
    import torch.nn.functional as F
    d = O.Dims(**MODEL_DIMS[args.model])
    mae_w, c_w, bidirect = MODEL_LOSS[args.model]
    B = args.batch
    torch.cuda.empty_cache()
    state = {k: v.requires_grad_(True) for k, v in oracle_state_from_model(model, dev).items()}
    params = list(state.values())
    opt = torch.optim.Adam(params, 2e-4, weight_decay=5e-7, betas=(0.95, 0.999), fused=True)
    audio = torch.randn(B, d.audio_len, d.mel, device=dev)
    imgs = torch.randn(B, d.in_chans, d.img, d.img, device=dev)
    plan = O.make_mask_plan(B, d, 7, two_pass=False).to(dev)
    # the library's own fused attention instead of the oracle's explicit softmax(QK^T)V (memory, and a fair comparator)
    orig_attention = O.attention

    def sdpa_attention(x, sd, pfx, heads):
        Bq, N, C = x.shape
        hd = C // heads
        qkv = F.linear(x, sd[pfx + "qkv.weight"], sd[pfx + "qkv.bias"]).reshape(Bq, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(Bq, N, C)
        return F.linear(o, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])

    O.attention = sdpa_attention
    try:
        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = O.forward_single_pass(audio, imgs, state, d, plan, mae_loss_weight=mae_w, contrast_loss_weight=c_w,
                                            bidirect=bidirect)
            opt.zero_grad(set_to_none=True)
            out[0].backward()
            opt.step()
            return out[0]

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        assert bool(torch.isfinite(loss)), "library baseline: non-finite loss"
    finally:
        O.attention = orig_attention
        del state, params, opt
        torch.cuda.empty_cache()
    return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "what": "torch 2.11 eager, autocast(bfloat16), cuBLAS GEMMs + F.scaled_dot_product_attention + torch autograd + "
                    f"torch.optim.Adam(fused=True); same {MODEL_NAME[args.model]} single_pass step, same batch, same GPU; activations kept "
                    "by autograd (no recomputation)"}


# ----------------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch.distributed as dist
    import avsiam_b200
    from avsiam_b200 import B200DDP, CAVMAE_BASE, FusedAdam, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — avsiam_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    torch.manual_seed(0)                      # identical random-init weights on every rank
    dims = avsiam_b200.Dims(**MODEL_DIMS[args.model]) if MODEL_DIMS[args.model] else None
    model = CAVMAE_BASE(audio_length=1024, norm_pix_loss=False, modality_specific_depth=23, tr_pos=False,
                        arrangement=args.arrangement, dims=dims, bidirect_contrast=MODEL_LOSS[args.model][2]).to(dev)
    mae_w, c_w, _ = MODEL_LOSS[args.model]
    with torch.no_grad():                     # the reference zero-inits these (cav_mae_base.py:312-337); any value works
        for n in ("mask_token", "decoder_pos_embed_a", "decoder_pos_embed_v", "decoder_modality_a", "decoder_modality_v"):
            getattr(model, n).normal_(std=0.02)
    model.direct_grads = True
    net = B200DDP(model) if world > 1 else model
    adam_kw = dict(lr=2e-4, weight_decay=5e-7, betas=(0.95, 0.999))      # traintest_cavmae_base.py:64-66
    opt1 = FusedAdam(model.parameters(), model=model, **adam_kw)
    opt2 = FusedAdam(model.parameters(), model=model, **adam_kw) if args.arrangement == "two_pass" else None

    torch.manual_seed(87 + rank)              # run_cavmae_pretrain_base.py:113
    n_host = 2
    host_a = [torch.randn(B, 1024, 128).pin_memory() for _ in range(n_host)]
    host_v = [torch.randn(B, 3, 224, 224).pin_memory() for _ in range(n_host)]
    dev_a = [t.to(dev) for t in host_a]
    dev_v = [t.to(dev) for t in host_v]
    h2d_bytes = host_a[0].numel() * 4 + host_v[0].numel() * 4

    graphed = None
    if args.graph and args.arrangement == "single_pass":
        # forward + reverse pass + Adam as ONE CUDA graph (avsiam_b200.GraphedTrainStep); falls back to eager launches,
        # and says so on the JSON line, if this torch / NCCL combination cannot capture the step
        try:
            graphed = avsiam_b200.GraphedTrainStep(net, opt1, 0.75, 0.75, mae_loss_weight=mae_w, contrast_loss_weight=c_w)
            graphed(torch.randn(B, 1024, 128, device=dev), torch.randn(B, 3, 224, 224, device=dev))
            torch.cuda.synchronize()
        except Exception as e:   # noqa: BLE001
            print(f"bench.py: CUDA-graph capture unavailable ({type(e).__name__}: {str(e)[:200]}); eager launches", file=sys.stderr)
            graphed = None
            opt1.capturable = False

    def eager_step(a, v):
        out = net(a, v, 0.75, 0.75, mae_loss_weight=mae_w, contrast_loss_weight=c_w)
        opt1.zero_grad()
        out[0].backward()
        opt1.step()
        return out[0]

    def train_step(a, v):
        if graphed is not None:
            return graphed(a, v)
        if args.arrangement == "single_pass":
            out = net(a, v, 0.75, 0.75, mae_loss_weight=mae_w, contrast_loss_weight=c_w)
            opt1.zero_grad()
            out[0].backward()
            opt1.step()
        else:                                 # literal two-pass body, traintest_cavmae_base.py:131-152
            out = net(a, v, 0.75, 0.75, mae_loss_weight=0, contrast_loss_weight=1)
            opt1.zero_grad()
            out[0].backward()
            opt1.step()
            out = net(a, v, 0.75, 0.75, mae_loss_weight=1, contrast_loss_weight=0)
            opt2.zero_grad()
            out[0].backward()
            opt2.step()
        return out[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- (0) untimed correctness gate: step-0 loss against the fp32 oracle (every rank; global InfoNCE at N > 1)
    step0 = None
    if not args.no_loss_check:
        step0 = check_step0_loss(model, net, args, dev, rank, world, dist)

    # ---- (1) device-resident inputs: `value`
    for i in range(args.warmup):
        train_step(dev_a[i % n_host], dev_v[i % n_host])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = train_step(dev_a[i % n_host], dev_v[i % n_host])
    e1.record()
    barrier()
    launches = ops.launch_count() if graphed is None else graphed.launches_per_step * args.steps   # graph kernel nodes
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    loss_val = float(loss)
    if loss_val != loss_val:
        raise SystemExit("bench.py: loss is NaN")
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    if args.detail:      # developer aid: per-shape device time of one step, printed to stderr
        with ops.KernelTimer(detail=True) as kt:
            (eager_step if graphed is not None else train_step)(dev_a[0], dev_v[0])
        rows = sorted(kt.summary().items(), key=lambda kv: -kv[1]["ms"])
        for k, v in rows:
            rate = v["work"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
            print(f"[detail] {v['ms']:9.3f} ms  {v['calls']:4d} calls  {rate:8.1f} T/s  {k}", file=sys.stderr)
    if args.kernel_only:  # ncu runs: stop after the device-resident timed region
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_per_step,
                              "gpu_launches": int(launches), "note": "kernel-only run (profiling aid)"}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- (2) end to end through the public module call with HOST inputs (double-buffered pinned H2D + loss D2H)
    copy_stream = torch.cuda.Stream()
    stage_a = [torch.empty_like(dev_a[0]) for _ in range(2)]
    stage_v = [torch.empty_like(dev_v[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(args.steps + args.warmup, dtype=torch.float32).pin_memory()

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            stage_a[s].copy_(host_a[i % n_host], non_blocking=True)
            stage_v[s].copy_(host_v[i % n_host], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n, base):
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            s = i % 2
            torch.cuda.current_stream().wait_event(ready[s])
            l = train_step(stage_a[s], stage_v[s])
            consumed[s].record()
            loss_host[base + i].copy_(l.detach(), non_blocking=True)

    for s in range(2):
        consumed[s].record()
    e2e_loop(args.warmup, 0)
    barrier()
    e0.record()
    e2e_loop(args.steps, args.warmup)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    assert bool(torch.isfinite(loss_host).all()), "e2e: non-finite loss"

    # ---- (3) per-family device time of one more step (roofline of the dominant kernel family)
    peaks, peaks_src = load_peaks()
    with ops.KernelTimer() as kt:
        (eager_step if graphed is not None else train_step)(dev_a[0], dev_v[0])   # per-kernel events need eager launches
    fam = kt.summary()
    fam_total = sum(f["ms"] for f in fam.values())
    gemm = fam.get("gemm", {"ms": 0.0, "work": 0.0, "calls": 0})
    gemm_tflops = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    if os.path.exists(tpath) and args.arrangement == "single_pass" and B == 256 and args.model == "vit_b":
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    roofline = {
        "kernel": "avs::gemm2_bf16_kernel (two-CTA tcgen05.mma.cta_group::2 + TMEM + TMA: Linear / patch-embed fwd, dgrad) + "
                  "avs::gemm_bf16_kernel (one-CTA: wgrad)",
        "bound": "tensor", "achieved": gemm_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": gemm_tflops / peaks["bf16_tflops_sustained"], "peak_source": f"{peaks_src} (sustained cuBLAS bf16)",
        "traffic": traffic, "traffic_unit": "bytes of DRAM read + write per launch (average over the launches of one step)",
        "traffic_source": traffic_src, "launches_per_step": gemm["calls"], "ms_per_step": gemm["ms"],
        "share_of_kernel_time": gemm["ms"] / fam_total if fam_total else None,
    }
    families = {k: {"ms": round(v["ms"], 3), "calls": v["calls"],
                    "rate": (v["work"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else None)}
                for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    if args.model == "vit_b":
        gflop_sample = TRAIN_GFLOP_PER_SAMPLE[args.arrangement]
    else:
        md = model.dims
        gflop_sample = train_gflop_per_sample(md, int(md.Ta * 0.25), int(md.Tv * 0.25), with_mae=mae_w != 0)
    step_util = (value / world) * gflop_sample * 1e9 / (peaks["bf16_tflops"] * 1e12)

    # ---- (4) CPU baseline (rank 0, N=1 only): bounded sample of the same step on the host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.model == "vit_b":
        cores = torch.get_num_threads()
        sps, _ = time_cpu(args.arrangement, args.cpu_batch, 3, 1)
        cpu_baseline = {"value": sps, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"3 steps x batch {args.cpu_batch} (1 warm-up) of the same ViT-B/16 "
                                  f"{args.arrangement} step, fp32 torch CPU, oracle/avsiam_oracle.py"}

    # ---- (5) library-call comparator on the same GPU (rank 0, N=1 only): torch eager bf16
    library_baseline = None
    if rank == 0 and world == 1 and args.library_baseline and args.arrangement == "single_pass":
        try:
            library_baseline = time_library_baseline(args, dev, model)
        except torch.cuda.OutOfMemoryError as e:    # reported, not hidden: the comparator keeps every activation
            library_baseline = {"unavailable": f"out of memory at batch {B}: {str(e)[:120]}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, B, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "cuda_graph": graphed is not None,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "library_baseline": library_baseline,
            "step0_loss_check": None if step0 is None else {"loss": step0[0], "oracle_fp32": step0[1], "rtol": 1e-3},
            "tensor_pipe_util_vs_burst_peak": step_util,
            "train_gflop_per_sample": gflop_sample,
            "kernel_families_ms": families,
            "loss": loss_val,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured graph holds NCCL kernels: release it (and everything else that references the communicators)
        # before the process group goes away, or the teardown waits forever
        if graphed is not None:
            graphed.graph = None
            graphed.static_out = None
        graphed = None
        torch.cuda.synchronize()
        dist.barrier()
        # bounded teardown: the measurement is printed; a communicator that refuses to die must not hold the launcher
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(30.0)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="avsiam_b200", choices=["avsiam_b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (BASELINE.json config 2: 256)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="batch of the CPU reference sample (config 1: 2)")
    ap.add_argument("--arrangement", default="single_pass", choices=["single_pass", "two_pass"])
    ap.add_argument("--model", default="vit_b", choices=sorted(MODEL_DIMS), help="encoder geometry (config 5: vit_l, vit_h)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="issue every launch from Python (no CUDA graph)")
    ap.add_argument("--no-loss-check", action="store_true", help="skip the untimed step-0 loss check against the fp32 oracle")
    ap.add_argument("--no-library-baseline", dest="library_baseline", action="store_false",
                    help="skip the torch-eager bf16 comparator (N=1 only)")
    ap.add_argument("--detail", action="store_true", help="print a per-shape kernel time table to stderr")
    ap.add_argument("--kernel-only", action="store_true", help="stop after the device-resident timed region (ncu aid)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3                       # timing rule: W >= 3
    if args.model == "vit_h" and args.arrangement != "single_pass":
        raise SystemExit("bench.py: --model vit_h runs the single_pass arrangement only (CAVMAE_HUGE as shipped)")
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
