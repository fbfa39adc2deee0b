"""Worker of tests/test_nccl_multi_gpu.py — launched by torch.distributed.run, one process per GPU (NCCL).

Checks the data-parallel path on hardware against the fp32 oracle evaluated on the GLOBAL batch by the same process
(on its GPU, plain torch): the reference's semantics are DDP's mean of the per-rank losses, where every rank's loss is
its own MAE loss plus the InfoNCE over the all-gathered embeddings (cav_mae_base.py:724-735, gather_layer.py:21-37,
traintest_cavmae_base.py:58-59), i.e. the objective  mean_r(mae_r) + c_w * nce(global batch).
Asserted on every rank: the rank's loss, the all-reduced gradient of EVERY parameter (cosine >= 0.999 and norm within
3 %), identical gradients and identical post-Adam weights on all ranks. Exits non-zero on any failure.
"""
import dataclasses
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from avsiam_b200 import B200DDP, CAVMAE_BASE, Dims, FusedAdam  # noqa: E402
from oracle import avsiam_oracle as O  # noqa: E402
from oracle.make_golden import synth_inputs  # noqa: E402


def cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))


def slice_plan(plan: O.MaskPlan, r: int, B: int) -> O.MaskPlan:
    return O.MaskPlan(plan.ids_shuffle_a[r * B:(r + 1) * B], plan.ids_shuffle_v[r * B:(r + 1) * B])


def run_case(name, d, arrangement, B, dev, rank, world, seed):
    torch.manual_seed(1)
    model = CAVMAE_BASE(dims=Dims(**dataclasses.asdict(d)), arrangement=arrangement)
    sd = O.init_state(d, seed=0)
    if rank != 0:                          # B200DDP must broadcast rank 0's weights like DDP does
        sd_other = O.init_state(d, seed=5)
        model.load_state_dict(O.with_aliases(sd_other), strict=True)
    else:
        model.load_state_dict(O.with_aliases(sd), strict=True)
    model = model.to(dev)
    model.direct_grads = True
    net = B200DDP(model, device_ids=[dev.index], output_device=dev.index, find_unused_parameters=True)
    opt = FusedAdam(model.parameters(), 2e-4, weight_decay=5e-7, betas=(0.95, 0.999))
    audio, imgs = synth_inputs(world * B, d, seed)
    mae_w, c_w = 1.0, 0.01
    state = {k: v.to(dev).requires_grad_(True) for k, v in sd.items() if ".head." not in k}
    if arrangement == "single_pass":
        gplan = O.make_mask_plan(world * B, d, seed + 1, two_pass=False)
        plans = [slice_plan(gplan, r, B) for r in range(world)]
        ref = O.forward_single_pass(audio.to(dev), imgs.to(dev), state, d, gplan.to(dev), mae_loss_weight=mae_w,
                                    contrast_loss_weight=c_w)
        ref_total = ref[0]
        # the rank's own loss: its MAE part + the global InfoNCE
        with torch.no_grad():
            sl = slice(rank * B, (rank + 1) * B)
            mine = O.forward_single_pass(audio[sl].to(dev), imgs[sl].to(dev), state, d, plans[rank].to(dev),
                                         mae_loss_weight=mae_w, contrast_loss_weight=0.0)
            ref_rank_loss = float(mine[0]) + float(ref[4])
    else:
        plans = [O.make_mask_plan(B, d, seed + 10 + r, two_pass=True) for r in range(world)]
        maes, cas, cvs = [], [], []
        for r in range(world):
            sl = slice(r * B, (r + 1) * B)
            a, v, p = audio[sl].to(dev), imgs[sl].to(dev), plans[r].to(dev)
            maes.append(O.forward(a, v, state, d, p, mae_loss_weight=mae_w, contrast_loss_weight=0.0)[0])
            ca, cv = O.forward_encoder_mmixed(a, v, state, d, p)
            cas.append(ca.mean(dim=1)); cvs.append(cv.mean(dim=1))
        nce, _ = O.contrastive(torch.cat(cas), torch.cat(cvs), bidirect=True)
        ref_total = sum(maes) / world + c_w * nce
        ref_rank_loss = float(maes[rank]) + c_w * float(nce)
    ref_total.backward()
    ref_grads = {k: v.grad for k, v in state.items() if v.grad is not None}

    model.mask_plan = plans[rank]
    sl = slice(rank * B, (rank + 1) * B)
    out = net(audio[sl].to(dev), imgs[sl].to(dev), 0.75, 0.75, mae_loss_weight=mae_w, contrast_loss_weight=c_w)
    loss = float(out[0])
    rtol = 1e-3 if d is O.VIT_B else 2e-2      # same bounds as tests/test_model_gpu.py (reference geometry / TINY)
    assert abs(loss - ref_rank_loss) <= rtol * abs(ref_rank_loss) + 1e-4, (name, "rank loss", loss, ref_rank_loss)
    opt.zero_grad()
    out[0].backward()
    torch.cuda.synchronize()
    arena = model.arena
    # weights were broadcast from rank 0
    named = dict(model.named_parameters())
    # gradients: every parameter against the global-batch oracle
    worst, bad = 1.0, []
    used = set(model.last_used_names())
    assert used == set(ref_grads), (sorted(used - set(ref_grads))[:4], sorted(set(ref_grads) - used)[:4])
    for k, g in ref_grads.items():
        if float(g.norm()) < 1e-12:
            continue
        mine = arena.grad(k)
        c = cos(mine, g)
        worst = min(worst, c)
        rn = float(mine.double().norm() / g.double().norm())
        if c < 0.999 or abs(rn - 1) > 0.03:
            bad.append((k, c, rn))
    assert not bad, (name, bad[:6])
    # identical on all ranks (bitwise: every rank holds the all-reduced buffer)
    used_sum = torch.stack([arena.grad(k).double().sum() for k in sorted(used)]).sum().reshape(1)
    sums = [torch.zeros_like(used_sum) for _ in range(world)]
    dist.all_gather(sums, used_sum)
    assert all(float(s) == float(sums[0]) for s in sums), (name, "gradients differ across ranks", sums)
    opt.step()
    torch.cuda.synchronize()
    w_sum = arena.flat[:arena.n_hot].double().sum().reshape(1)
    sums = [torch.zeros_like(w_sum) for _ in range(world)]
    dist.all_gather(sums, w_sum)
    assert all(float(s) == float(sums[0]) for s in sums), (name, "weights differ across ranks after Adam", sums)
    upd = named["vit_base.blocks.0.attn.qkv.weight"].detach().cpu() - sd["vit_base.blocks.0.attn.qkv.weight"]
    assert float(upd.abs().max()) > 1e-5, "Adam did not move rank 0's broadcast weights"
    if rank == 0:
        print(f"[nccl W={world}] {name}: rank-0 loss {loss:.6f} (oracle {ref_rank_loss:.6f}), global objective "
              f"{float(ref_total):.6f}, worst gradient cosine {worst:.6f} over {len(ref_grads)} parameters, "
              f"all-reduced bytes {model.grad_sync.bytes_last_step}", flush=True)
    del model, net, opt, state, ref_grads
    torch.cuda.empty_cache()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    try:
        run_case("TINY single_pass B=5/rank", O.TINY, "single_pass", 5, dev, rank, world, 300)
        run_case("TINY two_pass B=6/rank", O.TINY, "two_pass", 6, dev, rank, world, 310)
        run_case("ViT-B/16 single_pass B=3/rank", O.VIT_B, "single_pass", 3, dev, rank, world, 320)
    finally:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("[nccl] all cases passed", flush=True)


if __name__ == "__main__":
    main()
