"""Multi-GPU invariants on real NCCL, run under torchrun with N >= 2:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py

 1. the inner loop is per-sample: every rank attacks its shard of ONE global batch with its own
    replica; the gathered result must equal the single-process result on the whole batch BIT FOR
    BIT (no collective on the data path, batch-composition-independent kernels);
 2. the outer step: each rank's gradient of its shard, all-reduced with `parallel.allreduce_mean_`
    (one NCCL all-reduce of the flat fp32 buffer), must equal the single-GPU full-batch gradient up
    to fp32 summation order, also for UNEVEN shards; after the fused AdamW step the replicas hold
    identical parameters;
 3. the per-block bucketed all-reduce overlapped with the backward (even shards) gives the
    parameters of one blocking all-reduce (to fp32 summation order; bit for bit at N = 2) and
    bit-identical replicas.
Prints one JSON line on rank 0 and writes gpurun_out/check_multi_gpu.json; exits non-zero on failure.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustvlm_b200 import attacks, model as rm, parallel  # noqa: E402
from robustvlm_b200.engine import FMT_BF16, LossSpec, VitEngine  # noqa: E402
from robustvlm_b200.train import AdversarialTrainer  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def main():
    rank, world = parallel.init()
    assert world >= 2, "run under torchrun with at least 2 ranks"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    name, B, eps = "tiny-14", 12 + 1, 4 / 255          # 13 samples: uneven shards
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, 3, 224, 224, generator=g).to(dev)
    d0 = ((torch.rand(B, 3, 224, 224, generator=g) * 2 - 1) * eps).to(dev)
    vt, vto = rm.VisionTransformer(name, seed=1).to(dev), rm.VisionTransformer(name, seed=2).to(dev)
    m, mo = rm.ClipVisionModel(vt, None, None).eval(), rm.ClipVisionModel(vto, None, None).eval()
    res = {"world": world, "batch": B}
    # ---- 1. inner loop: sharded == whole, bit for bit
    with torch.no_grad():
        eo = mo(x, output_normalize=False)
    full = attacks.pgd(m, rm.ComputeLossWrapper(eo, None, "mean", "l2", 100.0), x, None, "linf", eps,
                       5, 1 / 255, False, perturbation=d0, mode="max")
    lo, hi = parallel.shard_bounds(B, rank, world)
    mine = attacks.pgd(m, rm.ComputeLossWrapper(eo[lo:hi], None, "mean", "l2", 100.0), x[lo:hi], None,
                       "linf", eps, 5, 1 / 255, False, perturbation=d0[lo:hi], mode="max")
    gathered = parallel.gather_batch(mine)
    res["pgd_sharded_equals_whole"] = bool(torch.equal(gathered, full))
    res["shards"] = [list(parallel.shard_bounds(B, r, world)) for r in range(world)]
    # ---- 2. outer gradient: all-reduce of shard gradients == full-batch gradient
    sd = vt.state_dict()
    eng = VitEngine(sd, fmt_fwd=FMT_BF16, fmt_bwd=FMT_BF16, device=dev)
    names = list(sd)
    sizes = [sd[k].numel() for k in names]

    def flat_grads(xs, eos):
        flat = torch.zeros(sum(sizes), device=dev)
        views, off = {}, 0
        for k, n in zip(names, sizes):
            views[k] = flat[off:off + n].view(sd[k].shape)
            off += n
        eng.train_grads(xs, LossSpec("l2", False, emb_orig=eos), views, 1.0 / xs.shape[0])
        return flat

    g_full = flat_grads(full, eo)
    g_mine = flat_grads(full[lo:hi], eo[lo:hi])
    parallel.allreduce_mean_(g_mine, hi - lo, B)
    rel = float((g_mine - g_full).norm() / g_full.norm())
    cos = float(torch.nn.functional.cosine_similarity(g_mine, g_full, dim=0))
    res["grad_allreduce_rel_err"], res["grad_allreduce_cos"] = rel, cos
    # ---- 3. trainer: replicas stay identical after steps on different shards
    tr = AdversarialTrainer(m, mo, lr=1e-4, wd=1e-4, steps=20, warmup=2, iterations=2)
    for _ in range(2):
        tr.train_step(x[lo:hi], None, global_batch=B)
    p = tr.flat_p.clone()
    ref = p.clone()
    dist.broadcast(ref, src=0)
    res["replicas_identical_after_2_steps"] = bool(torch.equal(p, ref))
    # ---- 4. even shards: the per-block bucketed all-reduce overlapped with the backward
    # (parallel.BucketedAllReduce behind rvlm_vit_train_grads_ev's events) gives the same parameters,
    # bit for bit, as one blocking all-reduce after the backward
    Be = (B // world) * world
    le, he = parallel.shard_bounds(Be, rank, world)
    outs = []
    for overlap in (True, False):
        vt2 = rm.VisionTransformer(name, seed=1).to(dev)
        m2 = rm.ClipVisionModel(vt2, None, None).eval()
        tr2 = AdversarialTrainer(m2, mo, lr=1e-4, wd=1e-4, steps=20, warmup=2, attack="none")
        if not overlap:
            tr2._bucketed = False              # force the blocking path
        for _ in range(2):
            tr2.train_step(x[le:he], None, global_batch=Be)
        outs.append(tr2.flat_p.clone())
        used_overlap = tr2._bucketed not in (None, False)
        if overlap:
            res["bucketed_path_used"] = bool(used_overlap)
    # NCCL picks its algorithm (ring / tree / NVLS) by message size, so the 50 MB buckets and the
    # 1.2 GB buffer may be summed in different orders at N > 2: equal to fp32 rounding, bit for
    # bit at N = 2; either way every replica must hold the same bits
    res["bucketed_equals_blocking_bitwise"] = bool(torch.equal(outs[0], outs[1]))
    res["bucketed_vs_blocking_rel_diff"] = float((outs[0] - outs[1]).norm() / outs[1].norm())
    ref_b = outs[0].clone()
    dist.broadcast(ref_b, src=0)
    res["bucketed_replicas_identical"] = bool(torch.equal(outs[0], ref_b))
    flags = torch.tensor([int(res["pgd_sharded_equals_whole"]), int(rel < 2e-2 and cos > 0.999),
                          int(res["replicas_identical_after_2_steps"]),
                          int(res["bucketed_vs_blocking_rel_diff"] < 1e-6 and res["bucketed_path_used"]
                              and res["bucketed_replicas_identical"])],
                         device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok = bool(flags.min().item() == 1)
    if rank == 0:
        res["ok"] = ok
        print(json.dumps(res), flush=True)
        os.makedirs(OUT, exist_ok=True)
        json.dump(res, open(os.path.join(OUT, "check_multi_gpu.json"), "w"), indent=1)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
