"""world_size-2 gloo tests (CPU) of the N > 1 path's host logic: batch sharding leaves the
per-sample attack unchanged, and the outer-gradient all-reduce reproduces the full-batch
mean gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import attacks as oa
from oracle import vit


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from robustvlm_b200 import parallel
    torch.set_num_threads(1)
    r, w = parallel.init("gloo")
    assert (r, w) == (rank, world)
    cfg = vit.VitConfig(image_size=64, patch=32, width=128, layers=1, heads=2, mlp=256, out_dim=32)
    model = vit.ClipVisionModel(vit.build(cfg, seed=1)).eval()
    model_o = vit.ClipVisionModel(vit.build(cfg, seed=2)).eval()
    g = torch.Generator().manual_seed(5)
    B = 5                                           # uneven: 3 + 2
    x = torch.rand(B, 3, 64, 64, generator=g)
    d0 = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1) * 4 / 255
    with torch.no_grad():
        eo = model_o(x, output_normalize=False)
    xs, ds, es = (parallel.shard(t, rank, world) for t in (x, d0, eo))
    # inner loop: no collective, identical per-sample result (sign() cancels the 1/B of 'mean')
    adv_s = oa.pgd(model, oa.ComputeLossWrapper(es, None, "mean", "l2", 100.0), xs, None, "linf",
                   4 / 255, 2, 1 / 255, False, perturbation=ds.clone(), mode="max")
    adv = parallel.gather_batch(adv_s)
    # outer step: mean loss over the global batch; one all-reduce of the flat gradient
    params = [p for p in model.parameters()]
    for p in params:
        p.requires_grad_(True)
    loss = oa.l2(model(adv_s, output_normalize=False), es, "mean")
    grads = torch.autograd.grad(loss, params)
    flat = torch.cat([g_.reshape(-1) for g_ in grads])
    parallel.allreduce_mean_(flat, xs.shape[0], B)
    # the bucketed variant (per-block buckets + the rest) averages the same buffer identically
    from robustvlm_b200 import checkpoint
    rng = torch.Generator().manual_seed(100 + rank)
    layout = checkpoint.param_layout(model.model)
    n = sum(sz for *_, sz in layout)
    fg = torch.randn(n, generator=rng)
    whole = fg.clone()
    dist.all_reduce(whole)
    whole /= world
    bk = parallel.BucketedAllReduce(layout, cfg.layers, "cpu")
    bk.finish(bk.launch(fg), fg)
    assert torch.equal(fg, whole), "bucketed all-reduce differs from the monolithic one"
    if rank == 0:
        q.put((adv, flat))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_attack_and_gradient_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    adv, flat = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # single-process full-batch reference
    cfg = vit.VitConfig(image_size=64, patch=32, width=128, layers=1, heads=2, mlp=256, out_dim=32)
    model = vit.ClipVisionModel(vit.build(cfg, seed=1)).eval()
    model_o = vit.ClipVisionModel(vit.build(cfg, seed=2)).eval()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(5, 3, 64, 64, generator=g)
    d0 = (torch.rand(5, 3, 64, 64, generator=g) * 2 - 1) * 4 / 255
    with torch.no_grad():
        eo = model_o(x, output_normalize=False)
    ref = oa.pgd(model, oa.ComputeLossWrapper(eo, None, "mean", "l2", 100.0), x, None, "linf",
                 4 / 255, 2, 1 / 255, False, perturbation=d0.clone(), mode="max")
    # a handful of near-zero-gradient pixels may flip sign with the batch-size-dependent
    # summation order of BLAS; everything else is identical
    assert float((adv == ref).float().mean()) > 0.999
    params = list(model.parameters())
    loss = oa.l2(model(adv, output_normalize=False), eo, "mean")
    full = torch.cat([g_.reshape(-1) for g_ in torch.autograd.grad(loss, params)])
    assert float((flat - full).norm() / full.norm()) < 1e-5


def test_grad_buckets_partition_the_flat_buffer():
    """one contiguous bucket per transformer block + the rest, covering [0, n) exactly once"""
    from robustvlm_b200 import checkpoint, parallel
    from robustvlm_b200.model import VisionTransformer
    for arch, L in (("tiny-32", 2), ("ViT-B-32", 12)):
        layout = checkpoint.param_layout(VisionTransformer(arch))
        n = sum(sz for *_, sz in layout)
        blocks, rest = parallel.grad_buckets(layout, L)
        spans = sorted(blocks + rest)
        assert len(blocks) == L and spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        for name, _, off, sz in layout:
            if name.startswith("transformer.resblocks."):
                lo, hi = blocks[int(name.split(".")[2])]
                assert lo <= off and off + sz <= hi


def test_shard_bounds_cover_everything():
    from robustvlm_b200 import parallel
    for n in (1, 5, 8, 2048, 2049):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
