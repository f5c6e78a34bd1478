"""Whole FARE training step (BASELINE.json config 5 / SURVEY.md §8d "also report whole-train-step
img/s"): model_orig forward, 10-step ell_inf PGD, clean forward, adversarial forward + full
backward, ONE NCCL all-reduce of the flat fp32 gradient, fused AdamW -- per-rank batch 256,
ViT-L/14, synthetic data.  Run plain (1 GPU) or under torchrun (N GPUs, weak scaling):

    python tools/bench_train.py --steps 3 --warmup 1
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py

Prints one JSON line on rank 0 (time = max over ranks, CUDA events) and writes it to
gpurun_out/bench_train_n{N}.json.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustvlm_b200 import model as rm, parallel  # noqa: E402
from robustvlm_b200.train import AdversarialTrainer  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ViT-L-14")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--iterations", type=int, default=10)
    args = ap.parse_args()
    rank, world = parallel.init()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    vt, vto = rm.VisionTransformer(args.model, seed=1), rm.VisionTransformer(args.model, seed=1)
    m = rm.ClipVisionModel(vt.to(dev), None, None)
    mo = rm.ClipVisionModel(vto.to(dev), None, None).eval()
    tr = AdversarialTrainer(m, mo, lr=1e-5, wd=1e-4, steps=20000, warmup=1400,
                            iterations=args.iterations)             # README.md:282 (FARE)
    g = torch.Generator().manual_seed(100 + rank)
    data = torch.rand(args.batch, 3, 224, 224, generator=g).to(dev)
    y = torch.zeros(args.batch, dtype=torch.long, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(args.warmup):
        losses.append(float(tr.train_step(data, y)["loss"]))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        log = tr.train_step(data, y)
    e1.record()
    barrier()
    losses.append(float(log["loss"]))
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    # the attack alone, same batch (for the split)
    from robustvlm_b200 import attacks
    with torch.no_grad():
        eo = mo(data, output_normalize=False)
    w = rm.ComputeLossWrapper(eo, None, "mean", "l2", 100.0)
    m.eval()
    d0 = torch.empty_like(data).uniform_(-4 / 255, 4 / 255)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a0.record()
    attacks.pgd(m, w, data, y, "linf", 4 / 255, args.iterations, 1 / 255, False, perturbation=d0,
                mode="max")
    a1.record()
    torch.cuda.synchronize()
    if rank == 0:
        step_ms = float(ms)
        line = {"metric": "FARE training step images/sec (ViT-L/14, 10-step PGD inner loop + outer "
                          "fwd/bwd + all-reduce + AdamW)",
                "value": world * args.batch / (step_ms / 1e3), "unit": "img/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
                "attack_ms": a0.elapsed_time(a1), "outer_ms": step_ms - a0.elapsed_time(a1),
                "per_gpu_batch": args.batch, "global_batch": world * args.batch,
                "grad_allreduce_bytes": tr.n * 4 if world > 1 else 0,
                "loss_first_last": [losses[0], losses[-1]], "scaling": "weak", "data": "synthetic",
                "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        print(json.dumps(line), flush=True)
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, f"bench_train_n{world}.json"), "w") as f:
            f.write(json.dumps(line) + "\n")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
