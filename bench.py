#!/usr/bin/env python
"""Benchmark of the adversarial-perturbation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one complete inner attack call -- `pgd()` with 10 iterations, ell_inf,
eps = 4/255, stepsize 1/255, FARE (l2-embedding) loss, ViT-L/14, 256 synthetic
224x224x3 images per GPU (BASELINE.json configs[1]; weak scaling: every rank attacks its
own 256 images, the inner loop needs no collective -- SURVEY.md §8e).
value   = images attacked per second over all ranks, inputs resident in HBM
e2e     = the same through the reference-facing call `robustvlm_b200.attacks.pgd` with
          pinned HOST inputs: H2D of the batch and D2H of the adversarial images inside
          the timed region.
roofline = achieved tensor throughput of the dominant kernel (the tcgen05 GEMM):
          algorithmic 2*M*N*K per launch / CUDA-event time of that launch, summed over
          every GEMM launch of the timed steps, against the measured bf16 peak.
--impl reference times the reference's own CPU PyTorch path for the same metric (the
CPU oracle restatement of train/pgd_train.py::pgd around the restated open_clip ViT --
/root/reference itself cannot travel to the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EPS, STEPSIZE, ITERS = 4 / 255, 1 / 255, 10
FWD_GF = {"ViT-L-14": 162.03, "ViT-B-32": 8.82}      # BASELINE.md §2, GFLOP / image
BWD_GF = {"ViT-L-14": 168.52, "ViT-B-32": 8.91}
METRIC = "adversarial images/sec (ViT-L/14, 10-step PGD linf)"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4)
                          if len(s) >= 6 and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def synth(B, seed, device="cpu"):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, 224, 224, generator=g)
    d0 = (torch.rand(B, 3, 224, 224, generator=g) * 2 - 1) * EPS
    return x.to(device), d0.to(device)


# ------------------------------------------------------------------------------------
def _effective_cpus():
    """CPUs this process may really use: affinity mask, capped by a cgroup v2 / v1 CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = max(1, min(n, int(q / per + 0.5)))
        except (OSError, ValueError):
            pass
    return n


def _cpu_arm_setup(model_name, Bs):
    """Oracle model + FARE loss + inputs for the CPU arm, with the intra-op thread count
    CALIBRATED on one no-grad forward (more threads than the box really gives -- or than a
    batch of Bs images can feed -- makes the CPU arm slower, which would flatter the GPU)."""
    import torch
    from oracle import attacks as oa
    from oracle import vit as ovit
    model = ovit.ClipVisionModel(ovit.build(model_name, seed=1)).eval()
    model_orig = ovit.ClipVisionModel(ovit.build(model_name, seed=2)).eval()
    x, d0 = synth(Bs, seed=0)
    n = _effective_cpus()
    cands = sorted({c for c in (4, 8, 16, 32, 64, n // 2, n) if 1 <= c <= n} | {n})
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            model(x[:2], output_normalize=False)          # warm the pool at this size
            t0 = time.perf_counter()
            model(x, output_normalize=False)
            t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = c, t
        elif t > 1.5 * best_t:
            break
    torch.set_num_threads(best)
    with torch.no_grad():
        e_orig = model_orig(x, output_normalize=False)
    wrap = oa.ComputeLossWrapper(e_orig, None, "mean", "l2", 100.0)
    return oa, model, wrap, x, d0, best, n


def torch_cuda_baseline(model_name, Bs=64, calls=2, dev_index=0):
    """The "reference's single-GPU PyTorch path" of BASELINE.json's >= 10x target, measured on the
    same box: the oracle port of the reference's pgd() (train/pgd_train.py:5-68) around the
    restated open_clip ViT, executed by STOCK PyTorch on the GPU (cuBLAS / SDPA), in fp32 (torch
    defaults, what the reference runs) and with TF32 matmuls allowed (a stronger informal
    baseline).  Bounded: `Bs` images, one warm-up + `calls` timed full 10-iteration calls per
    precision, CUDA-event timing, median."""
    import torch
    from oracle import attacks as oa
    from oracle import vit as ovit
    dev = torch.device("cuda", dev_index)
    model = ovit.ClipVisionModel(ovit.build(model_name, seed=1)).eval().to(dev)
    model_orig = ovit.ClipVisionModel(ovit.build(model_name, seed=2)).eval().to(dev)
    for p in list(model.parameters()) + list(model_orig.parameters()):
        p.requires_grad_(False)
    x, d0 = synth(Bs, seed=0, device=dev)
    res = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                e_orig = model_orig(x, output_normalize=False)
            wrap = oa.ComputeLossWrapper(e_orig, None, "mean", "l2", 100.0)

            def one():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                oa.pgd(model, wrap, x, None, "linf", EPS, ITERS, STEPSIZE, False,
                       perturbation=d0.clone(), mode="max")
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / 1e3
            one()
            ts = sorted(one() for _ in range(max(1, calls)))
            res["tf32" if tf32 else "fp32"] = Bs / ts[len(ts) // 2]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    del model, model_orig
    torch.cuda.empty_cache()
    return {"fp32_img_s": res["fp32"], "tf32_img_s": res["tf32"], "batch": Bs, "calls": calls,
            "what": "oracle port of train/pgd_train.py::pgd + restated open_clip ViT on stock "
                    f"PyTorch {torch.__version__} CUDA (cuBLAS / SDPA), same box, same GPU"}


def run_reference_cuda(args):
    """`--impl reference --ref-device cuda`: the stock-PyTorch-on-GPU baseline as its own line."""
    r = torch_cuda_baseline(args.model, args.ref_batch, max(1, args.steps))
    line = {"impl": "reference-torch-cuda", "metric": METRIC, "unit": "img/s", "n_gpus": 1,
            "value": r["fp32_img_s"], "value_tf32": r["tf32_img_s"], "steps": args.steps,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.ref_batch),
            "note": r["what"]}
    print(json.dumps(line), flush=True)


def run_reference(args):
    """CPU baseline arm: the oracle port of the reference's pgd() around the restated ViT, full
    10-iteration calls (nothing extrapolated) on a bounded number of images per step so that
    `--steps K --warmup W` ends within a few minutes."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        return run_reference_cuda(args)
    name = args.model
    oa, model, wrap, x8, d8, cores, avail = _cpu_arm_setup(name, 8)

    def one(Bs):
        w = oa.ComputeLossWrapper(wrap.embedding_orig[:Bs], None, "mean", "l2", 100.0)
        t0 = time.perf_counter()
        oa.pgd(model, w, x8[:Bs], None, "linf", EPS, ITERS, STEPSIZE, False,
               perturbation=d8[:Bs].clone(), mode="max")
        return time.perf_counter() - t0

    # images per step: the largest of 8 / 4 / 2 whose K + W full calls fit ~200 s (probe: one call
    # of 2 images; CPU GEMM efficiency is flat over these batch sizes)
    t2 = one(2)
    calls = max(1, args.steps + args.warmup)
    Bs = 2
    for cand in (8, 4):
        if t2 * cand / 2 * calls <= 200.0:
            Bs = cand
            break
    for _ in range(args.warmup):
        one(Bs)
    times = [one(Bs) for _ in range(args.steps)]
    per_call = sum(times) / len(times)
    val = Bs / per_call
    sample = (f"{Bs} images per step, all {ITERS} PGD iterations of every call timed (no "
              f"extrapolation); {name} fp32, torch {torch.__version__}, {cores} threads "
              f"(calibrated; {avail} CPUs available)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "img/s",
            "n_gpus": args.gpus, "device": "cpu (host cores of the box; this arm leaves the GPUs idle)",
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_call * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, args.batch),
            "cpu_baseline": {"value": val, "unit": "img/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gemm_traffic():
    """Measured DRAM bytes per GEMM launch (mean over the four forward GEMM variants of one
    block) from the committed ncu --set full capture; None when the summary is absent."""
    p = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    try:
        d = json.load(open(p))
        return d["mean_per_launch"], d["mean_algorithmic_per_launch"]
    except (OSError, KeyError, ValueError):
        return None, None


def workload_config(args, B):
    return {"workload": f"{args.model} FARE(l2) {ITERS}-step PGD linf eps=4/255 step=1/255, "
                        f"{B} synthetic 224x224x3 images per GPU, random-init weights",
            "per_gpu_batch": B, "chunk": args.chunk, "parallelism": f"dp{args.gpus}",
            "l2_flush": "inputs larger than L2 (154 MB images + ~40 GB activations per step)",
            "shortcuts": "last block evaluated on the class-token rows only (declared, "
                         "SURVEY.md 8d); FLOP figures stay the full algorithmic ones"}


def cpu_baseline_sample(model_name, seconds=20.0, keep=None):
    """Bounded CPU sample of the same workload (oracle port), rank 0, N = 1 only.  `keep`, if a
    dict, receives the sample's inputs / outputs so the GPU path can be compared on them (the
    second half of BASELINE.json's metric: delta max-abs-err vs the reference)."""
    import torch
    Bs = 8
    oa, model, wrap, x, d0, cores, avail = _cpu_arm_setup(model_name, Bs)
    t0 = time.perf_counter()
    oa.pgd(model, wrap, x, None, "linf", EPS, 1, STEPSIZE, False, perturbation=d0.clone(),
           mode="max")
    t1 = time.perf_counter() - t0
    n_it = int(max(1, min(ITERS, seconds // max(t1, 1e-3))))
    t0 = time.perf_counter()
    x_adv = oa.pgd(model, wrap, x, None, "linf", EPS, n_it, STEPSIZE, False,
                   perturbation=d0.clone(), mode="max")
    t = (time.perf_counter() - t0) / n_it * ITERS
    if keep is not None:
        # the reference's embedding and first-iteration gradient at x + delta0 (one more fwd + bwd)
        for p_ in model.parameters():
            p_.requires_grad_(False)
        xa = (x + d0).requires_grad_(True)
        e0 = model(xa, output_normalize=False)
        (g0,) = torch.autograd.grad(wrap(e0, None), xa)
        keep.update(x=x, d0=d0, x_adv=x_adv, n_it=n_it, e_orig=wrap.embedding_orig,
                    state_dict=model.model.state_dict(), e0=e0.detach(), g0=g0)
    return {"value": Bs / t, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"{Bs} images x {n_it} of {ITERS} PGD iterations (scaled), {model_name} fp32 "
                      f"oracle port of train/pgd_train.py::pgd, {cores} threads (calibrated; "
                      f"{avail} CPUs available)"}


# ------------------------------------------------------------------------------------
def delta_vs_reference(m, keep, dev):
    """BASELINE.json's second metric: delta = x_adv - x of the CUDA path against the CPU arm's
    (the oracle port of train/pgd_train.py::pgd, fp32) on the SAME weights, images, start and
    iteration count as the cpu_baseline sample, plus the embedding error and the first
    iteration's sign agreement.  A different fp32 summation order alone flips sign(g) where
    g ~ 0, and one flipped sign moves that entry by 2*stepsize per
    iteration, so the max-abs error is quantised in steps of 2/255 * n_it; the informative
    figures are the fraction of identical entries, the mismatch among gradients that are not
    near zero, and the loss reached."""
    import torch
    from robustvlm_b200 import attacks, model as rmodel
    from robustvlm_b200.engine import LossSpec
    m.model.load_state_dict(keep["state_dict"])            # same weights as the CPU arm
    x, d0, ref = keep["x"].to(dev), keep["d0"].to(dev), keep["x_adv"]
    wrap = rmodel.ComputeLossWrapper(keep["e_orig"].to(dev), None, "mean", "l2", 100.0)
    out = attacks.pgd(m, wrap, x, None, "linf", EPS, keep["n_it"], STEPSIZE, False,
                      perturbation=d0, mode="max")
    with torch.no_grad():
        e0 = m(x + d0, output_normalize=False).cpu()
        l_ours = float(wrap(m(out, output_normalize=False), None))
        l_ref = float(wrap(m(ref.to(dev), output_normalize=False), None))
    g = m.engine().loss_grad(x, LossSpec("l2", False, emb_orig=keep["e_orig"].to(dev)),
                             delta=d0)["grad"].cpu()
    g0 = keep["g0"]
    mis = torch.sign(g) != torch.sign(g0)
    big = g0.abs() > 0.05 * g0.abs().median()
    d_o, d_r = out.cpu() - keep["x"], ref - keep["x"]
    return {"max_abs_err": float((d_o - d_r).abs().max()),
            "mean_abs_err": float((d_o - d_r).abs().mean()),
            "identical_frac": float((d_o == d_r).float().mean()),
            "emb_rel_err": float((e0 - keep["e0"]).norm() / keep["e0"].norm()),
            "mismatch_frac": float(mis.float().mean()),
            "mismatch_frac_big": float((mis & big).float().sum() / big.float().sum()),
            "images": int(x.shape[0]), "iterations": int(keep["n_it"]),
            "fare_loss_ours": l_ours, "fare_loss_ref_delta": l_ref,
            "linf_ours": float(d_o.abs().max()), "eps": EPS}


def fare_train_step(args, dev, rank, world, timed_steps=3):
    """Config 5's unit of work (BASELINE.json configs[4]): the WHOLE FARE step of
    train/adversarial_training_clip.py:296-366 on this rank's shard -- model_orig forward, 10-step
    PGD, clean forward, adversarial forward + full backward, all-reduce of the outer gradient
    (per-block buckets overlapped with the backward, parallel.BucketedAllReduce), fused AdamW --
    through robustvlm_b200.train.AdversarialTrainer.  Returns global images / s (max over ranks)."""
    import torch
    import torch.distributed as dist
    from robustvlm_b200 import model as rmodel
    from robustvlm_b200.train import AdversarialTrainer
    B = args.batch
    vt = rmodel.VisionTransformer(args.model, seed=1, precision=args.precision).to(dev)
    vt_orig = rmodel.VisionTransformer(args.model, seed=1, precision=args.precision).to(dev)
    m = rmodel.ClipVisionModel(vt, None, None).eval()
    m_orig = rmodel.ClipVisionModel(vt_orig, None, None).eval()
    tr = AdversarialTrainer(m, m_orig, lr=1e-5, wd=1e-4, steps=20000, warmup=1400, loss="l2",
                            inner_loss="l2", norm="linf", eps=EPS, stepsize=STEPSIZE,
                            iterations=ITERS, attack="pgd", output_normalize=False)
    x, _ = synth(B, seed=500 + rank, device=dev)
    tr.train_step(x, None, global_batch=B * world)          # warm-up (allocations, NCCL channels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(timed_steps):
        tr.train_step(x, None, global_batch=B * world)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / timed_steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    peak = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    del tr, m, m_orig, vt, vt_orig
    torch.cuda.empty_cache()
    return {"value": B * world / (ms / 1e3), "unit": "img/s", "ms_per_step": ms,
            "global_batch": B * world, "steps": timed_steps, "peak_mem_gib": peak,
            "allreduce": ("none (1 GPU)" if world == 1 else
                          "NCCL AVG of the 1.216 GB fp32 gradient in per-block buckets started "
                          "behind per-block events and overlapped with the backward"),
            "what": "whole FARE step: model_orig fwd + 10-step PGD + clean fwd + adv fwd/bwd + "
                    "all-reduce + AdamW (BASELINE.json configs[4] per-GPU work)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from robustvlm_b200 import _native as nat
    from robustvlm_b200 import attacks, model as rmodel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    x_h, d0_h = synth(B, seed=100 + rank)
    x_h, d0_h = x_h.pin_memory(), d0_h.pin_memory()
    x, d0 = x_h.to(dev), d0_h.to(dev)
    chunk = min(args.chunk, B)
    vt_orig = rmodel.VisionTransformer(args.model, seed=2, precision=args.precision).to(dev)
    m_orig = rmodel.ClipVisionModel(vt_orig, None, None).eval()
    with torch.no_grad():
        e_orig = torch.cat([m_orig(x[i:i + chunk], output_normalize=False)
                            for i in range(0, B, chunk)])
    del m_orig, vt_orig              # release the frozen tower's engine and workspace
    torch.cuda.empty_cache()
    wrap = rmodel.ComputeLossWrapper(e_orig, None, "mean", "l2", 100.0)
    out_h = torch.empty_like(x_h).pin_memory()

    def make(mode):
        vt = rmodel.VisionTransformer(args.model, seed=1, precision=mode).to(dev)
        return rmodel.ClipVisionModel(vt, None, None).eval()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False, sample=True, finish=None):
        barrier()
        sampler = ClockSampler(local) if (rank == 0 and sample) else None
        if sampler:
            sampler.start()
        if profile:
            nat.lib().rvlm_profile_gemm(int(profile))   # 1: GEMM launches, 2: every bracketed kernel
        nat.lib().rvlm_launch_count_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish:
            finish()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        fn_l = nat.lib().rvlm_launch_count
        fn_l.restype = nat.ctypes.c_longlong
        launches = int(fn_l())
        prof = None
        if profile:
            from ctypes import byref, c_double, c_longlong
            tms, tfl, nl = c_double(), c_double(), c_longlong()
            nat.check(nat.lib().rvlm_profile_gemm_collect(byref(tms), byref(tfl), byref(nl)))
            nat.lib().rvlm_profile_gemm(0)
            import ctypes
            import json as _json
            need = nat.lib().rvlm_profile_breakdown(None, 0)
            buf = ctypes.create_string_buffer(int(need) + 8)
            nat.lib().rvlm_profile_breakdown(buf, len(buf))
            try:
                bd = _json.loads(buf.value.decode())
            except ValueError:
                bd = {}
            prof = (tms.value, tfl.value, nl.value, bd)
        if sampler:
            sampler.stop_flag = True
            sampler.join(timeout=2)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof, (sampler.summary() if sampler else None)

    m = make(args.precision)

    def step_resident(mm=None):
        return attacks.pgd(mm or m, wrap, x, None, "linf", EPS, ITERS, STEPSIZE, False,
                           perturbation=d0, mode="max", chunk=chunk)

    def step_e2e():
        xd = x_h.to(dev, non_blocking=True)
        dd = d0_h.to(dev, non_blocking=True)
        adv = attacks.pgd(m, wrap, xd, None, "linf", EPS, ITERS, STEPSIZE, False,
                          perturbation=dd, mode="max", chunk=chunk)
        out_h.copy_(adv, non_blocking=True)
        return adv

    for _ in range(args.warmup):
        step_resident()
    ms, launches, prof, clocks = timed(step_resident, args.steps, profile=1)
    # one more step with EVERY kernel bracketed (the events between back-to-back short kernels cost
    # ~1 % of the step, so this pass is not the one `value` is taken from)
    ms_bd, _, prof_bd, _ = timed(step_resident, 1, profile=2, sample=False)
    class E2EPipeline:
        """The end-to-end step as a data loader would drive it: every step's inputs come from
        pinned host memory and its result goes back to pinned host memory INSIDE the timed region,
        on a second stream, so that the copies of step i + 1 / step i - 1 run under the compute of
        step i (double-buffered device inputs; the public call attacks.pgd is unchanged)."""

        def __init__(self):
            self.cs = torch.cuda.Stream(device=dev)
            self.xd = [torch.empty_like(x), torch.empty_like(x)]
            self.dd = [torch.empty_like(d0), torch.empty_like(d0)]
            self.ev_in = [torch.cuda.Event(), torch.cuda.Event()]
            self.ev_cmp = [torch.cuda.Event(), torch.cuda.Event()]
            self.reset()

        def reset(self):
            torch.cuda.synchronize()
            self.i, self.primed = 0, False

        def _h2d(self, b):
            with torch.cuda.stream(self.cs):
                self.xd[b].copy_(x_h, non_blocking=True)
                self.dd[b].copy_(d0_h, non_blocking=True)
                self.ev_in[b].record(self.cs)

        def step(self):
            main = torch.cuda.current_stream()
            b = self.i & 1
            if not self.primed:                  # the first step's inputs: nothing to hide under
                self.cs.wait_stream(main)
                self._h2d(b)
                self.primed = True
            main.wait_event(self.ev_in[b])
            adv = attacks.pgd(m, wrap, self.xd[b], None, "linf", EPS, ITERS, STEPSIZE, False,
                              perturbation=self.dd[b], mode="max", chunk=chunk)
            self.ev_cmp[b].record(main)
            if self.i > 0:                       # the other buffer was read by step i - 1
                self.cs.wait_event(self.ev_cmp[b ^ 1])
            self._h2d(b ^ 1)                     # next step's inputs, under this step's compute
            self.cs.wait_event(self.ev_cmp[b])   # this step's result, under the next step's compute
            with torch.cuda.stream(self.cs):
                out_h.copy_(adv, non_blocking=True)
            adv.record_stream(self.cs)
            self.i += 1
            return adv

        def finish(self):
            torch.cuda.current_stream().wait_stream(self.cs)

    e2e_how = "copies overlapped with the neighbouring steps' compute on a second stream"
    try:
        pipe = E2EPipeline()
        for _ in range(min(args.warmup, 1)):
            pipe.step()
        pipe.finish()
        pipe.reset()
        ms_e2e, _, _, _ = timed(pipe.step, args.steps, finish=pipe.finish)
        torch.cuda.synchronize()
        e2e_ok = bool(torch.equal(out_h, step_resident().cpu()))   # same inputs, deterministic path
        del pipe
    except Exception as exc:  # noqa: BLE001 -- fall back to the plain sequence, say so
        e2e_how = f"sequential (pipelined path failed: {type(exc).__name__})"
        for _ in range(min(args.warmup, 1)):
            step_e2e()
        ms_e2e, _, _, _ = timed(step_e2e, args.steps)
        e2e_ok = None

    n_img = B * world * args.steps
    value = n_img / (ms / 1e3)
    e2e = n_img / (ms_e2e / 1e3)
    line = None
    if rank == 0:
        peak_tf, _, which = read_peaks()
        gemm_ms, gemm_fl, gemm_n, _ = prof
        breakdown = prof_bd[3]
        achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        algo_tf = (FWD_GF[args.model] + BWD_GF[args.model]) * ITERS * B * args.steps / 1e3
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": PRECISION_DTYPE[args.precision],
            "data": "synthetic", "config": workload_config(args, B),
            "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": int(2 * x_h.numel() * 4),
                    "d2h_bytes_per_step": int(x_h.numel() * 4), "copies": e2e_how,
                    "result_equals_resident_run": e2e_ok},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                         "traffic": gemm_traffic()[0],
                         "traffic_algorithmic": gemm_traffic()[1],
                         "traffic_source": "profiles/r02_gemm_traffic.json (ncu dram bytes per launch of "
                                           "this round's kernels, mean of the 4 forward GEMMs of a "
                                           "block; a profiler cannot run inside the timed bench)",
                         "peak_source": f"{which} bf16_tflops_sustained",
                         "kernel": "gemm16_2cta_kernel (tcgen05 cta_group::2, all epilogue variants)",
                         "launches": gemm_n,
                         "gemm_ms_per_step": gemm_ms / args.steps,
                         "gemm_share_of_step": gemm_ms / ms if ms else None,
                         "whole_step_algorithmic_tflops": algo_tf / (ms / 1e3)},
            # CUDA-event time of the bracketed kernels in one extra step, by kernel / GEMM shape;
            # "other" = everything not bracketed (head, class-token attention, layout)
            "kernel_breakdown": kernel_breakdown(breakdown, ms_bd, 1),
        }
    # ---- the other precision modes on the same workload (N = 1 only, bounded: 1 + 2 steps each)
    modes = {}
    if world == 1 and not args.no_modes:
        models = {args.precision: m}
        for mode in [k for k in ("fast16", "split", "fast") if k != args.precision]:
            mm = make(mode)
            step_resident(mm)
            ms_m, _, _, _ = timed(lambda mm=mm: step_resident(mm), 2, sample=False)
            modes[mode] = {"value": B * 2 / (ms_m / 1e3), "unit": "img/s",
                           "ms_per_step": ms_m / 2, "dtype": PRECISION_DTYPE[mode]}
            models[mode] = mm
        modes[args.precision] = {"value": value, "unit": "img/s", "ms_per_step": ms / args.steps,
                                 "dtype": PRECISION_DTYPE[args.precision], "default": True}
        if not args.no_cpu_baseline:
            keep = {}
            line["cpu_baseline"] = cpu_baseline_sample(args.model, keep=keep)
            for mode, mm in models.items():
                modes[mode]["delta_vs_ref"] = delta_vs_reference(mm, keep, dev)
            line["delta_vs_ref"] = modes[args.precision]["delta_vs_ref"]
        line["modes"] = modes
        models.clear()
        mm = None
    m = None                      # (the closures above keep the cell, not the engine)
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if world == 1 and rank == 0 and not args.no_torch_baseline:
        try:
            tb = torch_cuda_baseline(args.model, args.ref_batch, 2, local)
            tb["speedup_vs_fp32"] = value / tb["fp32_img_s"]
            tb["speedup_vs_tf32"] = value / tb["tf32_img_s"]
            line["torch_cuda_baseline"] = tb
        except Exception as e:  # noqa: BLE001  (an OOM of the stock path must not lose the line)
            line["torch_cuda_baseline"] = {"error": repr(e)[:200]}
        torch.cuda.empty_cache()
    if not args.no_train_step:
        try:
            ts = fare_train_step(args, dev, rank, world)
        except Exception as e:  # noqa: BLE001  (never lose the headline line to an add-on)
            ts = {"error": repr(e)[:300]}
        if rank == 0:
            line["train_step"] = ts
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


PRECISION_DTYPE = {
    "fast": "f16 fwd operands / bf16 bwd operands, f32 accumulate + residual stream",
    "fast16": "f16 fwd operands / f16 bwd operands (per-sample rescaled), f32 accumulate + residual stream",
    "split": "f16 hi+lo fwd operands (3 MMAs per product, fp32-class) / f16 bwd operands, f32 accumulate",
    "bf16": "bf16 operands, f32 accumulate",
}


def kernel_breakdown(bd, ms_total, steps):
    out, acc = {}, 0.0
    for k, v in sorted(bd.items(), key=lambda kv: -kv[1]["ms"]):
        per = v["ms"] / steps
        acc += v["ms"]
        out[k] = {"ms_per_step": round(per, 3), "launches_per_step": v["launches"] // steps,
                  "us_per_launch": round(1e3 * v["ms"] / max(v["launches"], 1), 2),
                  "share": round(v["ms"] / ms_total, 4) if ms_total else None}
        if v["gflop"] > 0:
            out[k]["tflops"] = round(v["gflop"] / v["ms"], 1)
    out["other"] = {"ms_per_step": round((ms_total - acc) / steps, 3),
                    "share": round(1 - acc / ms_total, 4) if ms_total else None}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="ViT-L-14")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--chunk", type=int, default=256, help="engine micro-batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fast16", choices=["fast", "fast16", "split"],
                    help="engine precision mode of the headline value (robustvlm_b200.engine.PRECISIONS)")
    ap.add_argument("--no-modes", action="store_true", help="skip the other precision modes")
    ap.add_argument("--no-torch-baseline", action="store_true",
                    help="skip the stock-PyTorch-on-GPU baseline (N = 1)")
    ap.add_argument("--no-train-step", action="store_true",
                    help="skip the whole-FARE-step measurement (config 5's per-GPU work)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' times stock PyTorch on the GPU (informal)")
    ap.add_argument("--ref-batch", type=int, default=64)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
