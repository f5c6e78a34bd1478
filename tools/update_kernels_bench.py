"""HBM roofline of the standalone perturbation-update kernels (SURVEY.md §8d "HBM bandwidth bounds
only the standalone update/LN/softmax kernels (report GB/s for those)").

B = 256 images of 3x224x224 fp32 (154 MB per tensor, working set 0.5-1 GB >> 126 MB L2), each entry
point timed with CUDA events over `--reps` back-to-back launches after warm-up.  Algorithmic
bytes = the tensors the reference's arithmetic must read / write once.  Writes
gpurun_out/update_kernels.json; run the same command under
`ncu --set full -k regex:'pgd_|apgd_|l2_sumsq' -c 16` for the per-kernel DRAM counters.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustvlm_b200 import _native as nat  # noqa: E402
from robustvlm_b200._native import c_float, c_int, c_int64  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peak_gbs():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_gbs_burst", "hbm_gbps_burst", "hbm_gbs", "hbm_gbps"):
            if k in p:
                return float(p[k]), f"MEASURED_PEAKS.json {k}"
        for k, v in p.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v), f"MEASURED_PEAKS.json {k}"
    except Exception:  # noqa: BLE001
        pass
    return 7700.0, "fallback 7.7 TB/s (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    B, n = args.batch, 3 * 224 * 224
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    T = lambda: torch.rand(B, n, device=dev, generator=g)  # noqa: E731
    x, delta, vel, grad, x_adv, x_old = T(), (T() - 0.5) * 0.03, torch.zeros(B, n, device=dev), T() - 0.5, T(), T()
    x_best, g_best, x_best_adv, mask = T(), T(), T(), torch.ones(B, n, device=dev)
    step = torch.full((B,), 2 / 255, device=dev)
    fn = nat.lib().rvlm_update_workspace_bytes
    fn.restype = nat.c_size_t
    ws = torch.empty(int(fn(c_int(B))), dtype=torch.uint8, device=dev)
    loss = torch.rand(B, device=dev)
    pred = torch.zeros(B, dtype=torch.int32, device=dev)
    loss_steps = torch.zeros(100, B, device=dev)
    lb, lbl, red = torch.zeros(B, device=dev), torch.zeros(B, device=dev), torch.ones(B, device=dev)
    acc, flags = torch.ones(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)
    S = nat.stream_ptr
    P = nat.ptr
    eps = 4 / 255
    cases = {
        # name: (callable, algorithmic tensors read+written)
        "pgd_update_linf (g,delta,v,x -> delta,v)": (lambda: nat.call(
            "rvlm_pgd_update_linf", P(delta), P(vel), P(grad), P(x), c_int(B), c_int64(n),
            c_float(1 / 255), c_float(eps), c_int(1), c_float(0.9), S()), 6),
        "pgd_update_linf_masked (+mask)": (lambda: nat.call(
            "rvlm_pgd_update_linf_masked", P(delta), P(vel), P(grad), P(x), P(mask), c_int(1),
            c_int(B), c_int64(n), c_float(1 / 255), c_float(eps), c_int(1), c_float(0.9), S()), 7),
        "pgd_update_l2 (4 passes; algorithmic g,delta,v,x -> delta,v)": (lambda: nat.call(
            "rvlm_pgd_update_l2", P(delta), P(vel), P(grad), P(x), c_int(B), c_int64(n),
            c_float(0.3), c_float(2.0), c_int(1), c_float(0.9), P(ws), S()), 6),
        "apgd_step_linf (x_adv,x_old,x,g -> x_adv,x_old)": (lambda: nat.call(
            "rvlm_apgd_step_linf", P(x_adv), P(x_old), P(x), P(grad), P(step), c_int(B), c_int64(n),
            c_float(eps), c_float(0.75), S()), 6),
        "apgd_step_l2 (4 passes; algorithmic as linf)": (lambda: nat.call(
            "rvlm_apgd_step_l2", P(x_adv), P(x_old), P(x), P(grad), P(step), c_int(B), c_int64(n),
            c_float(2.0), c_float(0.75), P(ws), S()), 6),
        "apgd_bookkeep, every sample fooled + improved (x_adv,g -> x_best,g_best,x_best_adv)": (
            lambda: nat.call(
                "rvlm_apgd_bookkeep", P(x_adv), P(grad), P(x_best), P(g_best), P(x_best_adv), P(loss),
                P(pred), P(loss_steps), P(lb.fill_(-1.0)), P(lbl), P(red), P(step), P(acc), P(flags),
                c_int(B), c_int64(n), c_int(100), c_int(5), c_int(0), c_float(0.0), S()), 5),
        "add_f32 (x,delta -> x_adv)": (lambda: nat.call(
            "rvlm_add_f32", P(x_best), P(x), P(delta), c_int64(B * n), S()), 3),
    }
    peak, src = peak_gbs()
    tensor_bytes = B * n * 4
    out = {"batch": B, "tensor_mb": tensor_bytes / 1e6, "peak_gbs": peak, "peak_source": src,
           "kernels": {}}
    for name, (f, ntens) in cases.items():
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        gbs = ntens * tensor_bytes / (us * 1e-6) / 1e9
        out["kernels"][name] = {"us": us, "algorithmic_bytes": ntens * tensor_bytes,
                                "achieved_gbs": gbs, "frac_of_peak": gbs / peak}
        print(f"{name:90s} {us:9.1f} us  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of peak", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "update_kernels.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
