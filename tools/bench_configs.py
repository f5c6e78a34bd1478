"""Secondary measurements for BASELINE.json configs 3 and 4 (bench.py measures config 2).

  config 3: ViT-L/14 TeCoA, 10-step PGD linf eps=2/255, CE-logit loss, batch 256
  config 4: ViT-L/14 APGD-CE l2 eps=2.0, 100 iterations (1 restart timed), batch 128
Writes gpurun_out/bench_configs.json (copied into profiles/ by hand).
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustvlm_b200 import attacks, model as rm  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
dev = torch.device("cuda:0")


def timed(fn, n=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n / 1e3, out


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip3", action="store_true", help="config 4 only")
    ap.add_argument("--n-iter", type=int, default=100, help="APGD iterations of config 4")
    ap.add_argument("--restarts", type=int, default=2, help="APGD restarts of config 4")
    args = ap.parse_args()
    res = {}
    vt = rm.VisionTransformer("ViT-L-14", seed=1).to(dev)
    m = rm.ClipVisionModel(vt, None, None).eval()
    g = torch.Generator().manual_seed(0)
    text = torch.nn.functional.normalize(torch.randn(768, 1000, generator=g), dim=0).to(dev)
    # ---- config 3
    B = 256
    x = torch.rand(B, 3, 224, 224, generator=g).to(dev)
    y = torch.randint(0, 1000, (B,), generator=g).to(dev)
    eps = 2 / 255
    d0 = ((torch.rand(B, 3, 224, 224, generator=g) * 2 - 1) * eps).to(dev)
    wrap = rm.ComputeLossWrapper(None, text, "mean", "ce", 100.0)
    f3 = lambda: attacks.pgd(m, wrap, x, y, "linf", eps, 10, 1 / 255, True, perturbation=d0,  # noqa: E731
                             mode="max")
    if not args.skip3:
        timed(f3)
        t3, adv = timed(f3, 2)
        with torch.no_grad():
            l_clean = float(wrap(m(x, output_normalize=True), y))
            l_adv = float(wrap(m(adv, output_normalize=True), y))
        res["config3_tecoa_pgd10_linf"] = {"batch": B, "seconds": t3, "adv_img_per_s": B / t3,
                                           "ce_clean": l_clean, "ce_adv": l_adv}
        print(res["config3_tecoa_pgd10_linf"], flush=True)
    # ---- config 4
    B = 128
    x = x[:B].contiguous()
    cm = rm.ClassificationModel(m, text)
    with torch.no_grad():
        yc = cm(x).argmax(1)
    # random-init towers are broken by the first restart (robust accuracy 0 %), after which
    # perturb() has nothing left to attack; to time `restarts` FULL restarts the still-robust set
    # is held fixed (best_loss=True: every restart attacks all points, autopgd_base.py:482-502)
    atk = attacks.APGDAttack(cm, n_iter=args.n_iter, norm="L2", n_restarts=args.restarts, eps=2.0,
                             seed=0, alpha=2.0)
    if args.n_iter == 100:          # untimed warm-up (workspace allocation, kernel attributes)
        attacks.APGDAttack(cm, n_iter=2, norm="L2", n_restarts=1, eps=2.0, seed=0,
                           alpha=2.0).perturb(x, yc)
    t0 = time.time()
    t4, adv = timed(lambda: atk.perturb(x, yc, best_loss=True))
    with torch.no_grad():
        racc = float((cm(adv).argmax(1) == yc).float().mean())
    res["config4_apgd_ce_l2_100it"] = {"batch": B, "seconds": t4, "n_iter": args.n_iter,
                                       "restarts": args.restarts,
                                       "model_evals_per_s": B * (args.n_iter + 1) * args.restarts / t4,
                                       "adv_img_per_s": B / t4,
                                       "adv_img_per_s_per_restart": B * args.restarts / t4,
                                       "robust_acc_after": racc, "wall_s": time.time() - t0,
                                       "max_l2": float((adv - x).flatten(1).norm(dim=1).max())}
    print(res["config4_apgd_ce_l2_100it"], flush=True)
    os.makedirs(OUT, exist_ok=True)
    json.dump(res, open(os.path.join(OUT, "bench_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
