"""Workload for ONE ncu capture that contains every kernel of the library once or twice at the
ViT-L/14 bench shapes (B = 256, T = 257, W = 1024, F = 4096), on a TWO-layer tower so the capture
stays short.  Only the calls between cudaProfilerStart / Stop are captured:

    ncu --profile-from-start off --clock-control none --metrics <tools/ncu_summary.py KEYS> \
        -f -o gpurun_out/r02_kernels python tools/profile_kernels.py
    python tools/ncu_summary.py report gpurun_out/r02_kernels.ncu-rep > profiles/r02_kernels_ncu.txt

`--step` instead runs ONE full-size 10-iteration PGD call (24 layers) for the launch list
(`ncu --metrics gpu__time_duration.sum`).

Phases (NVTX ranges rvlm.* mark forward / head / backward / update inside each):
  1. fast16: one fused ell_inf PGD iteration (all forward / backward kernels, fused delta update)
  2. split : one forward (SPLIT GEMM instantiations, split attention, lo-emitting LN / epilogues)
  3. TeCoA CE head, DLR head, ell_2 PGD update and the APGD step / bookkeeping kernels
  4. the outer training step (weight-gradient GEMMs, column reductions, AdamW) at B = 128
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustvlm_b200 import attacks, model as rm  # noqa: E402
from robustvlm_b200.engine import LossSpec  # noqa: E402
from robustvlm_b200.train import AdversarialTrainer  # noqa: E402

ARCH = dict(image=224, patch=14, width=1024, layers=2, heads=16, mlp=4096, out_dim=768)


def one_step():
    """one complete inner attack call of the bench workload (config 2)"""
    B, eps = 256, 4 / 255
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, 3, 224, 224, generator=g).to(dev)
    d0 = ((torch.rand(B, 3, 224, 224, generator=g) * 2 - 1) * eps).to(dev)
    eo = torch.randn(B, 768, generator=g).to(dev)
    m = rm.ClipVisionModel(rm.VisionTransformer("ViT-L-14", seed=1).to(dev), None, None).eval()
    w = rm.ComputeLossWrapper(eo, None, "mean", "l2", 100.0)
    m.engine()                                   # weight conversion outside the captured range
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    attacks.pgd(m, w, x, None, "linf", eps, 10, 1 / 255, False, perturbation=d0, mode="max")
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profile step done")


def main():
    if "--step" in sys.argv:
        return one_step()
    B, eps = int(os.environ.get("RVLM_PROFILE_B", "256")), 4 / 255
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, 3, 224, 224, generator=g).to(dev)
    d0 = ((torch.rand(B, 3, 224, 224, generator=g) * 2 - 1) * eps).to(dev)
    text = torch.nn.functional.normalize(torch.randn(768, 1000, generator=g), dim=0).to(dev)
    y = torch.randint(0, 1000, (B,), generator=g).to(dev)
    eo = torch.randn(B, 768, generator=g).to(dev)
    for mode in ("fast16", "split"):
        vt = rm.VisionTransformer(ARCH, seed=1, precision=mode).to(dev)
        m = rm.ClipVisionModel(vt, None, None).eval()
        w = rm.ComputeLossWrapper(eo, None, "mean", "l2", 100.0)
        eng = m.engine()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        if mode == "fast16":
            attacks.pgd(m, w, x, None, "linf", eps, 1, 1 / 255, False, perturbation=d0, mode="max")
            torch.cuda.profiler.stop()
            # head variants and the standalone update kernels: profiled alone (the model
            # evaluations around them would only repeat the kernels above)
            cm = rm.ClassificationModel(m, text)
            atk2 = attacks.APGDAttack(cm, n_iter=1, norm="L2", n_restarts=1, eps=2.0, seed=0)
            atki = attacks.APGDAttack(cm, n_iter=1, norm="Linf", n_restarts=1, eps=eps, seed=0)
            yc = cm(x).argmax(1)
            import robustvlm_b200.attacks as A
            real = eng.loss_grad

            def wrapped(*a, **k):            # capture the update kernels, not the model evaluations
                torch.cuda.profiler.stop()
                r = real(*a, **k)
                torch.cuda.synchronize()
                torch.cuda.profiler.start()
                return r
            eng.loss_grad = wrapped
            torch.cuda.profiler.start()
            atk2.perturb(x, yc)
            atki.perturb(x, yc)
            torch.cuda.profiler.stop()
            eng.loss_grad = real
            wce = rm.ComputeLossWrapper(None, text, "mean", "ce", 100.0)
            xs = x[:256]
            eng.forward(xs, output_normalize=True, keep_for_backward=False)
            torch.cuda.synchronize()
            # heads: CE, DLR, targeted DLR (forward + backward of the tower not captured twice: the
            # range covers the whole call, its GEMMs are the same kernels)
            torch.cuda.profiler.start()
            eng.loss_grad(xs, LossSpec("ce", True, text=text, logit_scale=100.0, targets=y, pred_on_logits=True), want_grad=False)
            eng.loss_grad(xs, LossSpec("dlr", True, text=text, logit_scale=100.0, targets=y, pred_on_logits=True), want_grad=False)
            # ell_2 PGD update inside the fused loop
            attacks.pgd(m, wce, x, y, "l2", 2.0, 1, 0.5, True, perturbation=d0, mode="max")
        else:
            eng.forward(x, d0, output_normalize=False, keep_for_backward=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        del m, vt
        torch.cuda.empty_cache()
    vt = rm.VisionTransformer(ARCH, seed=1).to(dev)
    vto = rm.VisionTransformer(ARCH, seed=1).to(dev)
    tr = AdversarialTrainer(rm.ClipVisionModel(vt, None, None).eval(),
                            rm.ClipVisionModel(vto, None, None).eval(), attack="none")
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr.train_step(x[:128], None)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profile workload done")


if __name__ == "__main__":
    main()
