"""Workload for ONE `ncu --set full` capture that contains every kernel of the library once or
twice at the ViT-L/14 bench shapes (B = 256, T = 257, W = 1024, F = 4096), on a TWO-layer tower
so the capture stays short:

    ncu --set full --import-source on --clock-control none -f -o gpurun_out/r02_kernels \
        python tools/profile_kernels.py
    python tools/ncu_summary.py report gpurun_out/r02_kernels.ncu-rep > profiles/r02_kernels_ncu.txt

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


def main():
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
        attacks.pgd(m, w, x, None, "linf", eps, 1, 1 / 255, False, perturbation=d0, mode="max")
        if mode == "fast16":
            wce = rm.ComputeLossWrapper(None, text, "mean", "ce", 100.0)
            attacks.pgd(m, wce, x, y, "l2", 2.0, 1, 0.5, True, perturbation=d0, mode="max")
            cm = rm.ClassificationModel(m, text)
            atk = attacks.APGDAttack(cm, n_iter=3, norm="L2", n_restarts=1, eps=2.0, seed=0)
            atk.perturb(x[:64], cm(x[:64]).argmax(1))
            atk = attacks.APGDAttack(cm, n_iter=3, norm="Linf", n_restarts=1, eps=eps, seed=0, loss="dlr")
            atk.perturb(x[:64], cm(x[:64]).argmax(1))
            eng = m.engine()
            eng.loss_grad(x[:64], LossSpec("dlr-targeted", True, text=text, logit_scale=100.0,
                                           targets=y[:64], pred_on_logits=True, y_target=(y[:64] + 1) % 1000))
        torch.cuda.synchronize()
        del m, vt
        torch.cuda.empty_cache()
    vt = rm.VisionTransformer(ARCH, seed=1).to(dev)
    vto = rm.VisionTransformer(ARCH, seed=1).to(dev)
    tr = AdversarialTrainer(rm.ClipVisionModel(vt, None, None).eval(),
                            rm.ClipVisionModel(vto, None, None).eval(), iterations=1)
    tr.train_step(x[:128], None)
    torch.cuda.synchronize()
    print("profile workload done")


if __name__ == "__main__":
    main()
