"""Generate golden vectors from the REAL reference (authoring container only).

    python tests/golden/make_golden.py        # writes tests/golden/reference_vectors.pt

The reference ships no tests or fixtures for this path, so the golden
vectors are outputs of the reference's own functions -- imported in place from
/root/reference (never copied) -- on small seeded inputs:
  * vlm_eval/attacks/utils.py  project_perturbation / normalize_grad
  * train/pgd_train.py         pgd            (ell_inf and ell_2, FARE and TeCoA losses)
  * train/apgd_train.py        apgd_train     (ell_inf and ell_2)
  * autoattack/autopgd_base.py APGDAttack.perturb (ell_inf and ell_2, 2 restarts)
  * train/adversarial_training_clip.py  l2 / ce / ComputeLossWrapper
around the oracle's restated ViT (the model is not part of the reference tree).  The
model is a 2-layer, 64x64-pixel configuration so every tensor stays small.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import, vit  # noqa: E402

CFG = vit.VitConfig(image_size=64, patch=32, width=128, layers=2, heads=2, mlp=512, out_dim=64)


def build_models():
    m = vit.ClipVisionModel(vit.build(CFG, seed=1)).eval()
    mo = vit.ClipVisionModel(vit.build(CFG, seed=2)).eval()
    return m, mo


def main():
    ns = ref_import.load()
    assert not hasattr(ns, "import_error"), ns.import_error
    g = torch.Generator().manual_seed(1234)
    B, C = 3, 10
    eps_inf = 4 / 255
    x = torch.rand(B, 3, 64, 64, generator=g)
    d0 = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1) * eps_inf
    y = torch.randint(0, C, (B,), generator=g)
    text = torch.nn.functional.normalize(torch.randn(64, C, generator=g), dim=0)
    m, mo = build_models()
    with torch.no_grad():
        e_orig = mo(x, output_normalize=False)
    out = {"meta": {"cfg": CFG.__dict__, "seeds": (1, 2), "B": B, "C": C},
           "x": x, "d0": d0, "y": y, "text": text, "e_orig": e_orig}

    # primitives
    p = torch.randn(B, 3, 64, 64, generator=g) * 0.05
    out["project_linf"] = ns.attack_utils.project_perturbation(p, eps_inf, "linf")
    out["project_l2"] = ns.attack_utils.project_perturbation(p * 40, 2.0, "l2")
    out["normgrad_linf"] = ns.attack_utils.normalize_grad(p, "linf")
    out["normgrad_l2"] = ns.attack_utils.normalize_grad(p, "l2")
    out["prim_in"] = p
    # losses
    emb = torch.randn(B, 64, generator=g)
    out["loss_in"] = emb
    out["l2_none"] = ns.l2(emb, e_orig, "none")
    out["l2_mean"] = ns.l2(emb, e_orig, "mean")
    out["ce_mean"] = ns.compute_loss("ce", emb, y, None, 100.0, text, "mean")
    # pgd
    fare = ns.ComputeLossWrapper(e_orig, None, "mean", "l2", 100.0)
    tecoa = ns.ComputeLossWrapper(None, text, "mean", "ce", 100.0)
    out["pgd_fare_linf"] = ns.pgd(m, fare, x, y, "linf", eps_inf, 4, 1 / 255, False,
                                  perturbation=d0.clone(), mode="max")
    out["pgd_fare_l2"] = ns.pgd(m, fare, x, y, "l2", 1.0, 4, 0.3, False,
                                perturbation=d0.clone(), mode="max")
    out["pgd_tecoa_linf_min"] = ns.pgd(m, tecoa, x, y, "linf", 2 / 255, 3, 1 / 255, True,
                                       perturbation=d0.clone() * 0.5, mode="min")
    # apgd_train
    tecoa_none = ns.ComputeLossWrapper(None, text, "none", "ce", 100.0)
    out["apgd_train_linf"] = ns.apgd_train(m, x, y, "linf", eps_inf, n_iter=10, loss_fn=tecoa_none)
    out["apgd_train_l2"] = ns.apgd_train(m, x, y, "l2", 1.0, n_iter=10, loss_fn=tecoa_none)
    # evaluation APGD on the zero-shot classifier
    cm = vit.ClassificationModel(m.model, text, 100.0)
    with torch.no_grad():
        y_clean = cm(x).argmax(1)
    out["y_clean"] = y_clean
    for norm, eps in (("Linf", eps_inf), ("L2", 1.0)):
        atk = ns.APGDAttack(cm, n_iter=20, norm=norm, n_restarts=2, eps=eps, seed=0, alpha=2.0)
        out[f"apgd_eval_{norm}"] = atk.perturb(x, y_clean).detach()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
