"""Golden vectors for the targeted APGD (authoring container only; same conventions as
make_golden.py, kept in a separate file so the first fixture stays byte-stable).

    python tests/golden/make_golden_targeted.py   # writes tests/golden/reference_vectors_targeted.pt

Outputs of the REAL reference `autoattack/autopgd_base.py::APGDAttack_targeted.perturb`
(imported in place from /root/reference) around the oracle's restated ViT, on the inputs of
reference_vectors.pt.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import, vit  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ns = ref_import.load()
    gold = torch.load(os.path.join(HERE, "reference_vectors.pt"), weights_only=False)
    cfg = vit.VitConfig(**gold["meta"]["cfg"])
    m = vit.ClipVisionModel(vit.build(cfg, seed=gold["meta"]["seeds"][0])).eval()
    cm = vit.ClassificationModel(m.model, gold["text"], 100.0)
    x, y_clean = gold["x"], gold["y_clean"]
    out = {"meta": {"n_iter": 20, "n_target_classes": 3, "n_restarts": 1, "seed": 0, "alpha": 2.0}}
    for norm, eps in (("Linf", 4 / 255), ("L2", 1.0)):
        atk = ns.APGDAttack_targeted(cm, n_iter=20, norm=norm, n_restarts=1, eps=eps, seed=0,
                                     n_target_classes=3, alpha=2.0)
        out[f"apgd_t_{norm}"] = atk.perturb(x, y_clean).detach()
        with torch.no_grad():
            out[f"apgd_t_{norm}_pred"] = cm(out[f"apgd_t_{norm}"]).argmax(1)
    # AutoAttack orchestrator (autoattack/autoattack.py:12-261), custom version as in
    # CLIP_eval/clip_robustbench.py:289-300: apgd-ce then apgd-t on the still-robust points
    import importlib
    aa = importlib.import_module("autoattack.autoattack")
    g = torch.Generator().manual_seed(77)
    xa = torch.rand(12, 3, 64, 64, generator=g)
    with torch.no_grad():
        ya = cm(xa).argmax(1)
    ya[0] = (ya[0] + 1) % 10                      # one point misclassified from the start
    eps = 1.2 / 255                              # mid-plateau: robust accuracy is 0.25 for eps in [1.0, 1.5]/255
    A = aa.AutoAttack(cm, norm="Linf", eps=eps, version="custom", attacks_to_run=["apgd-ce", "apgd-t"],
                      alpha=2.0, verbose=False, device="cpu", iterations_apgd=10, seed=0)
    A.apgd.n_restarts = 1
    A.apgd_targeted.n_target_classes = 3
    x_adv, y_adv = A.run_standard_evaluation(xa, ya, bs=5, return_labels=True)
    out["aa"] = {"x": xa, "y": ya, "eps": eps, "x_adv": x_adv, "y_adv": y_adv, "bs": 5,
                 "iterations_apgd": 10, "n_target_classes": 3,
                 "robust_acc": float((y_adv == ya).float().mean())}
    print("AutoAttack robust accuracy", out["aa"]["robust_acc"])
    path = os.path.join(HERE, "reference_vectors_targeted.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
