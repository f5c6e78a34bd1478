"""Golden vectors for the VLM-evaluation attacks (authoring container only).

    python tests/golden/make_golden_vlm.py     # writes tests/golden/reference_vectors_vlm.pt

Outputs of the REAL reference `vlm_eval/attacks/pgd.py::PGD.perturb` and
`vlm_eval/attacks/apgd.py::{APGD.perturb, apgd}` (imported in place from /root/reference)
around `tests/util.py::ToyVlmLoss`, whose gradient is bit-identical on CPU and CUDA, so the CUDA
path can be compared bit for bit.  Inputs are [1, frames, 3, H, W] (OpenFlamingo layout: dim 1 =
context frames ..., query frame last).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_import  # noqa: E402
from util import ToyVlmLoss  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPE = (1, 3, 3, 32, 32)


def cases():
    """(name, kind, kwargs) -- shared with the tests through the fixture's meta."""
    return [
        ("pgd_linf_context", "pgd", dict(norm="linf", eps=8 / 255, mode="max", mask_out="context",
                                         image_space=True, iterations=12, stepsize=2 / 255)),
        ("pgd_linf_query_min", "pgd", dict(norm="linf", eps=4 / 255, mode="min", mask_out="query",
                                           image_space=True, iterations=7, stepsize=1 / 255)),
        ("pgd_linf_nomask_free", "pgd", dict(norm="linf", eps=4 / 255, mode="max", mask_out=None,
                                             image_space=False, iterations=6, stepsize=1 / 255)),
        ("pgd_l2_int_mask", "pgd", dict(norm="l2", eps=1.5, mode="max", mask_out=1,
                                        image_space=True, iterations=8, stepsize=0.4)),
        ("apgd_linf_context", "apgd", dict(norm="linf", eps=8 / 255, mask_out="context",
                                           initial_stepsize=1.0, iterations=40)),
        ("apgd_linf_none_pert", "apgd", dict(norm="linf", eps=4 / 255, mask_out="none",
                                             initial_stepsize=None, iterations=25,
                                             decrease_every=0.15, decrease_every_max=0.05,
                                             pert_init=True)),
        ("apgd_l2_query", "apgd", dict(norm="l2", eps=1.5, mask_out="query", initial_stepsize=2.0,
                                       iterations=30)),
    ]


def main():
    ns = ref_import.load()
    g = torch.Generator().manual_seed(2024)
    x = torch.rand(SHAPE, generator=g)
    pert = (torch.rand(SHAPE, generator=g) * 2 - 1) * 4 / 255
    model = ToyVlmLoss(x, seed=5)
    out = {"meta": {"shape": SHAPE, "model_seed": 5, "cases": cases()}, "x": x, "pert": pert}
    for name, kind, kw in cases():
        kw = dict(kw)
        if kind == "pgd":
            it, ss = kw.pop("iterations"), kw.pop("stepsize")
            atk = ns.vlm_pgd.PGD(model.scalar, **kw)
            res = atk.perturb(x.clone(), it, ss, perturbation=pert.clone())
        else:
            it = kw.pop("iterations")
            pi = pert.clone() if kw.pop("pert_init", False) else None
            atk = ns.vlm_apgd.APGD(model, **kw)
            res = atk.perturb(x.clone(), it, pert_init=pi)
        out[name] = res.detach()
        print(name, "max|delta|", float((res - x).abs().max()), "loss", float(model(res)),
              "loss clean", float(model(x)))
    # the raw apgd() triple once (x_best, loss_best, x_best_adv)
    xb, lb, xba = ns.vlm_apgd.apgd(model, x.clone(), norm="linf", eps=8 / 255, n_iter=20, alpha=1.0)
    out["apgd_raw"] = {"x_best": xb.detach(), "loss_best": lb.detach(), "x_best_adv": xba.detach()}
    path = os.path.join(HERE, "reference_vectors_vlm.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
