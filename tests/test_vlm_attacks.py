"""VLM-evaluation attacks (vlm_eval/attacks/{pgd,apgd}.py).

CPU part: the oracle restatement vs (a) the live reference imported in place (authoring
container only) and (b) the committed golden vectors made from the real reference
(tests/golden/make_golden_vlm.py) -- bit-exact.
GPU part: the CUDA path (`robustvlm_b200.vlm_attacks`, update kernels through the C ABI) vs
the same golden vectors.  The toy loss's gradient is bit-identical on CPU and CUDA
(tests/util.py::ToyVlmLoss), so the ell_inf trajectories must agree BIT FOR BIT; ell_2 within
2e-6 (different, fixed summation order of the per-sample norms)."""
import os
import sys

import pytest
import torch

from oracle import ref_import
from oracle import vlm_attacks as ov

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import ToyVlmLoss  # noqa: E402

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                    "reference_vectors_vlm.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(PATH, weights_only=False)


def _run(mod, model, x, pert, kind, kw):
    kw = dict(kw)
    if kind == "pgd":
        it, ss = kw.pop("iterations"), kw.pop("stepsize")
        return mod.PGD(model.scalar, **kw).perturb(x.clone(), it, ss, perturbation=pert.clone())
    it = kw.pop("iterations")
    pi = pert.clone() if kw.pop("pert_init", False) else None
    return mod.APGD(model, **kw).perturb(x.clone(), it, pert_init=pi)


def test_oracle_vs_golden(gold):
    model = ToyVlmLoss(gold["x"], seed=gold["meta"]["model_seed"])
    for name, kind, kw in gold["meta"]["cases"]:
        out = _run(ov, model, gold["x"], gold["pert"], kind, kw)
        assert torch.equal(out, gold[name]), name
    xb, lb, xba = ov.apgd(model, gold["x"].clone(), norm="linf", eps=8 / 255, n_iter=20, alpha=1.0)
    r = gold["apgd_raw"]
    assert torch.equal(xb, r["x_best"]) and torch.equal(xba, r["x_best_adv"])
    assert torch.equal(lb.reshape(-1), r["loss_best"].reshape(-1))


def test_apgd_actually_halves_step_and_restarts(gold):
    """The fixture exercises the check-point branches: with the concave toy loss the best
    iterate is not the last one in at least one case (restart-from-best happened)."""
    r = gold["apgd_raw"]
    assert not torch.equal(r["x_best"], r["x_best_adv"])


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted")
def test_oracle_vs_reference_live():
    ns = ref_import.load()
    g = torch.Generator().manual_seed(7)
    x = torch.rand(1, 2, 3, 16, 16, generator=g)
    pert = (torch.rand(x.shape, generator=g) * 2 - 1) * 2 / 255
    model = ToyVlmLoss(x, seed=11)
    for norm, eps, ss in (("linf", 4 / 255, 1 / 255), ("l2", 1.0, 0.3)):
        for mask_out in ("context", "query", 0, None):
            for image_space in (True, False):
                a = ov.PGD(model.scalar, norm, eps, "max", mask_out, image_space).perturb(
                    x.clone(), 5, ss, perturbation=pert.clone())
                b = ns.vlm_pgd.PGD(model.scalar, norm, eps, "max", mask_out, image_space).perturb(
                    x.clone(), 5, ss, perturbation=pert.clone())
                assert torch.equal(a, b), (norm, mask_out, image_space)
        for mask_out, n_it in (("context", 30), ("none", 17)):
            torch.manual_seed(3)
            a = ov.APGD(model, norm, eps, mask_out, initial_stepsize=1.0, random_init=True
                        ).perturb(x.clone(), n_it)
            torch.manual_seed(3)
            b = ns.vlm_apgd.APGD(model, norm, eps, mask_out, initial_stepsize=1.0, random_init=True
                                 ).perturb(x.clone(), n_it)
            assert torch.equal(a, b), (norm, mask_out)
    with pytest.raises(NotImplementedError):
        ov.set_mask(x, "bogus")
    with pytest.raises(AssertionError):
        ov.apgd(model, torch.cat([x, x]), "linf", 4 / 255)


# --------------------------------------------------------------------------------------
# GPU: product path vs the golden vectors
# --------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_vs_golden(gold):
    from robustvlm_b200 import vlm_attacks as va
    x, pert = gold["x"].cuda(), gold["pert"].cuda()
    model = ToyVlmLoss(gold["x"], seed=gold["meta"]["model_seed"]).to("cuda")
    for name, kind, kw in gold["meta"]["cases"]:
        out = _run(va, model, x, pert, kind, kw).cpu()
        ref = gold[name]
        if kw["norm"] == "linf":
            assert torch.equal(out, ref), (name, float((out - ref).abs().max()))
        else:
            assert float((out - ref).abs().max()) < 2e-6, (name, float((out - ref).abs().max()))
    xb, lb, xba = va.apgd(model, x.clone(), norm="linf", eps=8 / 255, n_iter=20, alpha=1.0)
    r = gold["apgd_raw"]
    assert torch.equal(xb.cpu(), r["x_best"]) and torch.equal(xba.cpu(), r["x_best_adv"])
    assert torch.equal(lb.cpu().reshape(-1), r["loss_best"].reshape(-1))


@pytest.mark.gpu
def test_cuda_vlm_api_edges(gold):
    from robustvlm_b200 import _native as nat
    from robustvlm_b200 import vlm_attacks as va
    x = gold["x"].cuda()
    model = ToyVlmLoss(gold["x"], seed=1).to("cuda")
    # masks: frames that are masked out stay untouched
    out = va.PGD(model.scalar, "linf", 4 / 255, "max", "context").perturb(x, 3, 1 / 255)
    assert torch.equal(out[:, :-1], x[:, :-1]) and not torch.equal(out[:, -1], x[:, -1])
    out = va.APGD(model, "linf", 4 / 255, mask_out=0).perturb(x, 5)
    assert torch.equal(out[:, 0], x[:, 0]) and not torch.equal(out[:, 1:], x[:, 1:])
    # return_loss, iterations = 0, half-precision inputs come back in their dtype
    o, loss = va.PGD(model.scalar, "linf", 4 / 255, "max", None).perturb(x, 2, 1 / 255,
                                                                           return_loss=True)
    assert loss.dim() == 0 and o.shape == x.shape
    assert torch.equal(va.PGD(model.scalar, "linf", 4 / 255).perturb(x, 0, 1 / 255), x)
    assert va.APGD(model, "linf", 4 / 255).perturb(x.half(), 3).dtype == torch.float16
    # the reference's exceptions
    with pytest.raises(NotImplementedError):
        va.PGD(model.scalar, "linf", 4 / 255, mask_out="bogus").perturb(x, 1, 1 / 255)
    with pytest.raises(NotImplementedError):
        va.PGD(model.scalar, "l1", 4 / 255, mask_out=None).perturb(x, 1, 1 / 255)
    with pytest.raises(ValueError):
        va.PGD(model.scalar, "linf", 4 / 255, mode="up").perturb(x, 1, 1 / 255)
    with pytest.raises(AssertionError):
        va.PGD(model.scalar, "linf", 4 / 255).perturb(x + 0.5, 1, 1 / 255)
    with pytest.raises(AssertionError):
        va.apgd(model, torch.cat([x, x]), "linf", 4 / 255)
    with pytest.raises(nat.RvlmError):
        va.APGD(model.to("cpu"), "linf", 4 / 255).perturb(gold["x"], 2)
