"""The reference-facing call surface (robustvlm_b200.attacks / .model) vs the CPU oracle.

The tests read like the reference's own usage (train/adversarial_training_clip.py:300-333,
CLIP_eval/clip_robustbench.py:289-300): build ClipVisionModel + ComputeLossWrapper, call
pgd / apgd_train / APGDAttack.perturb with the reference's argument conventions."""
import json
import os

import pytest
import torch

from oracle import attacks as oa
from oracle import vit as ovit

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
NAME = "tiny-32"


def _log(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "test_metrics.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **kw}) + "\n")
    print(name, kw)


def _pair(seed):
    """(ours on cuda, oracle on cpu) with identical weights."""
    from robustvlm_b200 import model as rm
    o = ovit.build(NAME, seed=seed)
    vt = rm.VisionTransformer(NAME)
    vt.load_state_dict(o.state_dict())
    vt = vt.cuda()
    return rm.ClipVisionModel(vt, None, None).eval(), ovit.ClipVisionModel(o).eval()


def _data(B, seed, eps):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, 224, 224, generator=g)
    d0 = (torch.rand(B, 3, 224, 224, generator=g) * 2 - 1) * eps
    return x, d0


def _loss_after(model_cpu, wrap_cpu, x_adv, y, normalize):
    with torch.no_grad():
        return float(wrap_cpu(model_cpu(x_adv.cpu(), output_normalize=normalize), y).mean())


@pytest.mark.parametrize("norm,eps,step", [("linf", 4 / 255, 1 / 255), ("l2", 2.0, 0.5)])
@pytest.mark.parametrize("path", ["fused", "generic"])
def test_pgd_fare(path, norm, eps, step):
    from robustvlm_b200 import attacks, model as rm
    B, iters = 4, 3
    ours, ref_m = _pair(1)
    ours_o, ref_o = _pair(2)
    x, d0 = _data(B, 7, min(eps, 0.05))
    with torch.no_grad():
        e_ref = ref_o(x, output_normalize=False)
        e_our = ours_o(x.cuda(), output_normalize=False)
    assert float((e_our.cpu() - e_ref).norm() / e_ref.norm()) < 1e-3
    wrap_ref = oa.ComputeLossWrapper(e_ref, None, "mean", "l2", 100.0)
    ref = oa.pgd(ref_m, wrap_ref, x, None, norm, eps, iters, step, False, perturbation=d0.clone(),
                 mode="max")
    if path == "fused":
        wrap = rm.ComputeLossWrapper(e_our, None, "mean", "l2", 100.0)
    else:  # a plain callable defeats the fused dispatch -> autograd through the engine VJP
        w0 = rm.ComputeLossWrapper(e_our, None, "mean", "l2", 100.0)
        wrap = lambda out, t: w0(out, t)  # noqa: E731
    out = attacks.pgd(ours, wrap, x.cuda(), None, norm, eps, iters, step, False,
                      perturbation=d0.clone().cuda(), mode="max")
    torch.cuda.synchronize()
    out = out.cpu()
    assert out.min() >= 0 and out.max() <= 1
    if norm == "linf":
        assert float((out - x).abs().max()) <= eps + 1e-6
    else:
        assert float((out - x).flatten(1).norm(dim=1).max()) <= eps * (1 + 1e-5)
    l_ours = _loss_after(ref_m, wrap_ref, out, None, False)
    l_ref = _loss_after(ref_m, wrap_ref, ref, None, False)
    same = float(((out - x) == (ref - x)).float().mean())
    _log("pgd_fare", path=path, norm=norm, same=same, loss_ours=l_ours, loss_ref=l_ref,
         max_abs_err=float((out - ref).abs().max()))
    # measured (profiles/r02_test_metrics.jsonl): loss within 7e-7 relative, 99.937 % of delta identical
    assert abs(l_ours - l_ref) <= 1e-5 * abs(l_ref)
    if norm == "linf":
        assert same > 0.9987


def test_pgd_tecoa_ce():
    from robustvlm_b200 import attacks, model as rm
    B, iters, eps, step = 4, 3, 2 / 255, 1 / 255
    ours, ref_m = _pair(1)
    x, d0 = _data(B, 11, eps)
    torch.manual_seed(0)
    text = torch.nn.functional.normalize(torch.randn(64, 10), dim=0)
    y = torch.randint(0, 10, (B,))
    wrap_ref = oa.ComputeLossWrapper(None, text, "mean", "ce", 100.0)
    ref = oa.pgd(ref_m, wrap_ref, x, y, "linf", eps, iters, step, True, perturbation=d0.clone(),
                 mode="max")
    wrap = rm.ComputeLossWrapper(None, text.cuda(), "mean", "ce", 100.0)
    out = attacks.pgd(ours, wrap, x.cuda(), y.cuda(), "linf", eps, iters, step, True,
                      perturbation=d0.clone().cuda(), mode="max").cpu()
    l_ours = _loss_after(ref_m, wrap_ref, out, y, True)
    l_ref = _loss_after(ref_m, wrap_ref, ref, y, True)
    same = float(((out - x) == (ref - x)).float().mean())
    _log("pgd_tecoa", same=same, loss_ours=l_ours, loss_ref=l_ref)
    assert same > 0.9987 and abs(l_ours - l_ref) <= 1e-5 * abs(l_ref)    # measured 0.99938, 1.5e-6


@pytest.mark.parametrize("norm,eps", [("linf", 4 / 255), ("l2", 2.0)])
def test_apgd_train(norm, eps):
    from robustvlm_b200 import attacks, model as rm
    B = 4
    ours, ref_m = _pair(1)
    x, _ = _data(B, 13, 0.01)
    torch.manual_seed(0)
    text = torch.nn.functional.normalize(torch.randn(64, 10), dim=0)
    y = torch.randint(0, 10, (B,))
    wrap_ref = oa.ComputeLossWrapper(None, text, "none", "ce", 100.0)
    ref = oa.apgd_train(ref_m, x, y, norm, eps, n_iter=10, loss_fn=wrap_ref)
    wrap = rm.ComputeLossWrapper(None, text.cuda(), "none", "ce", 100.0)
    out = attacks.apgd_train(ours, x.cuda(), y.cuda(), norm, eps, n_iter=10, loss_fn=wrap).cpu()
    l_ours = _loss_after(ref_m, wrap_ref, out, y, True)
    l_ref = _loss_after(ref_m, wrap_ref, ref, y, True)
    _log("apgd_train", norm=norm, loss_ours=l_ours, loss_ref=l_ref,
         max_abs_err=float((out - ref).abs().max()),
         pert=float((out - x).abs().max()))
    assert out.min() >= 0 and out.max() <= 1
    if norm == "linf":
        assert float((out - x).abs().max()) <= eps + 1e-6
    else:
        assert float((out - x).flatten(1).norm(dim=1).max()) <= eps * (1 + 1e-4)
    assert abs(l_ours - l_ref) <= 6e-5 * abs(l_ref)                       # measured 2.8e-5 / 1.9e-5


@pytest.mark.parametrize("norm,eps", [("Linf", 4 / 255), ("L2", 2.0)])
def test_apgd_attack_perturb(norm, eps):
    from robustvlm_b200 import attacks, model as rm
    B = 6
    ours, ref_m = _pair(1)
    x, _ = _data(B, 17, 0.01)
    torch.manual_seed(0)
    text = torch.nn.functional.normalize(torch.randn(64, 10), dim=0)
    cm_ref = ovit.ClassificationModel(ref_m.model, text, 100.0)
    cm = rm.ClassificationModel(ours, text.cuda())
    with torch.no_grad():
        logits_ref = cm_ref(x)
        logits = cm(x.cuda()).cpu()
    _log("apgd_attack_logits", norm=norm, err=float((logits - logits_ref).abs().max()),
         scale=float(logits_ref.abs().max()))
    assert float((logits - logits_ref).abs().max()) < 1e-3 * float(logits_ref.abs().max())
    y = logits_ref.argmax(1)
    R = oa.APGDAttack(cm_ref, n_iter=20, norm=norm, n_restarts=2, eps=eps, seed=0, alpha=2.0)
    A = attacks.APGDAttack(cm, n_iter=20, norm=norm, n_restarts=2, eps=eps, seed=0, alpha=2.0)
    ref = R.perturb(x, y)
    out = A.perturb(x.cuda(), y.cuda()).cpu()
    with torch.no_grad():
        racc_ref = float((cm_ref(ref).argmax(1) == y).float().mean())
        racc_ours = float((cm_ref(out).argmax(1) == y).float().mean())
    _log("apgd_attack", norm=norm, racc_ref=racc_ref, racc_ours=racc_ours,
         pert=float((out - x).abs().max()))
    assert out.min() >= 0 and out.max() <= 1
    if norm == "Linf":
        assert float((out - x).abs().max()) <= eps + 1e-6
    assert abs(racc_ours - racc_ref) < 1e-6            # the same points end up fooled


@pytest.mark.parametrize("loss", ["dlr", "dlr-targeted"])
def test_dlr_losses_in_the_head_kernel(loss):
    """DLR (autoattack/autopgd_base.py:195-201) and targeted DLR (:616-621): value and input
    gradient from the head kernel (top-4 logits on device) vs autograd through the oracle."""
    from robustvlm_b200.engine import LossSpec
    B = 6
    ours, ref_m = _pair(1)
    x, d0 = _data(B, 23, 0.01)
    torch.manual_seed(1)
    text = torch.nn.functional.normalize(torch.randn(64, 10), dim=0)
    cm_ref = ovit.ClassificationModel(ref_m.model, text, 100.0)
    xa = (x + d0).clamp(0, 1).requires_grad_(True)
    logits = cm_ref(xa)
    y = logits.argmax(1)
    y[::2] = (y[::2] + 3) % 10                     # half of the points start misclassified
    yt = logits.sort(dim=1)[1][:, -3].detach()
    R = oa.APGDAttack_targeted(cm_ref, eps=0.1)
    R.y_target = yt
    l_ref = oa.dlr_loss(logits, y) if loss == "dlr" else R.dlr_loss_targeted(logits, y)
    (g_ref,) = torch.autograd.grad(l_ref.sum(), xa)
    spec = LossSpec(loss, True, text=text.cuda(), logit_scale=100.0, targets=y.cuda(),
                    pred_on_logits=True, y_target=yt.cuda() if loss == "dlr-targeted" else None)
    r = ours.engine().loss_grad(xa.detach().cuda(), spec, want_logits=True)
    l_err = float((r["loss"].cpu() - l_ref.detach()).abs().max())
    cos = float(torch.nn.functional.cosine_similarity(r["grad"].cpu().flatten(1), g_ref.flatten(1),
                                                      dim=1).min())
    nrm = float(r["grad"].cpu().flatten(1).norm(dim=1).div(g_ref.flatten(1).norm(dim=1)).sub(1).abs().max())
    _log("dlr_head", loss=loss, loss_err=l_err, cos_min=cos, norm_ratio_err=nrm,
         logits_err=float((r["logits"].cpu() - logits.detach()).abs().max()))
    assert torch.equal(r["pred"].cpu(), logits.argmax(1) == y)
    assert l_err < 5e-3 and cos > 0.999 and nrm < 2e-2


def test_apgd_fused_dlr_eot_and_largereps():
    """APGDAttack with loss='dlr', eot_iter = 2 and `use_largereps` on the fused engine path
    (no autograd), against the oracle run with the same noise."""
    from robustvlm_b200 import attacks, model as rm
    B = 6
    ours, ref_m = _pair(1)
    x, _ = _data(B, 29, 0.01)
    torch.manual_seed(0)
    text = torch.nn.functional.normalize(torch.randn(64, 10), dim=0)
    cm_ref = ovit.ClassificationModel(ref_m.model, text, 100.0)
    cm = rm.ClassificationModel(ours, text.cuda())
    with torch.no_grad():
        y = cm_ref(x).argmax(1)
    real_randn_like = torch.randn_like
    res = {}
    for tag, kw in (("dlr_eot", dict(loss="dlr", eot_iter=2, norm="Linf", eps=4 / 255)),
                    ("largereps", dict(loss="ce", use_largereps=True, norm="Linf", eps=2 / 255))):
        R = oa.APGDAttack(cm_ref, n_iter=20, n_restarts=2, seed=0, alpha=2.0, **kw)
        A = attacks.APGDAttack(cm, n_iter=20, n_restarts=2, seed=0, alpha=2.0, **kw)
        outs = []
        for atk, xx, yy in ((R, x, y), (A, x.cuda(), y.cuda())):
            g = torch.Generator().manual_seed(77)      # decr_eps_pgd's randn_like: same noise for both
            torch.randn_like = lambda t, g=g: torch.randn(t.shape, generator=g).to(t.device)
            try:
                real_grad = torch.autograd.grad
                if atk is A:
                    def boom(*a, **k):
                        raise AssertionError("torch.autograd.grad called on the fused path")
                    torch.autograd.grad = boom
                outs.append(atk.perturb(xx, yy).cpu())
            finally:
                torch.randn_like = real_randn_like
                torch.autograd.grad = real_grad
        ref, out = outs
        with torch.no_grad():
            racc_ref = float((cm_ref(ref).argmax(1) == y).float().mean())
            racc_ours = float((cm_ref(out).argmax(1) == y).float().mean())
        res[tag] = (racc_ref, racc_ours, float((out - x).abs().max()), float((out - ref).abs().max()))
        eps = kw["eps"]
        assert out.min() >= 0 and out.max() <= 1 and float((out - x).abs().max()) <= eps + 1e-6
        assert abs(racc_ours - racc_ref) <= 1.0 / B + 1e-6
    _log("apgd_fused_dlr_eot_largereps", **{k: list(v) for k, v in res.items()})
