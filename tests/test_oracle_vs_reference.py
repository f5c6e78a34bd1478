"""Oracle vs the live reference, imported in place from /root/reference (skipped where
the tree is absent, e.g. on the GPU box).  CPU, bit-exact."""
import pytest
import torch

from oracle import attacks as oa
from oracle import ref_import, vit

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def ns():
    return ref_import.load()


@pytest.fixture(scope="module")
def setup():
    cfg = vit.VitConfig(image_size=64, patch=32, width=128, layers=2, heads=2, mlp=512, out_dim=64)
    m = vit.ClipVisionModel(vit.build(cfg, seed=3)).eval()
    mo = vit.ClipVisionModel(vit.build(cfg, seed=4)).eval()
    g = torch.Generator().manual_seed(99)
    x = torch.rand(4, 3, 64, 64, generator=g)
    d0 = (torch.rand(4, 3, 64, 64, generator=g) * 2 - 1) * 4 / 255
    y = torch.randint(0, 10, (4,), generator=g)
    text = torch.nn.functional.normalize(torch.randn(64, 10, generator=g), dim=0)
    with torch.no_grad():
        eo = mo(x, output_normalize=False)
    return m, x, d0, y, text, eo


@pytest.mark.parametrize("norm,eps,step", [("linf", 4 / 255, 1 / 255), ("l2", 2.0, 0.5)])
@pytest.mark.parametrize("mode", ["max", "min"])
def test_pgd(ns, setup, norm, eps, step, mode):
    m, x, d0, y, text, eo = setup
    a = oa.pgd(m, oa.ComputeLossWrapper(eo, None, "mean", "l2", 100.0), x, y, norm, eps, 3, step,
               False, perturbation=d0.clone(), mode=mode)
    b = ns.pgd(m, ns.ComputeLossWrapper(eo, None, "mean", "l2", 100.0), x, y, norm, eps, 3, step,
               False, perturbation=d0.clone(), mode=mode)
    assert torch.equal(a, b)


@pytest.mark.parametrize("norm,eps", [("linf", 4 / 255), ("l2", 2.0), ("l2", 1.5)])
@pytest.mark.parametrize("n_iter", [10, 25])
def test_apgd_train(ns, setup, norm, eps, n_iter):
    m, x, d0, y, text, eo = setup
    a = oa.apgd_train(m, x, y, norm, eps, n_iter=n_iter,
                      loss_fn=oa.ComputeLossWrapper(eo, text, "none", "ce", 100.0))
    b = ns.apgd_train(m, x, y, norm, eps, n_iter=n_iter,
                      loss_fn=ns.ComputeLossWrapper(eo, text, "none", "ce", 100.0))
    assert torch.equal(a, b)


@pytest.mark.parametrize("norm,eps", [("Linf", 4 / 255), ("L2", 2.0), ("L2", 1.5)])
@pytest.mark.parametrize("loss", ["ce", "dlr"])
def test_apgd_attack(ns, setup, norm, eps, loss):
    m, x, d0, y, text, eo = setup
    cm = vit.ClassificationModel(m.model, text, 100.0)
    with torch.no_grad():
        yc = cm(x).argmax(1)
    A = oa.APGDAttack(cm, n_iter=15, norm=norm, n_restarts=2, eps=eps, seed=0, loss=loss, alpha=2.0)
    R = ns.APGDAttack(cm, n_iter=15, norm=norm, n_restarts=2, eps=eps, seed=0, loss=loss, alpha=2.0)
    assert torch.equal(A.perturb(x, yc), R.perturb(x, yc))
    ra, rr = A.attack_single_run(x, yc), R.attack_single_run(x, yc)  # consumes the CPU RNG equally


@pytest.mark.parametrize("norm,eps", [("Linf", 4 / 255), ("L2", 2.0)])
def test_apgd_targeted(ns, setup, norm, eps):
    """autoattack/autopgd_base.py:584-707 (targeted DLR over the 2nd..4th ranked classes)"""
    m, x, d0, y, text, eo = setup
    cm = vit.ClassificationModel(m.model, text, 100.0)
    with torch.no_grad():
        yc = cm(x).argmax(1)
    kw = dict(n_iter=12, norm=norm, n_restarts=1, eps=eps, seed=0, n_target_classes=3, alpha=2.0)
    A, R = oa.APGDAttack_targeted(cm, **kw), ns.APGDAttack_targeted(cm, **kw)
    assert torch.equal(A.perturb(x, yc), R.perturb(x, yc))


def test_model_wrappers(ns, setup):
    m, x, d0, y, text, eo = setup
    r = ns.ClipVisionModel(m.model, None, m.normalize)
    with torch.no_grad():
        for flag in (False, True):
            assert torch.equal(m(x, output_normalize=flag), r(x, output_normalize=flag))


@pytest.mark.parametrize("norm,eps", [("Linf", 4 / 255), ("L2", 1.5)])
def test_apgd_attack_variants(ns, setup, norm, eps):
    """Less-travelled branches of autoattack/autopgd_base.py: EOT averaging (:267-286,371-386),
    perturb(best_loss=True) (:482-502), y=None (:470-473), use_rs=False (:219-220), an explicit
    x_init through attack_single_run (:221-224)."""
    m, x, d0, y, text, eo = setup
    cm = vit.ClassificationModel(m.model, text, 100.0)
    with torch.no_grad():
        yc = cm(x).argmax(1)
    kw = dict(n_iter=9, norm=norm, n_restarts=2, eps=eps, seed=3, alpha=2.0)
    A, R = oa.APGDAttack(cm, eot_iter=2, **kw), ns.APGDAttack(cm, eot_iter=2, **kw)
    assert torch.equal(A.perturb(x, yc), R.perturb(x, yc))
    A, R = oa.APGDAttack(cm, loss="dlr", **kw), ns.APGDAttack(cm, loss="dlr", **kw)
    torch.manual_seed(11)
    a = A.perturb(x, yc, best_loss=True)
    torch.manual_seed(11)
    b = R.perturb(x, yc, best_loss=True)
    assert torch.equal(a, b)
    assert torch.equal(A.perturb(x), R.perturb(x))                       # y = None -> predictions
    A, R = oa.APGDAttack(cm, use_rs=False, **kw), ns.APGDAttack(cm, use_rs=False, **kw)
    assert torch.equal(A.perturb(x, yc), R.perturb(x, yc))
    x0 = (x + d0).clamp(0, 1)
    ra, rr = A.attack_single_run(x, yc, x_init=x0), R.attack_single_run(x, yc, x_init=x0)
    for u, v in zip(ra, rr):
        assert torch.equal(u, v)


def test_apgd_train_initial_stepsize(ns, setup):
    """`initial_stepsize` (train/apgd_train.py:168-169), as the in-training evaluation passes it
    when clean_weight > 0 (train/adversarial_training_clip.py:407)."""
    m, x, d0, y, text, eo = setup
    for norm, eps in (("linf", 4 / 255), ("l2", 1.5)):
        a = oa.apgd_train(m, x, y, norm, eps, n_iter=12, initial_stepsize=0.05 * eps,
                          loss_fn=oa.ComputeLossWrapper(eo, text, "none", "ce", 100.0))
        b = ns.apgd_train(m, x, y, norm, eps, n_iter=12, initial_stepsize=0.05 * eps,
                          loss_fn=ns.ComputeLossWrapper(eo, text, "none", "ce", 100.0))
        assert torch.equal(a, b)


def test_apgd_use_largereps(ns, setup):
    """`use_largereps` (autoattack/autopgd_base.py:493-498, 517-520, 550-582): the 3 eps / 2 eps / eps
    schedule chained through x_init, with the constructor's check-point schedule."""
    m, x, d0, y, text, eo = setup
    cm = vit.ClassificationModel(m.model, text, 100.0)
    with torch.no_grad():
        yc = cm(x).argmax(1)
    kw = dict(n_iter=20, norm="Linf", n_restarts=2, eps=2 / 255, seed=5, alpha=2.0,
              use_largereps=True)
    A, R = oa.APGDAttack(cm, **kw), ns.APGDAttack(cm, **kw)
    assert torch.equal(A.perturb(x, yc), R.perturb(x, yc))       # CPU tensors: randn_like on the CPU RNG
    assert (A.n_iter, A.eps, A.use_rs) == (R.n_iter, R.eps, R.use_rs)
