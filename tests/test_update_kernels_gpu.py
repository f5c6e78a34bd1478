"""Perturbation-update kernels vs the oracle primitives (bit-exact for ell_inf)."""
import math

import pytest
import torch

from oracle import attacks as oa

pytestmark = pytest.mark.gpu


def _nat():
    from robustvlm_b200 import _native as nat
    return nat


def _ws(B):
    nat = _nat()
    fn = nat.lib().rvlm_update_workspace_bytes
    fn.restype = nat.c_size_t
    return torch.empty(int(fn(nat.c_int(B))), dtype=torch.uint8, device="cuda")


@pytest.mark.parametrize("maximize", [1, 0])
@pytest.mark.parametrize("shape", [(4, 3, 224, 224), (3, 3, 32, 32)])
def test_pgd_linf_bit_exact(shape, maximize):
    nat = _nat()
    torch.manual_seed(1)
    eps, step, mom = 4 / 255, 1 / 255, 0.9
    x = torch.rand(shape)
    delta = (torch.rand(shape) * 2 - 1) * eps
    vel = torch.zeros(shape)
    xd, dd, vd = x.cuda(), delta.clone().cuda(), vel.clone().cuda()
    B, n = shape[0], math.prod(shape[1:])
    for it in range(4):
        g = torch.randn(shape) * 1e-3
        g[0, 0, 0, :5] = float("nan")
        g[1, 1, 3, :7] = 0.0
        # oracle sequence (train/pgd_train.py:40-59)
        gg = g.clone()
        gg[gg.isnan()] = 0.0
        gn = oa.normalize_grad(gg, "linf")
        vel = oa.normalize_grad(mom * vel + gn, "linf")
        delta = delta + step * vel if maximize else delta - step * vel
        delta = oa.project_perturbation(delta, eps, "linf")
        delta = torch.clamp(x + delta, 0, 1) - x
        g_d = g.cuda()
        nat.call("rvlm_pgd_update_linf", nat.ptr(dd), nat.ptr(vd), nat.ptr(g_d), nat.ptr(xd),
                 nat.c_int(B), nat.c_int64(n), nat.c_float(step), nat.c_float(eps),
                 nat.c_int(maximize), nat.c_float(mom), nat.stream_ptr())
        torch.cuda.synchronize()
        assert torch.equal(dd.cpu(), delta), f"iter {it}: delta differs"
        assert torch.equal(vd.cpu(), vel), f"iter {it}: velocity differs"


def test_pgd_l2_close():
    nat = _nat()
    torch.manual_seed(2)
    shape = (4, 3, 64, 64)
    eps, step, mom = 2.0, 0.5, 0.9
    x = torch.rand(shape)
    delta = torch.randn(shape) * 0.01
    vel = torch.zeros(shape)
    xd, dd, vd = x.cuda(), delta.clone().cuda(), vel.clone().cuda()
    B, n = shape[0], math.prod(shape[1:])
    ws = _ws(B)
    for it in range(4):
        g = torch.randn(shape) * 1e-3
        gn = oa.normalize_grad(g, "l2")
        vel = oa.normalize_grad(mom * vel + gn, "l2")
        delta = delta + step * vel
        delta = oa.project_perturbation(delta, eps, "l2")
        delta = torch.clamp(x + delta, 0, 1) - x
        g_d = g.cuda()
        nat.call("rvlm_pgd_update_l2", nat.ptr(dd), nat.ptr(vd), nat.ptr(g_d), nat.ptr(xd),
                 nat.c_int(B), nat.c_int64(n), nat.c_float(step), nat.c_float(eps), nat.c_int(1),
                 nat.c_float(mom), nat.ptr(ws), nat.stream_ptr())
        torch.cuda.synchronize()
        assert (dd.cpu() - delta).abs().max() <= 1e-6 * max(1.0, float(delta.abs().max())) + 2e-7
        assert (vd.cpu() - vel).abs().max() <= 1e-6


@pytest.mark.parametrize("norm", ["Linf", "L2"])
def test_apgd_step(norm):
    nat = _nat()
    torch.manual_seed(3)
    shape = (5, 3, 32, 32)
    eps = 4 / 255 if norm == "Linf" else 1.5
    x = torch.rand(shape)
    x_adv = (x + torch.randn(shape) * 0.01).clamp(0, 1)
    x_old = (x + torch.randn(shape) * 0.01).clamp(0, 1)
    step = torch.rand(shape[0], 1, 1, 1) * eps + 0.1 * eps
    B, n = shape[0], math.prod(shape[1:])
    xa_d, xo_d = x_adv.clone().cuda(), x_old.clone().cuda()
    ws = _ws(B)
    for it, a in enumerate([1.0, 0.75, 0.75]):
        g = torch.randn(shape)
        ref, ref_old = oa._apgd_step(norm, x, x_adv, x_old, g, step, eps, a, oa._l2_unit)
        fn = "rvlm_apgd_step_linf" if norm == "Linf" else "rvlm_apgd_step_l2"
        x_d, g_d, s_d = x.cuda(), g.cuda(), step.flatten().contiguous().cuda()  # keep alive
        args = [nat.ptr(xa_d), nat.ptr(xo_d), nat.ptr(x_d), nat.ptr(g_d),
                nat.ptr(s_d), nat.c_int(B), nat.c_int64(n),
                nat.c_float(eps), nat.c_float(a)]
        if norm == "L2":
            args.append(nat.ptr(ws))
        nat.call(fn, *args, nat.stream_ptr())
        torch.cuda.synchronize()
        if norm == "Linf":
            assert torch.equal(xa_d.cpu(), ref), f"iter {it}"
        else:
            assert (xa_d.cpu() - ref).abs().max() < 2e-6
        assert torch.equal(xo_d.cpu(), ref_old)
        x_adv, x_old = ref, ref_old
        xa_d.copy_(ref)


def test_apgd_bookkeeping_matches_oracle():
    """Drive the oracle's _ApgdBook and the device kernel with the same scripted losses."""
    nat = _nat()
    torch.manual_seed(4)
    B, n, n_iter = 6, 3 * 16 * 16, 10
    shape = (B, 3, 16, 16)
    loss0 = torch.rand(B)
    book = oa._ApgdBook(n_iter, B, 3, loss0, "cpu")
    x_best = torch.rand(shape); grad_best = torch.randn(shape)
    x_best_adv = x_best.clone()
    step = torch.full((B, 1, 1, 1), 0.03)
    acc = torch.ones(B, dtype=torch.bool)
    d = lambda t: t.clone().cuda()  # noqa: E731
    xb_d, gb_d, xba_d, st_d = d(x_best), d(grad_best), d(x_best_adv), d(step.flatten())
    lsteps = torch.zeros(n_iter, B).cuda(); lbest = d(loss0); lbl = d(loss0)
    red = torch.ones(B).cuda(); acc_d = torch.ones(B, dtype=torch.int32).cuda()
    flags = torch.zeros(B, dtype=torch.int32).cuda()
    sched = set(oa.checkpoints(n_iter))
    k, kmin, kdecr = oa.apgd_schedule(n_iter)
    for i in range(n_iter):
        x_adv = torch.rand(shape); grad = torch.randn(shape)
        loss = torch.rand(B) * (1.0 + 0.2 * i) if i % 3 else torch.rand(B) * 0.1
        pred = torch.rand(B) > 0.5
        xa_d, g_d = d(x_adv), d(grad)
        # oracle
        acc = torch.min(acc, pred)
        x_best_adv = torch.where(pred.view(-1, 1, 1, 1), x_best_adv, x_adv)
        k_now = book.k
        is_ck = i in sched
        x_adv_r, grad_r, x_best, grad_best, step = book.update(i, loss, x_adv, grad, x_best,
                                                               grad_best, step)
        loss_d, pred_d = loss.cuda(), pred.int().cuda()  # keep alive
        nat.call("rvlm_apgd_bookkeep", nat.ptr(xa_d), nat.ptr(g_d), nat.ptr(xb_d), nat.ptr(gb_d),
                 nat.ptr(xba_d), nat.ptr(loss_d), nat.ptr(pred_d), nat.ptr(lsteps),
                 nat.ptr(lbest), nat.ptr(lbl), nat.ptr(red), nat.ptr(st_d), nat.ptr(acc_d),
                 nat.ptr(flags), nat.c_int(B), nat.c_int64(n), nat.c_int(n_iter), nat.c_int(i),
                 nat.c_int(k_now if is_ck else 0), nat.c_float(float(k_now * 0.75)),
                 nat.stream_ptr())
        torch.cuda.synchronize()
        assert torch.equal(xa_d.cpu(), x_adv_r), f"x_adv iter {i}"
        assert torch.equal(g_d.cpu(), grad_r), f"grad iter {i}"
        assert torch.equal(xb_d.cpu(), x_best), f"x_best iter {i}"
        assert torch.equal(gb_d.cpu(), grad_best)
        assert torch.equal(xba_d.cpu(), x_best_adv)
        assert torch.equal(st_d.cpu(), step.flatten()), f"step iter {i}"
        assert torch.equal(lbest.cpu(), book.loss_best)
        assert torch.equal(acc_d.cpu().bool(), acc)
