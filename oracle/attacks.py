"""CPU oracle of the adversarial-perturbation algorithms (TEST INFRASTRUCTURE).

Restates, with explicit per-sample state and masks instead of index lists:
  * momentum PGD                 train/pgd_train.py:5-68
  * gradient normalisation / projection   vlm_eval/attacks/utils.py:8-26
  * training APGD                train/apgd_train.py:125-373
  * evaluation APGD              autoattack/autopgd_base.py:89-582
  * FARE / TeCoA losses          train/adversarial_training_clip.py:260-274,495-528
The arithmetic uses the same torch primitives in the same order as the
reference, so on CPU the results are bit-identical to it (checked by
tests/test_oracle_vs_reference.py in the authoring container and by the golden
vectors under tests/golden/).  Only ell_inf / ell_2 are restated (SURVEY.md §8a:
the L1/L0 branches are secondary / dead code).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LINF_NAMES = ("inf", "linf", "Linf")            # vlm_eval/attacks/utils.py:9
L2_NAMES = (2, 2.0, "l2", "L2", "2")            # vlm_eval/attacks/utils.py:12


# --------------------------------------------------------------------------- #
# losses  (train/adversarial_training_clip.py:495-528)
# --------------------------------------------------------------------------- #
def l2(out, targets, reduction="none"):
    """Squared ell_2 distance per row, NOT divided by the latent dimension (:509-521)."""
    assert out.shape == targets.shape, f"{out.shape} != {targets.shape}"
    assert out.shape[0] > 1
    per_row = F.mse_loss(out, targets, reduction="none").sum(dim=1)
    return per_row.mean() if reduction == "mean" else per_row


def ce(out, targets, reduction="mean"):
    """Cross entropy on logits (:523-528)."""
    assert out.shape[0] == targets.shape[0], (out.shape, targets.shape)
    assert out.shape[0] > 1
    return F.cross_entropy(out, targets, reduction=reduction)


def compute_loss(loss_str, embedding, targets, embedding_orig, logit_scale,
                 embedding_text_labels_norm=None, reduction="mean"):
    """:495-507"""
    if loss_str == "l2":
        return l2(out=embedding, targets=embedding_orig, reduction=reduction)
    if loss_str == "ce":
        return ce(out=embedding @ (logit_scale * embedding_text_labels_norm), targets=targets,
                  reduction=reduction)
    raise ValueError(f"loss {loss_str} not supported")


class ComputeLossWrapper:
    """:260-274"""

    def __init__(self, embedding_orig, embedding_text_labels_norm, reduction="mean", loss=None,
                 logit_scale=100.0):
        self.embedding_orig = embedding_orig
        self.embedding_text_labels_norm = embedding_text_labels_norm
        self.reduction = reduction
        self.loss_str = loss
        self.logit_scale = logit_scale

    def __call__(self, embedding, targets):
        return compute_loss(self.loss_str, embedding, targets, self.embedding_orig,
                            self.logit_scale, self.embedding_text_labels_norm, self.reduction)


@torch.no_grad()
def compute_acc(logits, targets):
    """:488-492"""
    return (logits.max(dim=1)[1].eq(targets).sum() / targets.shape[0]).item() * 100


# --------------------------------------------------------------------------- #
# PGD primitives (vlm_eval/attacks/utils.py:8-26)
# --------------------------------------------------------------------------- #
def project_perturbation(perturbation, eps, norm):
    if norm in LINF_NAMES:
        return torch.clamp(perturbation, -eps, eps)
    if norm in L2_NAMES:
        return torch.renorm(perturbation, p=2, dim=0, maxnorm=eps)
    raise NotImplementedError(f"Norm {norm} not supported")


def normalize_grad(grad, p):
    if p in LINF_NAMES:
        return grad.sign()
    if p in L2_NAMES:
        flat = grad.view(grad.shape[0], -1)
        return F.normalize(flat, p=2, dim=1).view_as(grad)
    return None  # the reference falls through and returns None for unknown norms (:19-26)


def _in_image_range(t):
    return bool(torch.max(t) < 1.0 + 1e-6) and bool(torch.min(t) > -1e-6)


def pgd(forward, loss_fn, data_clean, targets, norm, eps, iterations, stepsize, output_normalize,
        perturbation=None, mode="min", momentum=0.9, verbose=False, trace=None):
    """Momentum PGD (train/pgd_train.py:5-68).  Returns data_clean + delta (detached).

    `trace`, if a list, receives per-iteration dicts {grad, delta, loss} for tests."""
    assert _in_image_range(data_clean)                                     # :24
    delta = torch.zeros_like(data_clean) if perturbation is None else perturbation
    velocity = torch.zeros_like(data_clean)                                # :28
    if mode not in ("min", "max"):
        if iterations > 0:
            raise ValueError(f"Unknown mode: {mode}")                      # :54
    for it in range(iterations):
        delta = delta.detach().requires_grad_(True)                        # :30
        with torch.enable_grad():
            out = forward(data_clean + delta, output_normalize=output_normalize)  # :32
            loss = loss_fn(out, targets)                                   # :33
        if verbose:
            print(f"[{it}] {loss.item():.5f}")
        with torch.no_grad():
            g = torch.autograd.grad(loss, delta)[0]                        # :38
            nan = g.isnan()
            if nan.any():                                                  # :40-42
                print(f"attention: nan in gradient ({nan.sum()})")
                g[nan] = 0.0
            g = normalize_grad(g, p=norm)                                  # :44
            velocity = normalize_grad(momentum * velocity + g, p=norm)     # :46-47
            delta = delta - stepsize * velocity if mode == "min" else delta + stepsize * velocity
            delta = project_perturbation(delta, eps, norm)                 # :56
            delta = torch.clamp(data_clean + delta, 0, 1) - data_clean     # :57-59
            assert not delta.isnan().any()                                 # :60
            assert _in_image_range(data_clean + delta)                     # :61-63
            if trace is not None:
                trace.append({"loss": loss.detach().clone(), "delta": delta.clone()})
    return data_clean + delta.detach()                                     # :68


# --------------------------------------------------------------------------- #
# APGD shared pieces
# --------------------------------------------------------------------------- #
def _bview(v, ndims):
    return v.view(-1, *([1] * ndims))


def L2_norm(x, keepdim=False):
    """train/apgd_train.py:16-20 == autoattack/other_utils.py:38-42"""
    z = (x ** 2).view(x.shape[0], -1).sum(-1).sqrt()
    return _bview(z, x.dim() - 1) if keepdim else z


def check_oscillation(loss_steps, j, k, k3=0.75):
    """train/apgd_train.py:117-122 == autoattack/autopgd_base.py:170-175.

    Counts, over the last k iterations, how often the loss went up; negative row
    indices wrap exactly as Python indexing does in the reference (at the first
    check-point the index -1 reads the last, still zero, row)."""
    n = loss_steps.shape[0]
    t = torch.zeros(loss_steps.shape[1], device=loss_steps.device)
    for c in range(k):
        t += (loss_steps[(j - c) % n] > loss_steps[(j - c - 1) % n]).float()
    return (t <= k * k3 * torch.ones_like(t)).float()


def apgd_schedule(n_iter):
    """(k0, n_iter_min, size_decr): train/apgd_train.py:151-154, autopgd_base.py:151-153."""
    return (max(int(0.22 * n_iter), 1), max(int(0.06 * n_iter), 1), max(int(0.03 * n_iter), 1))


def checkpoints(n_iter):
    """Iteration indices i at which the step-size check fires."""
    k, kmin, decr = apgd_schedule(n_iter)
    out, c = [], 0
    for i in range(n_iter):
        c += 1
        if c == k:
            out.append(i)
            c = 0
            k = max(k - decr, kmin)
    return out


def _apgd_step(norm, x, x_adv, x_adv_old, grad, step_size, eps, a, normalize_fn,
               scale_then_divide=False):
    """One APGD proposal (train/apgd_train.py:205-254 == autopgd_base.py:328-350).
    Returns (x_adv_new, x_adv_old_new).  The ell_2 ascent direction is written
    `step_size * grad / (||grad|| + 1e-12)` in train/apgd_train.py:232-235 and
    vlm_eval/attacks/apgd.py:259-260 (multiply first: `scale_then_divide`) but
    `step_size * normalize(grad)` in autopgd_base.py:343 -- one rounding apart unless the step
    size is a power of two."""
    grad2 = x_adv - x_adv_old
    x_adv_old = x_adv.clone()
    if norm == "Linf":
        z = x_adv + step_size * torch.sign(grad)
        z = torch.clamp(torch.min(torch.max(z, x - eps), x + eps), 0.0, 1.0)
        z = torch.clamp(torch.min(torch.max(x_adv + (z - x_adv) * a + grad2 * (1 - a), x - eps),
                                  x + eps), 0.0, 1.0)
    elif norm == "L2":
        if scale_then_divide:
            z = x_adv + step_size * grad / (L2_norm(grad, keepdim=True) + 1e-12)
        else:
            z = x_adv + step_size * normalize_fn(grad)
        z = torch.clamp(x + normalize_fn(z - x) * torch.min(eps * torch.ones_like(x),
                                                            L2_norm(z - x, keepdim=True)), 0.0, 1.0)
        z = x_adv + (z - x_adv) * a + grad2 * (1 - a)
        z = torch.clamp(x + normalize_fn(z - x) * torch.min(eps * torch.ones_like(x),
                                                            L2_norm(z - x, keepdim=True)), 0.0, 1.0)
    else:
        raise NotImplementedError(norm)
    return z + 0.0, x_adv_old


def _l2_unit(v):
    """v / (||v||_2 + 1e-12) per sample (train/apgd_train.py:232-235; autopgd_base.py:180-193)."""
    return v / (L2_norm(v, keepdim=True) + 1e-12)


class _ApgdBook:
    """Best-point / step-size bookkeeping shared by both APGD variants
    (train/apgd_train.py:318-355 == autopgd_base.py:400-446)."""

    def __init__(self, n_iter, bs, ndims, loss0, device, thr_decr=0.75, schedule_n_iter=None):
        # the reference fixes the check-point schedule in the constructor (autopgd_base.py:157-159)
        # and `decr_eps_pgd` later changes n_iter only: schedule_n_iter = the constructor's n_iter
        self.k, self.k_min, self.k_decr = apgd_schedule(schedule_n_iter or n_iter)
        self.counter = 0
        self.thr = thr_decr
        self.ndims = ndims
        self.loss_steps = torch.zeros([n_iter, bs], device=device)
        self.loss_best = loss0.detach().clone()
        self.loss_best_last = self.loss_best.clone()
        self.reduced_last = torch.ones_like(self.loss_best)

    def update(self, i, loss_indiv, x_adv, grad, x_best, grad_best, step_size):
        y1 = loss_indiv.detach().clone()
        self.loss_steps[i] = y1 + 0
        better = _bview(y1 > self.loss_best, self.ndims)
        x_best = torch.where(better, x_adv, x_best)
        grad_best = torch.where(better, grad, grad_best)
        self.loss_best = torch.where(y1 > self.loss_best, y1, self.loss_best)
        self.counter += 1
        if self.counter == self.k:
            osc = check_oscillation(self.loss_steps, i, self.k, self.thr)
            no_impr = (1.0 - self.reduced_last) * (self.loss_best_last >= self.loss_best).float()
            flag = torch.max(osc, no_impr)
            self.reduced_last = flag.clone()
            self.loss_best_last = self.loss_best.clone()
            if flag.sum() > 0:
                m = _bview(flag > 0, self.ndims)
                step_size = torch.where(m, step_size / 2.0, step_size)
                x_adv = torch.where(m, x_best, x_adv)
                grad = torch.where(m, grad_best, grad)
            self.counter = 0
            self.k = max(self.k - self.k_decr, self.k_min)
        return x_adv, grad, x_best, grad_best, step_size


def apgd_train(model, x, y, norm, eps, n_iter=10, use_rs=False, loss_fn=None, verbose=False,
               is_train=True, initial_stepsize=None, trace=None):
    """Training-time APGD (train/apgd_train.py:125-373): no random start, last backward
    skipped, returns x_best_adv (Appendix C.1 of SURVEY.md)."""
    assert not model.training                                              # :127
    norm = norm.replace("linf", "Linf").replace("l2", "L2")               # :128
    if use_rs:
        raise NotImplementedError                                          # :135
    if norm not in ("Linf", "L2"):
        raise NotImplementedError(norm)
    ndims = x.dim() - 1
    bs = x.shape[0]
    x_adv = x.clone().clamp(0.0, 1.0)                                      # :132,139
    x_best = x_adv.clone()
    x_best_adv = x_adv.clone()
    alpha = 2.0 if not initial_stepsize else initial_stepsize / eps        # :158,168-169
    step_size = alpha * eps * torch.ones([bs, *[1] * ndims], device=x.device)

    def fwd_bwd(xa, need_grad=True):
        xa = xa.detach().requires_grad_(True)
        with torch.enable_grad():
            logits = model(xa, output_normalize=True)                      # :181,288
            loss_indiv = loss_fn(logits, y)
            loss = loss_indiv.sum()
        g = torch.autograd.grad(loss, [xa])[0].detach() if need_grad else None
        return logits.detach(), loss_indiv.detach(), g

    logits, loss_indiv, grad = fwd_bwd(x_adv)                              # :177-190
    grad_best = grad.clone()
    acc = logits.max(1)[1] == y
    book = _ApgdBook(n_iter, bs, ndims, loss_indiv, x.device)
    x_adv_old = x_adv.clone()

    for i in range(n_iter):                                                # :202
        a = 0.75 if i > 0 else 1.0
        x_adv, x_adv_old = _apgd_step(norm, x, x_adv.detach(), x_adv_old, grad, step_size, eps, a,
                                      _l2_unit, scale_then_divide=True)
        logits, loss_indiv, g_new = fwd_bwd(x_adv, need_grad=i < n_iter - 1)  # :288-295
        if g_new is not None:
            grad = g_new
        pred = logits.max(1)[1] == y                                       # :301
        acc = torch.min(acc, pred)
        x_best_adv = torch.where(_bview(~pred, ndims), x_adv, x_best_adv)  # :304-305
        x_adv, grad, x_best, grad_best, step_size = book.update(
            i, loss_indiv, x_adv, grad, x_best, grad_best, step_size)
        if trace is not None:
            trace.append({"x_adv": x_adv.clone(), "step_size": step_size.flatten().clone(),
                          "loss": loss_indiv.clone(), "loss_best": book.loss_best.clone()})
    return x_best_adv                                                      # :373


def dlr_loss(x, y):
    """autoattack/autopgd_base.py:195-201"""
    x_sorted, ind_sorted = x.sort(dim=1)
    ind = (ind_sorted[:, -1] == y).float()
    u = torch.arange(x.shape[0])
    return -(x[u, y] - x_sorted[:, -2] * ind - x_sorted[:, -1] * (1.0 - ind)) / (
        x_sorted[:, -1] - x_sorted[:, -3] + 1e-12)


class APGDAttack:
    """Evaluation APGD (autoattack/autopgd_base.py:89-582), ell_inf / ell_2, CE / DLR."""

    def __init__(self, predict, n_iter=100, norm="Linf", n_restarts=1, eps=None, seed=0, loss="ce",
                 eot_iter=1, rho=0.75, topk=None, verbose=False, device=None, use_largereps=False,
                 is_tf_model=False, logger=None, alpha=None, use_rs=True):
        assert norm in ("Linf", "L2") and eps is not None
        assert not is_tf_model
        self.model, self.n_iter, self.eps, self.norm = predict, n_iter, eps, norm
        self.n_restarts, self.seed, self.loss, self.eot_iter = n_restarts, seed, loss, eot_iter
        self.thr_decr, self.verbose, self.device = rho, verbose, device
        self.use_rs, self.alpha = use_rs, alpha
        self.use_largereps, self.n_iter_orig, self.eps_orig = use_largereps, n_iter, eps   # :143-145

    def normalize(self, x):
        """:180-193"""
        if self.norm == "Linf":
            t = x.abs().view(x.shape[0], -1).max(1)[0]
        else:
            t = (x ** 2).view(x.shape[0], -1).sum(-1).sqrt()
        return x / (_bview(t, x.dim() - 1) + 1e-12)

    def _criterion(self):
        if self.loss == "ce":
            return torch.nn.CrossEntropyLoss(reduction="none")             # :244
        if self.loss == "dlr":
            return dlr_loss
        if self.loss == "dlr-targeted":
            return self.dlr_loss_targeted                                  # :255-256
        if self.loss == "ce-targeted":
            return self.ce_loss_targeted
        raise ValueError("unknowkn loss")

    def _fwd_bwd(self, x_adv, y, crit):
        """EOT-averaged loss gradient (:267-286, 369-386)."""
        x_adv = x_adv.detach().requires_grad_(True)
        grad = torch.zeros_like(x_adv)
        for _ in range(self.eot_iter):
            with torch.enable_grad():
                logits = self.model(x_adv)
                loss_indiv = crit(logits, y)
                loss = loss_indiv.sum()
            grad += torch.autograd.grad(loss, [x_adv])[0].detach()
        grad /= float(self.eot_iter)
        return logits.detach(), loss_indiv.detach(), grad

    def attack_single_run(self, x, y, x_init=None):
        """:205-451 -> (x_best, acc, loss_best, x_best_adv)"""
        dev = x.device
        ndims = x.dim() - 1
        if self.use_rs:                                                    # :210-218 (CPU RNG)
            if self.norm == "Linf":
                t = 2 * torch.rand(x.shape).to(dev).detach() - 1
            else:
                t = torch.randn(x.shape).to(dev).detach()
            x_adv = x + self.eps * torch.ones_like(x).detach() * self.normalize(t)
        else:
            x_adv = x.clone()
        if x_init is not None:
            x_adv = x_init.clone()
        x_adv = x_adv.clamp(0.0, 1.0)                                      # :233
        x_best = x_adv.clone()
        x_best_adv = x_adv.clone()
        crit = self._criterion()
        logits, loss_indiv, grad = self._fwd_bwd(x_adv, y, crit)
        grad_best = grad.clone()
        acc = logits.max(1)[1] == y
        alpha = 2.0 if self.alpha is None else self.alpha                  # :295-299
        step_size = alpha * self.eps * torch.ones([x.shape[0], *([1] * ndims)]).to(dev).detach()
        x_adv_old = x_adv.clone()
        book = _ApgdBook(self.n_iter, x.shape[0], ndims, loss_indiv, dev, self.thr_decr,
                         schedule_n_iter=self.n_iter_orig)
        for i in range(self.n_iter):                                       # :326
            a = 0.75 if i > 0 else 1.0
            with torch.no_grad():
                x_adv, x_adv_old = _apgd_step(self.norm, x, x_adv.detach(), x_adv_old, grad,
                                              step_size, self.eps, a, self.normalize)
            logits, loss_indiv, grad = self._fwd_bwd(x_adv, y, crit)       # grad every iter
            pred = logits.max(1)[1] == y
            acc = torch.min(acc, pred)
            x_best_adv = torch.where(_bview(~pred, ndims), x_adv, x_best_adv)
            with torch.no_grad():
                x_adv, grad, x_best, grad_best, step_size = book.update(
                    i, loss_indiv, x_adv, grad, x_best, grad_best, step_size)
        return x_best, acc, book.loss_best, x_best_adv                     # :451

    def decr_eps_pgd(self, x, y, epss, iters, use_rs=True):
        """:550-582 (Linf branch)"""
        assert len(epss) == len(iters)
        assert self.norm in ("L1", "Linf")
        self.use_rs = False
        if not use_rs:
            x_init = None
        else:
            x_init = x + torch.randn_like(x)
            x_init = torch.clamp(x_init, 0.0, 1.0)
            x_init = torch.min(torch.max(x_init, x - self.eps), x + self.eps)
        for eps, niter in zip(epss, iters):
            self.n_iter = niter + 0
            self.eps = eps + 0.0
            if x_init is not None:
                x_init = torch.clamp(x_init, 0.0, 1.0)
                x_init = torch.min(torch.max(x_init, x - self.eps), x + self.eps)
            x_init, acc, loss, x_adv = self.attack_single_run(x, y, x_init=x_init)
        return x_init, acc, loss, x_adv

    def _largereps_schedule(self):
        """:493-498"""
        import math
        epss = [3.0 * self.eps_orig, 2.0 * self.eps_orig, 1.0 * self.eps_orig]
        iters = [math.ceil(c) for c in (0.3 * self.n_iter_orig, 0.3 * self.n_iter_orig,
                                        0.4 * self.n_iter_orig)]
        iters[-1] = self.n_iter_orig - sum(iters[:-1])
        return epss, iters

    def perturb(self, x, y=None, best_loss=False, x_init=None):
        """:453-548"""
        assert self.loss in ("ce", "dlr")
        if self.device is None:
            self.device = x.device
        x = x.detach().clone().float().to(self.device)
        y_pred = self.model(x).max(1)[1]                                   # :469
        y = y_pred.detach().clone().long() if y is None else y.detach().clone().long().to(self.device)
        adv = x.clone()
        acc = y_pred == y                                                  # :480
        if best_loss:
            adv_best = x.detach().clone()
            loss_best = torch.ones([x.shape[0]]).to(self.device) * (-float("inf"))
            for _ in range(self.n_restarts):
                best_curr, _, loss_curr, _ = self.attack_single_run(x, y)
                m = loss_curr > loss_best
                adv_best = torch.where(_bview(m, x.dim() - 1), best_curr, adv_best)
                loss_best = torch.where(m, loss_curr, loss_best)
            return adv_best
        if self.use_largereps:
            epss, iters = self._largereps_schedule()
        torch.random.manual_seed(self.seed)                                # :505
        for _ in range(self.n_restarts):                                   # :508
            idx = acc.nonzero().flatten()
            if idx.numel() == 0:
                continue
            if not self.use_largereps:                                     # :517-520
                res = self.attack_single_run(x[idx].clone(), y[idx].clone())
            else:
                res = self.decr_eps_pgd(x[idx].clone(), y[idx].clone(), epss, iters)
            _, acc_curr, _, adv_curr = res
            fooled = (acc_curr == 0).nonzero().flatten()
            acc[idx[fooled]] = 0                                           # :521-525
            adv[idx[fooled]] = adv_curr[fooled].clone()
        return adv


class APGDAttack_targeted(APGDAttack):
    """AutoPGD on the targeted DLR loss (autoattack/autopgd_base.py:584-707): for every target
    rank 2..n_target_classes+1 (by clean-logit order of the still-robust points) run
    n_restarts attacks towards that class."""

    def __init__(self, predict, n_iter=100, norm="Linf", n_restarts=1, eps=None, seed=0,
                 eot_iter=1, rho=0.75, topk=None, n_target_classes=9, verbose=False, device=None,
                 use_largereps=False, is_tf_model=False, logger=None, alpha=None, use_rs=True):
        super().__init__(predict, n_iter=n_iter, norm=norm, n_restarts=n_restarts, eps=eps,
                         seed=seed, loss="dlr-targeted", eot_iter=eot_iter, rho=rho, topk=topk,
                         verbose=verbose, device=device, use_largereps=use_largereps,
                         is_tf_model=is_tf_model, logger=logger, alpha=alpha, use_rs=use_rs)
        self.y_target = None
        self.n_target_classes = n_target_classes

    def dlr_loss_targeted(self, x, y):
        """:616-621"""
        x_sorted, _ = x.sort(dim=1)
        u = torch.arange(x.shape[0])
        return -(x[u, y] - x[u, self.y_target]) / (
            x_sorted[:, -1] - 0.5 * (x_sorted[:, -3] + x_sorted[:, -4]) + 1e-12)

    def ce_loss_targeted(self, x, y):
        """:623-624"""
        return -1.0 * F.cross_entropy(x, self.y_target, reduction="none")

    def perturb(self, x, y=None, x_init=None):
        """:627-707"""
        assert self.loss in ("dlr-targeted",)
        if self.device is None:
            self.device = x.device
        x = x.detach().clone().float().to(self.device)
        y_pred = self.model(x).max(1)[1]
        y = y_pred.detach().clone().long() if y is None else y.detach().clone().long().to(self.device)
        adv = x.clone()
        acc = y_pred == y
        torch.random.manual_seed(self.seed)                                # :661
        for target_class in range(2, self.n_target_classes + 2):           # :678
            for _ in range(self.n_restarts):
                idx = acc.nonzero().flatten()
                if idx.numel() == 0:
                    continue
                x_to_fool, y_to_fool = x[idx].clone(), y[idx].clone()
                output = self.model(x_to_fool)
                self.y_target = output.sort(dim=1)[1][:, -target_class]    # :691
                _, acc_curr, _, adv_curr = self.attack_single_run(x_to_fool, y_to_fool)
                fooled = (acc_curr == 0).nonzero().flatten()
                acc[idx[fooled]] = 0
                adv[idx[fooled]] = adv_curr[fooled].clone()
        return adv
