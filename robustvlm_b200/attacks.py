"""Attack-side call surface of the reference, on CUDA kernels.

  pgd                  train/pgd_train.py:5-68            (same signature)
  apgd_train           train/apgd_train.py:125-373        (same signature)
  APGDAttack           autoattack/autopgd_base.py:89-582  (same ctor / perturb / attack_single_run)
  APGDAttack_targeted  autoattack/autopgd_base.py:584-707 (targeted DLR, n_target_classes)
  project_perturbation, normalize_grad   vlm_eval/attacks/utils.py:8-26

Two execution paths behind each entry point, both on the GPU and both ending in the
hand-written update kernels (there is no CPU / eager fallback -- CPU tensors raise):

  fused    `forward` is a robustvlm_b200 ClipVisionModel around our VisionTransformer
           and the loss is a ComputeLossWrapper with loss in {l2, ce} (the FARE / TeCoA
           configurations): the whole model evaluation + loss + input gradient runs in
           the C engine; for pgd the complete loop is one C call (rvlm_pgd_run) with
           the delta-update fused into the last backward kernel.
  generic  any differentiable `forward` / `loss_fn`: torch.autograd supplies the
           gradient (through the engine's VJP when the tower is ours), the update,
           projection, APGD step and bookkeeping run in the C kernels.
"""
from __future__ import annotations

import math

import torch

from . import _native as nat
from ._native import c_float, c_int, c_int64
from .engine import NORM_L2, NORM_LINF, LossSpec
from .model import ClassificationModel, ClipVisionModel, ComputeLossWrapper, VisionTransformer

LINF_NAMES = ("inf", "linf", "Linf")   # vlm_eval/attacks/utils.py:9
L2_NAMES = (2, 2.0, "l2", "L2", "2")   # vlm_eval/attacks/utils.py:12


def _norm_id(norm):
    if norm in LINF_NAMES:
        return NORM_LINF
    if norm in L2_NAMES:
        return NORM_L2
    raise NotImplementedError(f"Norm {norm} not supported")


def _need_cuda(t, what):
    if not t.is_cuda:
        raise nat.RvlmError(f"{what} must be a CUDA tensor: robustvlm_b200 has no CPU path")


def _flat_n(t):
    return math.prod(t.shape[1:])


def _upd_ws(B, device):
    fn = nat.lib().rvlm_update_workspace_bytes
    fn.restype = nat.c_size_t
    return torch.empty(int(fn(c_int(B))), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------------------
# primitives (kept for API parity; both are thin wrappers over the update kernels' math)
# ----------------------------------------------------------------------------------------
def project_perturbation(perturbation, eps, norm):
    if _norm_id(norm) == NORM_LINF:
        return torch.clamp(perturbation, -eps, eps)
    return torch.renorm(perturbation, p=2, dim=0, maxnorm=eps)


def normalize_grad(grad, p):
    if _norm_id(p) == NORM_LINF:
        return grad.sign()
    return torch.nn.functional.normalize(grad.view(grad.shape[0], -1), p=2, dim=1).view_as(grad)


def _fusable(forward, loss_fn):
    return (isinstance(forward, ClipVisionModel) and isinstance(forward.model, VisionTransformer)
            and isinstance(loss_fn, ComputeLossWrapper) and loss_fn.loss_str in ("l2", "ce"))


def _loss_spec(loss_fn: ComputeLossWrapper, targets, output_normalize, pred_on_logits=False):
    return LossSpec(loss_fn.loss_str, output_normalize, emb_orig=loss_fn.embedding_orig,
                    text=loss_fn.embedding_text_labels_norm, logit_scale=loss_fn.logit_scale,
                    targets=targets, pred_on_logits=pred_on_logits)


# ----------------------------------------------------------------------------------------
# PGD
# ----------------------------------------------------------------------------------------
def pgd(forward, loss_fn, data_clean, targets, norm, eps, iterations, stepsize, output_normalize,
        perturbation=None, mode="min", momentum=0.9, verbose=False, chunk=None):
    """Minimize or maximize the given loss (train/pgd_train.py:5-68); returns
    data_clean + perturbation (detached)."""
    _need_cuda(data_clean, "data_clean")
    assert torch.max(data_clean) < 1.0 + 1e-6 and torch.min(data_clean) > -1e-6   # :24
    if mode not in ("min", "max"):
        raise ValueError(f"Unknown mode: {mode}")                                  # :54
    nid = _norm_id(norm)
    x = data_clean.detach().float().contiguous()
    if _fusable(forward, loss_fn) and not verbose:
        eng = forward.engine()
        spec = _loss_spec(loss_fn, targets, output_normalize)
        d0 = perturbation.detach().float().contiguous() if perturbation is not None else None
        return eng.pgd_run(x, spec, norm=nid, eps=float(eps), stepsize=float(stepsize),
                           iterations=int(iterations), maximize=(mode == "max"),
                           momentum=float(momentum), delta0=d0, chunk=chunk)
    # ---- generic path: autograd gradient + CUDA update kernels
    B, n = x.shape[0], _flat_n(x)
    delta = (torch.zeros_like(x) if perturbation is None
             else perturbation.detach().float().clone().contiguous())
    velocity = torch.zeros_like(x)                                                 # :28
    ws = _upd_ws(B, x.device) if nid == NORM_L2 else None
    for i in range(iterations):
        delta.requires_grad_(True)
        with torch.enable_grad():
            out = forward(x + delta, output_normalize=output_normalize)            # :32
            loss = loss_fn(out, targets)                                           # :33
            if verbose:
                print(f"[{i}] {loss.item():.5f}")
        grad = torch.autograd.grad(loss, delta)[0].detach().float().contiguous()   # :38
        delta = delta.detach()
        if nid == NORM_LINF:
            nat.call("rvlm_pgd_update_linf", nat.ptr(delta), nat.ptr(velocity), nat.ptr(grad),
                     nat.ptr(x), c_int(B), c_int64(n), c_float(stepsize), c_float(eps),
                     c_int(int(mode == "max")), c_float(momentum), nat.stream_ptr())
        else:
            nat.call("rvlm_pgd_update_l2", nat.ptr(delta), nat.ptr(velocity), nat.ptr(grad),
                     nat.ptr(x), c_int(B), c_int64(n), c_float(stepsize), c_float(eps),
                     c_int(int(mode == "max")), c_float(momentum), nat.ptr(ws), nat.stream_ptr())
    out = torch.empty_like(x)
    nat.call("rvlm_add_f32", nat.ptr(out), nat.ptr(x), nat.ptr(delta), c_int64(x.numel()),
             nat.stream_ptr())
    return out                                                                     # :68


# ----------------------------------------------------------------------------------------
# APGD core (device resident; the check-point schedule is data independent, so the host
# never reads a device value inside the loop)
# ----------------------------------------------------------------------------------------
def apgd_schedule(n_iter):
    """(k0, n_iter_min, size_decr): train/apgd_train.py:151-154 == autopgd_base.py:151-153"""
    return (max(int(0.22 * n_iter), 1), max(int(0.06 * n_iter), 1), max(int(0.03 * n_iter), 1))


def _apgd_core(fwd_bwd, x, x_start, norm, eps, n_iter, alpha, thr_decr, skip_last_grad,
               schedule=None, round_to=None):
    """fwd_bwd(x_adv, need_grad) -> (loss_indiv[B] f32, pred[B] int32, grad | None).
    `schedule` overrides (k0, n_iter_min, size_decr) (vlm_eval/attacks/apgd.py:194-198).
    `round_to`: the iterate is rounded to this dtype after every step, as the reference's
    `x_adv = x_adv_1.to(dtype=x_adv.dtype)` does for fp16 inputs (vlm_eval/attacks/apgd.py:289).
    Returns (x_best, acc(bool), loss_best, x_best_adv)."""
    B, n, dev = x.shape[0], _flat_n(x), x.device
    x_adv = x_start.clamp(0.0, 1.0).contiguous()
    x_best, x_best_adv = x_adv.clone(), x_adv.clone()
    loss, pred, grad = fwd_bwd(x_adv, True)
    grad = grad.contiguous()
    grad_best = grad.clone()
    acc = pred.to(torch.int32).clone()
    loss_best = loss.clone()
    loss_best_last = loss.clone()
    reduced_last = torch.ones_like(loss)
    step_size = torch.full((B,), float(alpha * eps), device=dev, dtype=torch.float32)
    x_adv_old = x_adv.clone()
    loss_steps = torch.zeros(n_iter, B, device=dev, dtype=torch.float32)
    flags = torch.zeros(B, device=dev, dtype=torch.int32)
    ws = _upd_ws(B, dev) if norm == "L2" else None
    k, k_min, k_decr = apgd_schedule(n_iter) if schedule is None else schedule
    counter3 = 0
    for i in range(n_iter):
        a = 0.75 if i > 0 else 1.0
        if norm == "Linf":
            nat.call("rvlm_apgd_step_linf", nat.ptr(x_adv), nat.ptr(x_adv_old), nat.ptr(x),
                     nat.ptr(grad), nat.ptr(step_size), c_int(B), c_int64(n), c_float(eps),
                     c_float(a), nat.stream_ptr())
        else:
            nat.call("rvlm_apgd_step_l2", nat.ptr(x_adv), nat.ptr(x_adv_old), nat.ptr(x),
                     nat.ptr(grad), nat.ptr(step_size), c_int(B), c_int64(n), c_float(eps),
                     c_float(a), nat.ptr(ws), nat.stream_ptr())
        if round_to is not None:
            x_adv.copy_(x_adv.to(round_to))
        need_grad = not (skip_last_grad and i == n_iter - 1)
        loss, pred, g_new = fwd_bwd(x_adv, need_grad)
        if g_new is not None:
            grad.copy_(g_new)
        counter3 += 1
        is_ck = counter3 == k
        loss_c, pred_c = loss.contiguous(), pred.to(torch.int32).contiguous()   # held across the call
        nat.call("rvlm_apgd_bookkeep", nat.ptr(x_adv), nat.ptr(grad), nat.ptr(x_best),
                 nat.ptr(grad_best), nat.ptr(x_best_adv), nat.ptr(loss_c),
                 nat.ptr(pred_c), nat.ptr(loss_steps),
                 nat.ptr(loss_best), nat.ptr(loss_best_last), nat.ptr(reduced_last),
                 nat.ptr(step_size), nat.ptr(acc), nat.ptr(flags), c_int(B), c_int64(n),
                 c_int(n_iter), c_int(i), c_int(k if is_ck else 0), c_float(float(k * thr_decr)),
                 nat.stream_ptr())
        if is_ck:
            counter3 = 0
            k = max(k - k_decr, k_min)
    return x_best, acc.bool(), loss_best, x_best_adv


def _autograd_fwd_bwd(model_call, loss_call, y):
    def fwd_bwd(x_adv, need_grad):
        xa = x_adv.detach().clone().requires_grad_(need_grad)
        with torch.enable_grad():
            logits = model_call(xa)
            loss_indiv = loss_call(logits, y)
            g = (torch.autograd.grad(loss_indiv.sum(), [xa])[0].detach().float()
                 if need_grad else None)
        pred = (logits.detach().max(1)[1] == y).to(torch.int32)
        return loss_indiv.detach().float(), pred, g
    return fwd_bwd


def apgd_train(model, x, y, norm, eps, n_iter=10, use_rs=False, loss_fn=None, verbose=False,
               is_train=True, initial_stepsize=None):
    """Training-time APGD (train/apgd_train.py:125-373); returns x_best_adv."""
    assert not model.training                                                      # :127
    norm = norm.replace("linf", "Linf").replace("l2", "L2")                        # :128
    if use_rs:
        raise NotImplementedError                                                  # :135
    if norm not in ("Linf", "L2"):
        raise NotImplementedError(f"norm {norm}: only Linf / L2 are built (SURVEY.md §8a)")
    _need_cuda(x, "x")
    x = x.detach().float().contiguous()
    alpha = 2.0 if not initial_stepsize else initial_stepsize / eps                # :158,168
    if _fusable(model, loss_fn):
        eng = model.engine()
        spec = _loss_spec(loss_fn, y, True)     # apgd_train always normalises (:181,288)

        def fwd_bwd(x_adv, need_grad):
            r = eng.loss_grad(x_adv, spec, want_grad=need_grad)
            return r["loss"], r["pred"].to(torch.int32), r["grad"]
    else:
        fwd_bwd = _autograd_fwd_bwd(lambda xa: model(xa, output_normalize=True), loss_fn, y)
    _, _, _, x_best_adv = _apgd_core(fwd_bwd, x, x.clone(), norm, float(eps), n_iter, alpha, 0.75,
                                     skip_last_grad=True)
    return x_best_adv                                                              # :373


def dlr_loss(x, y):
    """autoattack/autopgd_base.py:195-201"""
    x_sorted, ind_sorted = x.sort(dim=1)
    ind = (ind_sorted[:, -1] == y).float()
    u = torch.arange(x.shape[0], device=x.device)
    return -(x[u, y] - x_sorted[:, -2] * ind - x_sorted[:, -1] * (1.0 - ind)) / (
        x_sorted[:, -1] - x_sorted[:, -3] + 1e-12)


class APGDAttack:
    """AutoPGD (autoattack/autopgd_base.py:89-582): ell_inf / ell_2, CE / DLR, restarts over
    the still-robust points."""

    def __init__(self, predict, n_iter=100, norm="Linf", n_restarts=1, eps=None, seed=0, loss="ce",
                 eot_iter=1, rho=0.75, topk=None, verbose=False, device=None, use_largereps=False,
                 is_tf_model=False, logger=None, alpha=None, use_rs=True):
        assert norm in ("Linf", "L2", "L1") and eps is not None                    # :146-147
        if norm == "L1" or is_tf_model:
            raise NotImplementedError("only Linf / L2, PyTorch models (SURVEY.md §8a)")
        if use_largereps and norm != "Linf":
            raise NotImplementedError("use_largereps: Linf only (the reference asserts L1 / Linf, :552)")
        self.model, self.n_iter, self.eps, self.norm = predict, n_iter, eps, norm
        self.n_restarts, self.seed, self.loss, self.eot_iter = n_restarts, seed, loss, eot_iter
        self.thr_decr, self.topk, self.verbose, self.device = rho, topk, verbose, device
        self.use_rs, self.use_largereps, self.alpha, self.logger = use_rs, use_largereps, alpha, logger
        self.n_iter_orig, self.eps_orig, self.y_target = n_iter, eps, None

    def normalize(self, x):
        """:180-193"""
        if self.norm == "Linf":
            t = x.abs().view(x.shape[0], -1).max(1)[0]
        else:
            t = (x ** 2).view(x.shape[0], -1).sum(-1).sqrt()
        return x / (t.view(-1, *([1] * (x.dim() - 1))) + 1e-12)

    def _fused_ok(self):
        m = self.model
        return (self.loss in ("ce", "dlr", "dlr-targeted") and isinstance(m, ClassificationModel)
                and isinstance(m.visual.model, VisionTransformer) and m.resizer is None)

    def _fwd_bwd(self, y):
        m = self.model
        if self._fused_ok():
            # model + loss + input gradient inside the engine: CE, DLR (:195-201) and targeted DLR
            # (:616-621) are evaluated by the head kernel (top-4 logits on device), no autograd
            eng = m.visual.engine()
            spec = LossSpec(self.loss, True, text=m.text_embedding,
                            logit_scale=m.logit_scale_value, targets=y, pred_on_logits=True,
                            y_target=self.y_target if self.loss == "dlr-targeted" else None)

            def fused(x_adv, need_grad):
                r = eng.loss_grad(x_adv, spec, want_grad=need_grad)
                return r["loss"], r["pred"].to(torch.int32), r["grad"]
            base = fused
        else:
            if self.loss == "ce":
                crit = torch.nn.CrossEntropyLoss(reduction="none")                 # :244
            elif self.loss == "dlr":
                crit = dlr_loss
            elif self.loss == "dlr-targeted":
                crit = self.dlr_loss_targeted                                      # :255-256
            elif self.loss == "ce-targeted":
                crit = self.ce_loss_targeted
            else:
                raise ValueError("unknowkn loss")
            base = _autograd_fwd_bwd(m, crit, y)
        if self.eot_iter == 1:
            return base

        def eot(x_adv, need_grad):                                                 # :267-286
            # the reference evaluates the model eot_iter times and averages the gradients (for
            # randomised defences); the loss / prediction reported are those of the last pass
            acc_g, out = None, None
            for _ in range(self.eot_iter):
                out = base(x_adv, need_grad)
                if need_grad:
                    acc_g = out[2].clone() if acc_g is None else acc_g.add_(out[2])
            return out[0], out[1], (acc_g.div_(float(self.eot_iter)) if need_grad else None)
        return eot

    def decr_eps_pgd(self, x, y, epss, iters, use_rs=True):
        """:550-582 -- the `use_largereps` schedule: APGD runs at 3 eps, 2 eps, eps chained through
        x_init.  Quirks kept: `self.use_rs` stays False afterwards; the check-point schedule is
        the one of the ORIGINAL n_iter (set once in the constructor, :157-159)."""
        assert len(epss) == len(iters)
        assert self.norm in ("L1", "Linf")
        self.use_rs = False
        if not use_rs:
            x_init = None
        else:
            x_init = x + torch.randn_like(x)
            x_init = torch.clamp(x_init, 0.0, 1.0)
            x_init = torch.min(torch.max(x_init, x - self.eps), x + self.eps)
        res = None
        for eps, niter in zip(epss, iters):
            self.n_iter = niter + 0
            self.eps = eps + 0.0
            if x_init is not None:
                x_init = torch.clamp(x_init, 0.0, 1.0)
                x_init = torch.min(torch.max(x_init, x - self.eps), x + self.eps)
            res = self.attack_single_run(x, y, x_init=x_init)
            x_init = res[0]
        return res

    def _largereps_schedule(self):
        """:493-498"""
        epss = [3.0 * self.eps_orig, 2.0 * self.eps_orig, 1.0 * self.eps_orig]
        iters = [math.ceil(c) for c in (0.3 * self.n_iter_orig, 0.3 * self.n_iter_orig,
                                        0.4 * self.n_iter_orig)]
        iters[-1] = self.n_iter_orig - sum(iters[:-1])
        return epss, iters

    def attack_single_run(self, x, y, x_init=None):
        """:205-451 -> (x_best, acc, loss_best, x_best_adv)"""
        _need_cuda(x, "x")
        x = x.detach().float().contiguous()
        if self.use_rs:                                                            # :210-218
            if self.norm == "Linf":
                t = 2 * torch.rand(x.shape).to(x.device).detach() - 1               # CPU RNG
            else:
                t = torch.randn(x.shape).to(x.device).detach()
            x_adv = x + self.eps * torch.ones_like(x).detach() * self.normalize(t)
        else:
            x_adv = x.clone()
        if x_init is not None:
            x_adv = x_init.clone()
        alpha = 2.0 if self.alpha is None else self.alpha                          # :295-299
        return _apgd_core(self._fwd_bwd(y), x, x_adv, self.norm, float(self.eps), self.n_iter,
                          alpha, self.thr_decr, skip_last_grad=False,
                          schedule=apgd_schedule(self.n_iter_orig))                # :157-159

    def perturb(self, x, y=None, best_loss=False, x_init=None):
        """:453-548"""
        assert self.loss in ("ce", "dlr")
        _need_cuda(x, "x")
        if y is not None and y.dim() == 0:                                         # :462-464
            x.unsqueeze_(0)
            y.unsqueeze_(0)
        if self.device is None:
            self.device = x.device
        x = x.detach().clone().float().to(self.device)
        with torch.no_grad():
            y_pred = self.model(x).max(1)[1]                                       # :469
        y = y_pred.detach().clone().long() if y is None else y.detach().clone().long().to(self.device)
        adv = x.clone()
        acc = y_pred == y                                                          # :480
        if best_loss:
            adv_best = x.detach().clone()
            loss_best = torch.full((x.shape[0],), -float("inf"), device=self.device)
            for _ in range(self.n_restarts):
                best_curr, _, loss_curr, _ = self.attack_single_run(x, y)
                m = loss_curr > loss_best
                adv_best[m] = best_curr[m]
                loss_best[m] = loss_curr[m]
            return adv_best
        if self.use_largereps:                                                     # :493-498
            epss, iters = self._largereps_schedule()
        torch.random.manual_seed(self.seed)                                        # :505
        torch.cuda.random.manual_seed(self.seed)                                   # :506
        for _ in range(self.n_restarts):                                           # :508
            idx = acc.nonzero().flatten()          # the one host sync per restart (as :509)
            if idx.numel() == 0:
                continue
            if not self.use_largereps:                                             # :517-520
                res_curr = self.attack_single_run(x[idx].clone(), y[idx].clone())
            else:
                res_curr = self.decr_eps_pgd(x[idx].clone(), y[idx].clone(), epss, iters)
            _, acc_curr, _, adv_curr = res_curr
            fooled = (acc_curr == 0).nonzero().flatten()
            acc[idx[fooled]] = False                                               # :521-525
            adv[idx[fooled]] = adv_curr[fooled].clone()
        return adv


class APGDAttack_targeted(APGDAttack):
    """AutoPGD on the targeted DLR loss (autoattack/autopgd_base.py:584-707): for each target
    rank 2 .. n_target_classes+1 of the still-robust points' logits, n_restarts runs towards
    that class.  On our ClassificationModel the loss (top-4 logits) and its gradient are
    evaluated by the head kernel (RVLM_LOSS_DLR_TARGETED); step / projection / bookkeeping are
    the CUDA update kernels."""

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
        u = torch.arange(x.shape[0], device=x.device)
        return -(x[u, y] - x[u, self.y_target]) / (
            x_sorted[:, -1] - 0.5 * (x_sorted[:, -3] + x_sorted[:, -4]) + 1e-12)

    def ce_loss_targeted(self, x, y):
        """:623-624"""
        return -1.0 * torch.nn.functional.cross_entropy(x, self.y_target, reduction="none")

    def perturb(self, x, y=None, x_init=None):
        """:627-707"""
        assert self.loss in ("dlr-targeted",)
        _need_cuda(x, "x")
        if y is not None and y.dim() == 0:                                         # :634-636
            x.unsqueeze_(0)
            y.unsqueeze_(0)
        if self.device is None:
            self.device = x.device
        x = x.detach().clone().float().to(self.device)
        with torch.no_grad():
            y_pred = self.model(x).max(1)[1]
        y = y_pred.detach().clone().long() if y is None else y.detach().clone().long().to(self.device)
        adv = x.clone()
        acc = y_pred == y
        if self.use_largereps:                                                     # :666-671
            epss, iters = self._largereps_schedule()
        torch.random.manual_seed(self.seed)                                        # :661
        torch.cuda.random.manual_seed(self.seed)                                   # :662
        for target_class in range(2, self.n_target_classes + 2):                   # :678
            for _ in range(self.n_restarts):
                idx = acc.nonzero().flatten()
                if idx.numel() == 0:
                    continue
                x_to_fool, y_to_fool = x[idx].clone(), y[idx].clone()
                with torch.no_grad():
                    output = self.model(x_to_fool)
                self.y_target = output.sort(dim=1)[1][:, -target_class]            # :691
                if not self.use_largereps:
                    res_curr = self.attack_single_run(x_to_fool, y_to_fool)
                else:
                    res_curr = self.decr_eps_pgd(x_to_fool, y_to_fool, epss, iters)
                _, acc_curr, _, adv_curr = res_curr
                fooled = (acc_curr == 0).nonzero().flatten()
                acc[idx[fooled]] = False
                adv[idx[fooled]] = adv_curr[fooled].clone()
        return adv
