"""CPU oracle of the VLM-evaluation attacks (TEST INFRASTRUCTURE).

Restates, with masks instead of index lists:
  * `PGD`    vlm_eval/attacks/pgd.py:5-87     (momentum PGD with a frame mask, optional
             image-space clamp; the model callable returns the scalar loss itself)
  * `APGD`, `apgd`   vlm_eval/attacks/apgd.py:5-48,163-370  (training-APGD arithmetic with a
             gradient mask, `pert_init`, tunable check-point fractions, batch size 1,
             x_best_adv = the last iterate)
Same torch primitives in the same order as the reference, so on CPU the results are
bit-identical to it (tests/test_oracle_vs_reference.py, authoring container; golden vectors
tests/golden/reference_vectors_vlm.pt).  ell_inf / ell_2 only.
"""
from __future__ import annotations

import torch

from .attacks import (L2_norm, _ApgdBook, _apgd_step, _l2_unit, normalize_grad,
                      project_perturbation)


def set_mask(data, mask_out):
    """`_set_mask` (vlm_eval/attacks/pgd.py:75-87 == apgd.py:34-46): dim 1 indexes the frames
    of one sample (context images ..., query image last)."""
    mask = torch.ones_like(data)
    if mask_out is None:
        return mask
    if isinstance(mask_out, str):
        frames = {"context": slice(None, -1), "query": -1}.get(mask_out)
        if frames is None:
            raise NotImplementedError(f"Unknown mask_out: {mask_out}")
    elif isinstance(mask_out, int):
        frames = mask_out
    else:
        raise NotImplementedError(f"Unknown mask_out: {mask_out}")
    mask[:, frames] = 0
    return mask


class PGD:
    """vlm_eval/attacks/pgd.py:5-73"""

    def __init__(self, forward, norm, eps, mode="min", mask_out="context", image_space=True):
        self.model, self.norm, self.eps, self.momentum = forward, norm, eps, 0.9
        self.mode, self.mask_out, self.image_space = mode, mask_out, image_space

    def perturb(self, data_clean, iterations, stepsize, perturbation=None, verbose=False,
                return_loss=False):
        if self.image_space:
            assert torch.max(data_clean) < 1.0 + 1e-6 and torch.min(data_clean) > -1e-6   # :24
        if perturbation is None:
            perturbation = torch.zeros_like(data_clean, requires_grad=True)
        mask = set_mask(data_clean, self.mask_out)
        velocity = torch.zeros_like(data_clean)
        loss = None
        for _ in range(iterations):
            perturbation = perturbation.detach().requires_grad_(True)
            with torch.enable_grad():
                loss = self.model(data_clean + perturbation)                             # :33
            with torch.no_grad():
                g = torch.autograd.grad(loss, perturbation)[0]
                g = mask * g                                                             # :39
                nan = g.isnan()
                if nan.any():
                    g[nan] = 0.0                                                         # :40-42
                g = normalize_grad(g, p=self.norm)
                velocity = normalize_grad(self.momentum * velocity + g, p=self.norm)     # :46-47
                if self.mode == "min":
                    perturbation = perturbation - stepsize * velocity
                elif self.mode == "max":
                    perturbation = perturbation + stepsize * velocity
                else:
                    raise ValueError(f"Unknown mode: {self.mode}")
                perturbation = project_perturbation(perturbation, self.eps, self.norm)   # :56
                if self.image_space:
                    perturbation = torch.clamp(data_clean + perturbation, 0, 1) - data_clean
                assert not perturbation.isnan().any()
        out = data_clean + perturbation.detach()
        return (out, loss) if return_loss else out


def apgd(model, x, norm, eps, n_iter=10, use_rs=False, mask=None, alpha=None, n_iter_2=None,
         n_iter_min=None, pert_init=None, verbose=False, is_train=True):
    """vlm_eval/attacks/apgd.py:163-370 -> (x_best, loss_best, x_best_adv)"""
    assert x.shape[0] == 1                                                               # :166
    norm = norm.replace("l", "L")
    if norm not in ("Linf", "L2"):
        raise NotImplementedError(norm)
    ndims, bs = x.dim() - 1, x.shape[0]
    if not use_rs:
        x_adv = x.clone()
    elif norm == "Linf":
        x_adv = x + torch.zeros_like(x).uniform_(-eps, eps).detach()                     # :175-176
    else:
        t = torch.randn(x.shape).to(x.device).detach()
        x_adv = x + eps * torch.ones_like(x).detach() * t / (L2_norm(t, keepdim=True) + 1e-12)
    if pert_init is not None:
        assert not use_rs
        assert pert_init.shape == x.shape
        x_adv = x + pert_init                                                            # :183
    x_adv = x_adv.clamp(0.0, 1.0)
    x_best = x_adv.clone()
    x_best_adv = x_adv.clone()
    k0 = max(int((0.22 if n_iter_2 is None else n_iter_2) * n_iter), 1)                 # :194-197
    k_min = max(int((0.06 if n_iter_min is None else n_iter_min) * n_iter), 1)
    alpha = 2.0 if alpha is None else alpha
    step_size = alpha * eps * torch.ones([bs, *[1] * ndims], device=x.device)

    def fwd_bwd(xa, need_grad=True):
        xa = xa.detach().requires_grad_(True)
        with torch.enable_grad():
            loss_indiv = model(xa)                                                       # :220
            loss = loss_indiv.sum()
        g = None
        if need_grad:
            g = torch.autograd.grad(loss, [xa])[0].detach()
            if mask is not None:
                g *= mask                                                                # :224-225
        return loss_indiv.detach(), g

    loss_indiv, grad = fwd_bwd(x_adv)
    loss_indiv = loss_indiv.reshape(bs)
    grad_best = grad.clone()
    book = _ApgdBook(n_iter, bs, ndims, loss_indiv, x.device)
    book.k, book.k_min = k0, k_min
    x_adv_old = x_adv.clone()
    for i in range(n_iter):
        a = 0.75 if i > 0 else 1.0
        x_dtype = x_adv.dtype
        x_adv, x_adv_old = _apgd_step(norm, x, x_adv.detach(), x_adv_old, grad, step_size, eps, a,
                                      _l2_unit, scale_then_divide=True)
        x_adv = x_adv.to(dtype=x_dtype) + 0.0          # :289 (fp16 inputs: the iterate stays fp16)
        loss_indiv, g_new = fwd_bwd(x_adv, need_grad=i < n_iter - 1)                     # :293-297
        loss_indiv = loss_indiv.reshape(bs)
        if g_new is not None:
            grad = g_new
        x_best_adv = x_adv + 0.0                                                         # :303
        x_adv, grad, x_best, grad_best, step_size = book.update(
            i, loss_indiv, x_adv, grad, x_best, grad_best, step_size)
    return x_best, book.loss_best, x_best_adv


class APGD:
    """vlm_eval/attacks/apgd.py:5-48"""

    def __init__(self, model, norm, eps, mask_out="context", initial_stepsize=None,
                 decrease_every=None, decrease_every_max=None, random_init=False):
        self.model, self.norm, self.eps = model, norm, eps
        self.initial_stepsize, self.decrease_every = initial_stepsize, decrease_every
        self.decrease_every_max, self.random_init = decrease_every_max, random_init
        self.mask_out = mask_out if mask_out != "none" else None

    def perturb(self, data_clean, iterations, pert_init=None, verbose=False):
        mask = set_mask(data_clean, self.mask_out)
        data_adv, _, _ = apgd(self.model, data_clean, norm=self.norm, eps=self.eps,
                              n_iter=iterations, use_rs=self.random_init, mask=mask,
                              alpha=self.initial_stepsize, n_iter_2=self.decrease_every,
                              n_iter_min=self.decrease_every_max, pert_init=pert_init,
                              verbose=verbose)
        return data_adv

    def __str__(self):
        return "APGD"
