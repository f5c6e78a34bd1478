"""VLM-evaluation attacks of the reference on the CUDA update kernels.

  PGD          vlm_eval/attacks/pgd.py:5-87    (same ctor / perturb / _set_mask)
  APGD, apgd   vlm_eval/attacks/apgd.py:5-48,163-370  (same ctor / perturb; apgd returns
               (x_best, loss_best, x_best_adv))

These attacks wrap an arbitrary VLM (`eval_model(x) -> loss`, LLaVA / OpenFlamingo in
vlm_eval/run_evaluation.py:1010-1023), so the gradient comes from torch.autograd; everything
between two model evaluations -- mask, NaN scrub, sign / ell_2 normalisation, momentum, step,
projection, image clamp, APGD proposal, best-point / step-size bookkeeping and restarts -- runs
in the C-ABI kernels (`rvlm_pgd_update_*_masked`, `rvlm_mul_f32`, `rvlm_apgd_step_*`,
`rvlm_apgd_bookkeep`) with no host read of a device value inside the loop.  The arithmetic is
fp32.  For fp16 / bf16 inputs the reference keeps the iterate in the input dtype
(`x_adv = x_adv_1.to(dtype=x_adv.dtype)`, apgd.py:289) with fp32 arithmetic inside a step (the
fp32 step size promotes every expression): here the step kernels run in fp32 on the widened
values and the iterate is rounded to the input dtype after every step, which reproduces that
sequence except for `grad2 = x_adv - x_adv_old` (:205), which the reference rounds to the input
dtype (at most one ulp of it).  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _native as nat
from ._native import c_float, c_int, c_int64
from .attacks import _apgd_core, _flat_n, _need_cuda, _norm_id, _upd_ws
from .engine import NORM_LINF


def _set_mask(data, mask_out):
    """vlm_eval/attacks/pgd.py:75-87 == apgd.py:34-46"""
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
    """Minimize or maximize the loss `forward(x)` returns (vlm_eval/attacks/pgd.py:5-73)."""

    def __init__(self, forward, norm, eps, mode="min", mask_out="context", image_space=True):
        self.model = forward
        self.norm, self.eps, self.momentum = norm, eps, 0.9
        self.mode, self.mask_out, self.image_space = mode, mask_out, image_space

    def _set_mask(self, data):
        return _set_mask(data, self.mask_out)

    def perturb(self, data_clean, iterations, stepsize, perturbation=None, verbose=False,
                return_loss=False):
        _need_cuda(data_clean, "data_clean")
        if self.image_space:
            assert torch.max(data_clean) < 1.0 + 1e-6 and torch.min(data_clean) > -1e-6   # :24
        if self.mode not in ("min", "max"):
            raise ValueError(f"Unknown mode: {self.mode}")                               # :54
        nid = _norm_id(self.norm)
        dt = data_clean.dtype
        x = data_clean.detach().float().contiguous()
        B, n = x.shape[0], _flat_n(x)
        delta = (torch.zeros_like(x) if perturbation is None
                 else perturbation.detach().float().clone().contiguous())
        mask = self._set_mask(x)
        velocity = torch.zeros_like(x)
        ws = _upd_ws(B, x.device) if nid != NORM_LINF else None
        loss = None
        for i in range(iterations):
            delta.requires_grad_(True)
            with torch.enable_grad():
                loss = self.model((x + delta).to(dt))                                    # :33
                if verbose and (i % (iterations // 10 + 1) == 0 or i == iterations - 1):
                    print(f"[iteration] {i} [loss] {loss.item()}")
            grad = torch.autograd.grad(loss, delta)[0].detach().float().contiguous()
            delta = delta.detach()
            args = (nat.ptr(delta), nat.ptr(velocity), nat.ptr(grad), nat.ptr(x), nat.ptr(mask),
                    c_int(int(self.image_space)), c_int(B), c_int64(n), c_float(stepsize),
                    c_float(self.eps), c_int(int(self.mode == "max")), c_float(self.momentum))
            if nid == NORM_LINF:
                nat.call("rvlm_pgd_update_linf_masked", *args, nat.stream_ptr())
            else:
                nat.call("rvlm_pgd_update_l2_masked", *args, nat.ptr(ws), nat.stream_ptr())
        out = torch.empty_like(x)
        nat.call("rvlm_add_f32", nat.ptr(out), nat.ptr(x), nat.ptr(delta), c_int64(x.numel()),
                 nat.stream_ptr())
        out = out.to(dt)
        return (out, loss) if return_loss else out


def apgd(model, x, norm, eps, n_iter=10, use_rs=False, mask=None, alpha=None, n_iter_2=None,
         n_iter_min=None, pert_init=None, verbose=False, is_train=True):
    """vlm_eval/attacks/apgd.py:163-370 -> (x_best, loss_best, x_best_adv).  `model(x)` returns
    the per-sample loss [1] to MAXIMISE; batch size 1 as in the reference (:166)."""
    assert x.shape[0] == 1                                                               # :166
    norm = norm.replace("l", "L")
    if norm not in ("Linf", "L2"):
        raise NotImplementedError(f"norm {norm}: only Linf / L2 are built (SURVEY.md §8a)")
    _need_cuda(x, "x")
    dt = x.dtype
    xf = x.detach().float().contiguous()
    B, eps = xf.shape[0], float(eps)
    if not use_rs:
        x_adv = xf.clone()
    elif norm == "Linf":
        x_adv = xf + torch.zeros_like(xf).uniform_(-eps, eps)                            # :175-176
    else:
        t = torch.randn(xf.shape).to(xf.device)                                          # CPU RNG
        tn = (t ** 2).view(B, -1).sum(-1).sqrt().view(-1, *([1] * (xf.dim() - 1)))
        x_adv = xf + eps * torch.ones_like(xf) * t / (tn + 1e-12)                        # :177-179
    if pert_init is not None:
        assert not use_rs
        assert pert_init.shape == x.shape, f"pert_init.shape: {pert_init.shape}, x.shape: {x.shape}"
        x_adv = xf + pert_init.detach().float()                                          # :183
    k0 = max(int((0.22 if n_iter_2 is None else n_iter_2) * n_iter), 1)                 # :194-197
    k_min = max(int((0.06 if n_iter_min is None else n_iter_min) * n_iter), 1)
    k_decr = max(int(0.03 * n_iter), 1)
    alpha = 2.0 if alpha is None else alpha
    maskf = mask.detach().float().contiguous() if mask is not None else None
    zero_pred = torch.zeros(B, dtype=torch.int32, device=xf.device)   # x_best_adv = x_adv (:303)

    def fwd_bwd(xa, need_grad):
        xin = xa.detach().clone().requires_grad_(need_grad)
        with torch.enable_grad():
            loss_indiv = model(xin.to(dt))                                               # :220,289
            g = None
            if need_grad:
                g = torch.autograd.grad(loss_indiv.sum(), [xin])[0].detach().float().contiguous()
                if maskf is not None:                                                    # :224-225
                    nat.call("rvlm_mul_f32", nat.ptr(g), nat.ptr(g), nat.ptr(maskf),
                             c_int64(g.numel()), nat.stream_ptr())
        return loss_indiv.detach().float().reshape(B), zero_pred, g

    x_best, _, loss_best, x_best_adv = _apgd_core(
        fwd_bwd, xf, x_adv, norm, eps, n_iter, alpha, 0.75, skip_last_grad=True,
        schedule=(k0, k_min, k_decr), round_to=None if dt == torch.float32 else dt)
    return x_best.to(dt), loss_best, x_best_adv.to(dt)


class APGD:
    """vlm_eval/attacks/apgd.py:5-48"""

    def __init__(self, model, norm, eps, mask_out="context", initial_stepsize=None,
                 decrease_every=None, decrease_every_max=None, random_init=False):
        self.model, self.norm, self.eps = model, norm, eps
        self.initial_stepsize = initial_stepsize
        self.decrease_every, self.decrease_every_max = decrease_every, decrease_every_max
        self.random_init = random_init
        self.mask_out = mask_out if mask_out != "none" else None

    def _set_mask(self, data):
        return _set_mask(data, self.mask_out)

    def perturb(self, data_clean, iterations, pert_init=None, verbose=False):
        mask = self._set_mask(data_clean)
        data_adv, _, _ = apgd(self.model, data_clean, norm=self.norm, eps=self.eps,
                              n_iter=iterations, use_rs=self.random_init, mask=mask,
                              alpha=self.initial_stepsize, n_iter_2=self.decrease_every,
                              n_iter_min=self.decrease_every_max, pert_init=pert_init,
                              verbose=verbose)
        return data_adv

    def __str__(self):
        return "APGD"
