"""Shared helpers for the parity tests (oracle = CPU checker, see oracle/__init__.py)."""
import torch

from oracle import vit as ovit


def make_oracle(name, seed):
    return ovit.ClipVisionModel(ovit.build(name, seed=seed)).eval()


def make_engine(name, seed, **kw):
    from robustvlm_b200.engine import VitEngine
    m = ovit.build(name, seed=seed)
    return VitEngine(m.state_dict(), **kw), ovit.ClipVisionModel(m).eval()


def synth_batch(B, seed=0, eps=4 / 255, image=224):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, image, image, generator=g)
    d0 = (torch.rand(B, 3, image, image, generator=g) * 2 - 1) * eps
    return x, d0


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def sign_agreement(g, g_ref, tau=0.0):
    """fraction of elements with equal sign, and the mismatch fraction among elements
    whose reference magnitude exceeds tau * median|g_ref|."""
    g, g_ref = g.cpu(), g_ref.cpu()
    mism = torch.sign(g) != torch.sign(g_ref)
    med = g_ref.abs().median()
    big = g_ref.abs() > tau * med
    return 1.0 - float(mism.float().mean()), float((mism & big).float().sum() / big.float().sum().clamp_min(1))


class ToyVlmLoss:
    """Stand-in for `eval_model(x) -> loss[1]` of vlm_eval/run_evaluation.py:1010-1023 with a
    gradient that is bit-identical on CPU and CUDA: every backward op is one elementwise IEEE
    operation (d loss / d x = w + (-2c d) with d = x - t), and the loss value is reduced in
    fp64 so its fp32 rounding does not depend on the summation order.  Concave in x, so APGD
    overshoots, oscillates and halves its step size (all branches of the bookkeeping run)."""

    def __init__(self, center, seed, spread=6 / 255, c=40.0):
        g = torch.Generator().manual_seed(seed)
        c0 = center.detach().float().cpu()
        self.w = (torch.randn(c0.shape, generator=g) * 0.5).to(center.device)
        self.t = (c0 + spread * torch.randn(c0.shape, generator=g)).to(center.device)
        self.c = c

    def to(self, device):
        o = ToyVlmLoss.__new__(ToyVlmLoss)
        o.w, o.t, o.c = self.w.to(device), self.t.to(device), self.c
        return o

    def __call__(self, x):
        d = x.float() - self.t
        v = self.w * x.float() - self.c * (d * d)
        return v.double().flatten(1).sum(1).float()

    def scalar(self, x):                       # PGD's model returns the scalar loss
        return self(x).sum()
