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
