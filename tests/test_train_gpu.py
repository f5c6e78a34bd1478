"""Outer training step (a12): parameter gradients of the CUDA engine vs oracle autograd,
the fused AdamW vs torch.optim.AdamW, and a short FARE loop that must reduce the loss."""
import json
import os

import pytest
import torch

from oracle import attacks as oa
from oracle import vit as ovit

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _log(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "test_metrics.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **kw}) + "\n")
    print(name, kw)


@pytest.mark.parametrize("loss", ["l2", "ce"])
@pytest.mark.parametrize("name", ["tiny-32", "tiny-14"])
def test_param_grads_match_autograd(name, loss):
    from robustvlm_b200.engine import FMT_BF16, LossSpec, VitEngine
    B = 4
    o = ovit.build(name, seed=1)
    model = ovit.ClipVisionModel(o).train()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(B, 3, 224, 224, generator=g)
    D = o.cfg.out_dim
    e_orig = torch.randn(B, D, generator=g) * 0.5
    text = torch.nn.functional.normalize(torch.randn(D, 10, generator=g), dim=0)
    y = torch.randint(0, 10, (B,), generator=g)
    normalize = loss == "ce"
    for p in o.parameters():
        p.requires_grad_(True)
    lval = oa.compute_loss(loss, model(x, output_normalize=normalize), y, e_orig, 100.0, text, "mean")
    ref = dict(zip([n for n, _ in o.named_parameters()],
                   torch.autograd.grad(lval, list(o.parameters()))))
    eng = VitEngine(o.state_dict(), fmt_fwd=FMT_BF16, fmt_bwd=FMT_BF16)
    grads = {k: torch.full_like(v, float("nan"), device="cuda") for k, v in o.state_dict().items()}
    spec = LossSpec(loss, normalize, emb_orig=e_orig.cuda(), text=text.cuda(), targets=y.cuda())
    loss_b, _ = eng.train_grads(x.cuda(), spec, grads, 1.0 / B)
    torch.cuda.synchronize()
    assert abs(float(loss_b.mean()) - float(lval)) <= 0.05 * abs(float(lval))
    worst = ("", 1.0)
    for k, r in ref.items():
        gk = grads[k].cpu()
        assert torch.isfinite(gk).all(), k
        cos = torch.nn.functional.cosine_similarity(gk.flatten().double(), r.flatten().double(), dim=0)
        ratio = float(gk.norm() / r.norm().clamp_min(1e-30))
        if float(cos) < worst[1]:
            worst = (k, float(cos))
        assert float(cos) > 0.97, (k, float(cos))
        assert 0.9 < ratio < 1.1, (k, ratio)
    _log("param_grads", model=name, loss=loss, worst_key=worst[0], worst_cos=worst[1])


def test_adamw_matches_torch():
    from robustvlm_b200 import _native as nat
    torch.manual_seed(0)
    n = 10007
    p = torch.randn(n, device="cuda")
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-2)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(n, device="cuda")
        ref_p.grad = g.clone()
        opt.step()
        nat.call("rvlm_adamw_step", nat.ptr(p), nat.ptr(g), nat.ptr(m), nat.ptr(v), nat.c_int64(n),
                 nat.c_float(1e-3), nat.c_float(0.9), nat.c_float(0.95), nat.c_float(1e-8),
                 nat.c_float(1e-2), nat.c_int(step), nat.stream_ptr())
        torch.cuda.synchronize()
        assert float((p - ref_p.detach()).abs().max()) < 2e-6


def test_fare_training_loop_reduces_loss():
    from robustvlm_b200 import model as rm
    from robustvlm_b200.train import AdversarialTrainer
    name, B = "tiny-32", 8
    o = ovit.build(name, seed=1)
    vt, vto = rm.VisionTransformer(name), rm.VisionTransformer(name)
    vt.load_state_dict(o.state_dict())
    vto.load_state_dict(o.state_dict())
    m = rm.ClipVisionModel(vt.cuda(), None, None)
    mo = rm.ClipVisionModel(vto.cuda(), None, None).eval()
    tr = AdversarialTrainer(m, mo, lr=3e-4, wd=1e-4, steps=40, warmup=2, iterations=3)
    g = torch.Generator().manual_seed(0)
    data = torch.rand(B, 3, 224, 224, generator=g).cuda()
    y = torch.zeros(B, dtype=torch.long).cuda()
    losses = []
    for _ in range(12):
        losses.append(float(tr.train_step(data, y)["loss"]))
    _log("fare_loop", losses=losses)
    assert all(l == l for l in losses)
    assert losses[-1] < 0.7 * losses[0]
