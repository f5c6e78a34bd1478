"""Full-size parity: the CUDA path on ViT-L/14 (24 layers, W = 1024) against the CPU oracle on the
SAME weights and inputs, for BASELINE.json configs 2, 3 and 4 (VERDICT r01 "next" item 1).

  (a) embeddings (raw and normalised) within 1e-3 relative of the oracle          north_star
  (b) one-step gradient: cosine, relative error, sign-mismatch fraction -- overall and restricted
      to |g_ref| > 0.05 median|g_ref| (a different fp32 summation order alone flips the sign of
      near-zero gradients) -- for FARE with model_orig != model, FARE with
      model_orig == model (the cancellation site, train/adversarial_training_clip.py:509-521) and
      the TeCoA cross-entropy
  (c) 10-step PGD (train/pgd_train.py:29-68): identical-delta fraction, loss reached within 1 %
  (d) config 4: APGDAttack(n_iter=20, norm='L2', eps=2.0, n_restarts=2) on the same CPU-RNG
      start (autoattack/autopgd_base.py:205-548): loss reached, robust accuracy, eps-ball

Every engine precision mode is measured; the asserts are set at <= 2x the values measured on
B200 (profiles/r02_test_metrics.jsonl) so a regression fails.  The oracle runs on the host CPU
(one ViT-L/14 forward + backward of 4 images takes ~1-2 s)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import attacks as oa
from oracle import vit as ovit

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
B = 4
EPS_FARE, EPS_TECOA, STEP = 4 / 255, 2 / 255, 1 / 255

# engine precision modes (robustvlm_b200.engine.PRECISIONS): thresholds per mode =
#   emb rel err, [FARE-diff, FARE-same, CE] x (sign mismatch, mismatch among big), 10-step
#   identical fraction FARE / TeCoA
MODES = os.environ.get("RVLM_TEST_MODES", "fast,fast16,split").split(",")
# measured on B200 (profiles/r02_test_metrics.jsonl), asserts at <= 2x:
#   mode      emb rel   fare_diff            fare_same (cancellation)  ce                10-step identical
#   fast      4.0e-4    1.84e-3 / 0          1.57e-2 / 5.9e-3          1.92e-3 / 0       0.98939 / 0.99643
#   fast16    4.0e-4    2.8e-4  / 0          1.58e-2 / 5.9e-3          3.0e-4  / 0       0.99808 / 0.99923
#   split     1.27e-5   2.6e-4  / 0          5.6e-4  / 0               2.6e-4  / 0       0.99820 / 0.99930
LIMITS = {
    #          emb      fare_diff         fare_same         ce                same10 fare / ce   loss rel
    "fast":   (8.5e-4, (3.7e-3, 1e-4), (3.2e-2, 1.2e-2), (3.9e-3, 1e-4), 0.979, 0.9929, 1e-5),
    "fast16": (8.5e-4, (5.6e-4, 1e-4), (3.2e-2, 1.2e-2), (6.1e-4, 1e-4), 0.9961, 0.9984, 2e-6),
    "split":  (2.6e-5, (5.2e-4, 1e-4), (1.2e-3, 1e-4), (5.2e-4, 1e-4), 0.9964, 0.9986, 2e-6),
}
# config 4 (APGD-CE ell_2, 20 iterations x 2 restarts): logits error / |logits|, ||adv - adv_ref|| / ||adv_ref - x||
LIMITS4 = {"fast": (9e-4, 2.2e-2), "fast16": (9e-4, 2.2e-2), "split": (3e-5, 1.4e-3)}


def _log(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "test_metrics.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **kw}) + "\n")
    print(name, kw)


@pytest.fixture(scope="module")
def ref():
    """Everything the oracle contributes, computed once on the host CPU."""
    torch.set_num_threads(max(1, min(64, os.cpu_count() or 1)))
    o1, o2 = ovit.build("ViT-L-14", seed=1), ovit.build("ViT-L-14", seed=2)
    m1, m2 = ovit.ClipVisionModel(o1).eval(), ovit.ClipVisionModel(o2).eval()
    for p in list(m1.parameters()) + list(m2.parameters()):
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, 3, 224, 224, generator=g)
    d0 = (torch.rand(B, 3, 224, 224, generator=g) * 2 - 1) * EPS_FARE
    d0t = d0 * (EPS_TECOA / EPS_FARE)
    text = F.normalize(torch.randn(768, 1000, generator=g), dim=0)
    y = torch.randint(0, 1000, (B,), generator=g)
    r = dict(o1=o1, m1=m1, x=x, d0=d0, d0t=d0t, text=text, y=y)
    with torch.no_grad():
        r["e_adv"] = m1(x + d0, output_normalize=False)
        r["e_orig_diff"] = m2(x, output_normalize=False)
        r["e_orig_same"] = m1(x, output_normalize=False)

    def grad(wrap, normalize, d, tg):
        xa = (x + d).requires_grad_(True)
        return torch.autograd.grad(wrap(m1(xa, output_normalize=normalize), tg), xa)[0]

    r["w_diff"] = oa.ComputeLossWrapper(r["e_orig_diff"], None, "mean", "l2", 100.0)
    r["w_same"] = oa.ComputeLossWrapper(r["e_orig_same"], None, "mean", "l2", 100.0)
    r["w_ce"] = oa.ComputeLossWrapper(None, text, "mean", "ce", 100.0)
    r["g_diff"] = grad(r["w_diff"], False, d0, None)
    r["g_same"] = grad(r["w_same"], False, d0, None)
    r["g_ce"] = grad(r["w_ce"], True, d0t, y)
    r["pgd_fare"] = oa.pgd(m1, r["w_diff"], x, None, "linf", EPS_FARE, 10, STEP, False,
                           perturbation=d0.clone(), mode="max")
    r["pgd_ce"] = oa.pgd(m1, r["w_ce"], x, y, "linf", EPS_TECOA, 10, STEP, True,
                         perturbation=d0t.clone(), mode="max")
    return r


_MODELS = {}


def _ours(ref, mode):
    """our ClipVisionModel with the oracle's weights, one per precision mode (kept: building the
    16-bit weight copies of a ViT-L/14 takes a moment)"""
    from robustvlm_b200 import model as rm
    if mode not in _MODELS:
        _MODELS.clear()                                     # one resident engine at a time
        torch.cuda.empty_cache()
        vt = rm.VisionTransformer("ViT-L-14", precision=mode)
        vt.load_state_dict(ref["o1"].state_dict())
        _MODELS[mode] = rm.ClipVisionModel(vt.cuda(), None, None).eval()
    return _MODELS[mode]


def _sign_stats(g, g_ref):
    g, g_ref = g.cpu(), g_ref.cpu()
    mis = torch.sign(g) != torch.sign(g_ref)
    big = g_ref.abs() > 0.05 * g_ref.abs().median()
    # per-sample positive rescaling is allowed on the attack path: compare directions per sample
    cos = F.cosine_similarity(g.flatten(1).double(), g_ref.flatten(1).double(), dim=1)
    gs = g.flatten(1).double()
    gr = g_ref.flatten(1).double()
    scale = (gs * gr).sum(1, keepdim=True) / (gs * gs).sum(1, keepdim=True).clamp_min(1e-300)
    rel = ((gs * scale - gr).norm(dim=1) / gr.norm(dim=1)).max()
    return dict(mismatch=float(mis.float().mean()),
                mismatch_big=float((mis & big).float().sum() / big.float().sum()),
                cos_min=float(cos.min()), rel_err=float(rel))


@pytest.mark.parametrize("mode", MODES)
def test_embedding_vitl14(ref, mode):
    m = _ours(ref, mode)
    x, d0 = ref["x"].cuda(), ref["d0"].cuda()
    with torch.no_grad():
        e = m(x + d0, output_normalize=False).cpu()
        en = m(x + d0, output_normalize=True).cpu()
    rel = float((e - ref["e_adv"]).norm() / ref["e_adv"].norm())
    reln = float((en - F.normalize(ref["e_adv"], dim=-1)).norm(dim=1).max())
    _log("fullsize_embedding", mode=mode, rel_err=rel, rel_err_normalized=reln)
    assert rel <= LIMITS[mode][0] and reln <= LIMITS[mode][0]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", ["fare_diff", "fare_same", "ce"])
def test_one_step_gradient_vitl14(ref, mode, case):
    from robustvlm_b200.engine import LossSpec
    m = _ours(ref, mode)
    eng = m.engine()
    x = ref["x"].cuda()
    if case == "ce":
        spec = LossSpec("ce", True, text=ref["text"].cuda(), logit_scale=100.0, targets=ref["y"].cuda())
        d, g_ref = ref["d0t"].cuda(), ref["g_ce"]
    else:
        eo = ref["e_orig_diff"] if case == "fare_diff" else ref["e_orig_same"]
        spec = LossSpec("l2", False, emb_orig=eo.cuda())
        d, g_ref = ref["d0"].cuda(), ref["g_diff"] if case == "fare_diff" else ref["g_same"]
    out = eng.loss_grad(x, spec, delta=d)
    st = _sign_stats(out["grad"], g_ref)
    _log("fullsize_gradient", mode=mode, case=case, **st)
    lim = LIMITS[mode][1 + ["fare_diff", "fare_same", "ce"].index(case)]
    assert st["mismatch"] <= lim[0] and st["mismatch_big"] <= lim[1]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", ["fare", "tecoa"])
def test_ten_step_pgd_vitl14(ref, mode, case):
    """configs 2 and 3 at B = 4 against train/pgd_train.py::pgd run by the oracle"""
    from robustvlm_b200 import attacks, model as rm
    m, m1 = _ours(ref, mode), ref["m1"]
    x = ref["x"]
    if case == "fare":
        wrap = rm.ComputeLossWrapper(ref["e_orig_diff"].cuda(), None, "mean", "l2", 100.0)
        out = attacks.pgd(m, wrap, x.cuda(), None, "linf", EPS_FARE, 10, STEP, False,
                          perturbation=ref["d0"].cuda(), mode="max").cpu()
        r, w, tg, nrm, eps = ref["pgd_fare"], ref["w_diff"], None, False, EPS_FARE
    else:
        wrap = rm.ComputeLossWrapper(None, ref["text"].cuda(), "mean", "ce", 100.0)
        out = attacks.pgd(m, wrap, x.cuda(), ref["y"].cuda(), "linf", EPS_TECOA, 10, STEP, True,
                          perturbation=ref["d0t"].cuda(), mode="max").cpu()
        r, w, tg, nrm, eps = ref["pgd_ce"], ref["w_ce"], ref["y"], True, EPS_TECOA
    with torch.no_grad():
        l_ours = float(w(m1(out, output_normalize=nrm), tg))
        l_ref = float(w(m1(r, output_normalize=nrm), tg))
    same = float(((out - x) == (r - x)).float().mean())
    _log("fullsize_pgd10", mode=mode, case=case, identical_frac=same, loss_ours=l_ours,
         loss_ref=l_ref, max_abs_err=float((out - r).abs().max()),
         mean_abs_err=float((out - r).abs().mean()))
    assert float((out - x).abs().max()) <= eps + 1e-6 and out.min() >= 0 and out.max() <= 1
    assert abs(l_ours - l_ref) <= LIMITS[mode][6] * abs(l_ref)     # north_star: within 1 %
    assert same >= LIMITS[mode][4 if case == "fare" else 5]


@pytest.mark.parametrize("mode", MODES)
def test_config4_apgd_l2_vitl14(ref, mode):
    """APGD-CE ell_2 eps = 2.0 with restarts on the zero-shot classifier, against
    autoattack/autopgd_base.py:205-548 run by the oracle from the same CPU-RNG start."""
    from robustvlm_b200 import attacks, model as rm
    m = _ours(ref, mode)
    x, text = ref["x"], ref["text"]
    cm_ref = ovit.ClassificationModel(ref["o1"], text, 100.0)
    cm = rm.ClassificationModel(m, text.cuda())
    with torch.no_grad():
        lg_ref = cm_ref(x)
        lg = cm(x.cuda()).cpu()
    y = lg_ref.argmax(1)
    R = oa.APGDAttack(cm_ref, n_iter=20, norm="L2", n_restarts=2, eps=2.0, seed=0, alpha=2.0)
    A = attacks.APGDAttack(cm, n_iter=20, norm="L2", n_restarts=2, eps=2.0, seed=0, alpha=2.0)
    r = R.perturb(x, y)
    out = A.perturb(x.cuda(), y.cuda()).cpu()
    ce = torch.nn.CrossEntropyLoss(reduction="none")
    with torch.no_grad():
        l_ref, l_ours = ce(cm_ref(r), y), ce(cm_ref(out), y)
        racc_ref = float((cm_ref(r).argmax(1) == y).float().mean())
        racc_ours = float((cm_ref(out).argmax(1) == y).float().mean())
    l2o = (out - x).flatten(1).norm(dim=1)
    _log("fullsize_config4", mode=mode, logits_max_abs_err=float((lg - lg_ref).abs().max()),
         logits_scale=float(lg_ref.abs().max()), racc_ref=racc_ref, racc_ours=racc_ours,
         ce_ref=l_ref.tolist(), ce_ours=l_ours.tolist(), l2_ours=l2o.tolist(),
         l2_ref=(r - x).flatten(1).norm(dim=1).tolist(),
         adv_rel_diff=float((out - r).norm() / (r - x).norm().clamp_min(1e-12)))
    assert float((lg - lg_ref).abs().max()) <= LIMITS4[mode][0] * float(lg_ref.abs().max())
    assert float(l2o.max()) <= 2.0 * (1 + 1e-4) and out.min() >= 0 and out.max() <= 1
    assert racc_ours == racc_ref
    assert float((out - r).norm() / (r - x).norm().clamp_min(1e-12)) <= LIMITS4[mode][1]
    assert float((l_ours - l_ref).abs().max()) <= 0.02 * float(l_ref.abs().max())


@pytest.mark.parametrize("fmt_fwd", [0, 1], ids=["mixed_f16_fwd", "bf16_fwd"])
def test_outer_step_param_grads_vitl14(ref, fmt_fwd):
    """The outer FARE step at full size (row a12, train/adversarial_training_clip.py:340-362):
    every parameter gradient of `rvlm_vit_train_grads` against autograd through the fp32 oracle on
    the same adversarial batch -- with the trainer's engine (fp16 forward operands, bf16 gradient
    operands: embedding within 1e-3) and with the all-bf16 engine of round 1 (3.3e-3)."""
    from robustvlm_b200.engine import FMT_BF16, LossSpec, VitEngine
    o1, m1 = ref["o1"], ref["m1"]
    x_adv = ref["pgd_fare"]
    params = dict(o1.named_parameters())
    for p in params.values():
        p.requires_grad_(True)
    try:
        loss = oa.compute_loss("l2", m1(x_adv, output_normalize=False), None, ref["e_orig_diff"],
                               100.0, None, "mean")
        grads_ref = dict(zip(params, torch.autograd.grad(loss, list(params.values()))))
    finally:
        for p in params.values():
            p.requires_grad_(False)
    _MODELS.clear()
    torch.cuda.empty_cache()
    eng = VitEngine(o1.state_dict(), fmt_fwd=fmt_fwd, fmt_bwd=FMT_BF16)
    grads = {k: torch.full_like(v, float("nan"), device="cuda") for k, v in o1.state_dict().items()}
    spec = LossSpec("l2", False, emb_orig=ref["e_orig_diff"].cuda())
    loss_b, emb = eng.train_grads(x_adv.cuda(), spec, grads, 1.0 / B)
    torch.cuda.synchronize()
    loss_rel = abs(float(loss_b.mean()) - float(loss)) / abs(float(loss))
    worst, worst_ratio = ("", 1.0), 0.0
    num = den_a = den_b = 0.0
    for k, r in grads_ref.items():
        g = grads[k].cpu().double().flatten()
        rr = r.double().flatten()
        assert torch.isfinite(g).all(), k
        cos = float(torch.nn.functional.cosine_similarity(g, rr, dim=0))
        if cos < worst[1]:
            worst = (k, cos)
        worst_ratio = max(worst_ratio, abs(float(g.norm() / rr.norm().clamp_min(1e-30)) - 1.0))
        num += float(g @ rr)
        den_a += float(g @ g)
        den_b += float(rr @ rr)
    cos_all = num / (den_a * den_b) ** 0.5
    with torch.no_grad():
        e_ref = m1(x_adv, output_normalize=False)
    emb_rel = float((emb.cpu() - e_ref).norm() / e_ref.norm())
    _log("fullsize_outer_step", fmt_fwd=fmt_fwd, loss_rel_err=loss_rel, cos_all=cos_all,
         worst_key=worst[0], worst_cos=worst[1], worst_norm_ratio_err=worst_ratio, emb_rel_err=emb_rel)
    # measured (bf16 forward): loss 1.5e-4, cos 0.99999, worst tensor 0.99998, embedding 3.3e-3
    assert loss_rel <= 3e-4 and cos_all >= 0.99998 and worst[1] >= 0.99996
    assert emb_rel <= (8.5e-4 if fmt_fwd == 0 else 6.6e-3)
