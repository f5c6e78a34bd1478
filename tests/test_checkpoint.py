"""Checkpoint formats and the lr schedule -- host logic, CPU only.

The trainer's flat AdamW / SGD buffers must convert to and from the dictionaries
`torch.optim.{AdamW,SGD}.state_dict()` produces over `model.parameters()` (what the reference
saves next to every model checkpoint, train/adversarial_training_clip.py:238-239,468-479), and
`cosine_lr` must equal the copy of open_clip's scheduler vendored in the reference
(CLIP_benchmark/clip_benchmark/metrics/linear_probe.py:13-32)."""
import os
import sys

import pytest
import torch

from oracle import ref_import
from oracle import vit as ovit
from robustvlm_b200 import checkpoint as ck
from robustvlm_b200 import model as rm
from robustvlm_b200.train import cosine_lr


def _stepped(opt_cls, **kw):
    o = ovit.build("tiny-32", seed=1)
    params = list(o.parameters())
    opt = opt_cls(params, **kw)
    g = torch.Generator().manual_seed(0)
    for _ in range(3):
        for p in params:
            p.grad = torch.randn(p.shape, generator=g) * 0.01
        opt.step()
    return o, opt


def test_parameter_order_is_open_clip_layout():
    """Own parameters first (class_embedding, positional_embedding, proj), then the children in
    registration order -- the order the reference hands to the optimizer (:194)."""
    ours = [n for n, *_ in ck.param_layout(rm.VisionTransformer("tiny-32"))]
    ref = [n for n, _ in ovit.build("tiny-32", seed=0).named_parameters()]
    assert ours == ref
    assert ours[:4] == ["class_embedding", "positional_embedding", "proj", "conv1.weight"]
    assert list(rm.VisionTransformer("tiny-32").state_dict()) == list(
        ovit.build("tiny-32", seed=0).state_dict())


def test_adamw_state_dict_round_trip():
    o, opt = _stepped(torch.optim.AdamW, lr=1e-3, weight_decay=1e-4)
    sd = opt.state_dict()
    layout = ck.param_layout(o)
    n = sum(sz for *_, sz in layout)
    m, v = torch.zeros(n), torch.zeros(n)
    step, group = ck.load_optimizer_state_dict(sd, layout, m, v)
    assert step == 3 and group["lr"] == 1e-3 and tuple(group["betas"]) == (0.9, 0.999)
    out = ck.adamw_state_dict(layout, m, v, step, lr=group["lr"], betas=group["betas"],
                              eps=group["eps"], weight_decay=group["weight_decay"])
    assert set(out["state"]) == set(sd["state"])
    for i in sd["state"]:
        assert torch.equal(out["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])
        assert torch.equal(out["state"][i]["exp_avg_sq"], sd["state"][i]["exp_avg_sq"])
        assert float(out["state"][i]["step"]) == float(sd["state"][i]["step"])
    # ... and torch accepts it back and continues identically
    o2 = ovit.build("tiny-32", seed=1)
    o2.load_state_dict(o.state_dict())
    opt2 = torch.optim.AdamW(list(o2.parameters()), lr=5.0)       # lr comes from the state
    opt2.load_state_dict(out)
    g = torch.Generator().manual_seed(9)
    for p, p2 in zip(o.parameters(), o2.parameters()):
        p.grad = torch.randn(p.shape, generator=g)
        p2.grad = p.grad.clone()
    opt.step()
    opt2.step()
    for p, p2 in zip(o.parameters(), o2.parameters()):
        assert torch.equal(p, p2)


def test_sgd_state_dict_round_trip_and_fresh_state():
    o, opt = _stepped(torch.optim.SGD, lr=1e-2, momentum=0.9, weight_decay=1e-4)
    layout = ck.param_layout(o)
    n = sum(sz for *_, sz in layout)
    buf = torch.zeros(n)
    step, group = ck.load_optimizer_state_dict(opt.state_dict(), layout, buf)
    assert step >= 1 and group["momentum"] == 0.9
    out = ck.sgd_state_dict(layout, buf, step, lr=1e-2, momentum=0.9, weight_decay=1e-4)
    for i, st in opt.state_dict()["state"].items():
        assert torch.equal(out["state"][i]["momentum_buffer"], st["momentum_buffer"])
    # before the first step torch holds no state
    assert ck.adamw_state_dict(layout, buf, buf, 0, lr=1e-3)["state"] == {}
    with pytest.raises(ValueError):
        ck.load_optimizer_state_dict(opt.state_dict(), layout[:-1], buf)


def test_saving_rules_and_loading(tmp_path):
    sd = ovit.build("tiny-32", seed=2).state_dict()
    opt_fn = lambda: {"state": {}, "param_groups": [{"lr": 1.0, "params": []}]}   # noqa: E731
    out = str(tmp_path)
    assert ck.save_periodic(sd, opt_fn, out, 7, 100) == []
    assert ck.save_periodic(sd, opt_fn, out, 10, 100) == ["step_10"]            # steps // 10
    assert ck.save_periodic(sd, opt_fn, out, 10, 100, save_checkpoints=False) == []
    assert ck.save_periodic(sd, opt_fn, out, 200, 2000) == ["step_200", "fallback_200"]
    assert ck.save_periodic(sd, opt_fn, out, 400, 2000) == ["step_400", "fallback_400"]
    files = sorted(os.listdir(os.path.join(out, "checkpoints")))
    assert "fallback_400.pt" in files and "fallback_400_opt.pt" in files
    assert not any(f.startswith("fallback_200") for f in files)                 # :476-478
    assert ck.step_of("fallback_400_opt.pt") == 400 and ck.step_of("final.pt") is None
    back = ck.load_visual_state_dict(os.path.join(out, "checkpoints", "step_10.pt"))
    assert list(back) == list(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    full = {"visual." + k: v for k, v in sd.items()}
    full["logit_scale"] = torch.tensor(4.6)
    assert list(ck.load_visual_state_dict(full)) == list(sd)                    # a whole CLIP
    assert list(ck.load_visual_state_dict({"vision_encoder_state_dict": dict(sd)})) == list(sd)
    vt = rm.VisionTransformer("tiny-32")
    vt.load_state_dict(back)                                                    # strict
    with pytest.raises(KeyError):
        ck.load_visual_state_dict({"foo": torch.zeros(1)})


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted")
def test_cosine_lr_equals_reference_copy():
    ref_import.load()                       # puts the reference root (autoattack, ...) on sys.path
    sys.path.insert(0, os.path.join(ref_import.REFERENCE_ROOT, "CLIP_benchmark"))
    import importlib
    ref = importlib.import_module("clip_benchmark.metrics.linear_probe").cosine_lr
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1e-5)
    sched = ref(opt, 1e-5, 1400, 20000)
    ours = cosine_lr(1e-5, 1400, 20000)
    for step in (0, 1, 699, 1399, 1400, 1401, 10000, 19999, 20000):
        sched(step)
        assert opt.param_groups[0]["lr"] == pytest.approx(ours(step), rel=1e-12, abs=0)


def test_cosine_lr_known_values():
    lr = cosine_lr(1e-5, 1400, 20000)
    assert lr(0) == pytest.approx(1e-5 / 1400) and lr(1399) == pytest.approx(1e-5)
    assert lr(1400) == pytest.approx(1e-5) and lr(20000) == pytest.approx(0.0, abs=1e-20)
    assert lr(10700) == pytest.approx(0.5e-5)


def test_evaluation_state_round_trip(tmp_path):
    """autoattack/state.py: the evaluation state survives a save / load, in the reference's keys,
    and a file written the reference's way (the set as its repr) is read too"""
    import json
    from robustvlm_b200.autoattack import EvaluationState
    p = tmp_path / "state.json"
    st = EvaluationState({"apgd-ce", "apgd-t"}, path=p)
    st.to_disk()
    st.robust_flags = torch.tensor([True, False, True, True])
    st.clean_accuracy = 0.75
    st.add_run_attack("apgd-ce")
    st.to_disk(force=True)
    d = json.load(open(p))
    assert set(d) == {"_attacks_to_run", "path", "_run_attacks", "_robust_flags", "_last_saved",
                      "_SAVE_TIMEOUT", "_clean_accuracy"}
    back = EvaluationState.from_disk(p)
    assert back.attacks_to_run == {"apgd-ce", "apgd-t"} and back.run_attacks == {"apgd-ce"}
    assert back.robust_flags.tolist() == [True, False, True, True] and back.clean_accuracy == 0.75
    d["_attacks_to_run"] = str({"apgd-ce", "apgd-t"})          # what json.dump(default=str) writes
    json.dump(d, open(p, "w"))
    assert EvaluationState.from_disk(p).attacks_to_run == {"apgd-ce", "apgd-t"}
