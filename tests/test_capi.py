"""The C-ABI library loads on a machine without a GPU and exports every symbol
include/rvlm.h declares; the Python shims mirror the reference's signatures and refuse
CPU tensors (no fallback)."""
import inspect
import os

import pytest
import torch

from robustvlm_b200 import _native as nat

LIB_BUILT = os.path.exists(nat.LIB_PATH)


@pytest.mark.skipif(not LIB_BUILT, reason="librvlm_b200.so not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    lib = nat.lib()
    syms = nat.declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.rvlm_abi_version.restype = nat.c_int
    assert lib.rvlm_abi_version() == 2            # RVLM_ABI_VERSION (split-operand forward)
    assert lib.rvlm_strerror(0) == b"ok"
    assert b"aligned" in lib.rvlm_strerror(-23)


def test_python_signatures_mirror_the_reference():
    from robustvlm_b200 import attacks, model
    assert list(inspect.signature(attacks.pgd).parameters)[:13] == [
        "forward", "loss_fn", "data_clean", "targets", "norm", "eps", "iterations", "stepsize",
        "output_normalize", "perturbation", "mode", "momentum", "verbose"]   # train/pgd_train.py:5-19
    assert list(inspect.signature(attacks.apgd_train).parameters) == [
        "model", "x", "y", "norm", "eps", "n_iter", "use_rs", "loss_fn", "verbose", "is_train",
        "initial_stepsize"]                                                  # train/apgd_train.py:125-126
    assert list(inspect.signature(attacks.APGDAttack.__init__).parameters)[1:] == [
        "predict", "n_iter", "norm", "n_restarts", "eps", "seed", "loss", "eot_iter", "rho",
        "topk", "verbose", "device", "use_largereps", "is_tf_model", "logger", "alpha", "use_rs"]
    assert list(inspect.signature(model.ComputeLossWrapper.__init__).parameters)[1:] == [
        "embedding_orig", "embedding_text_labels_norm", "reduction", "loss", "logit_scale"]
    assert list(inspect.signature(model.ClipVisionModel.__init__).parameters)[1:] == [
        "model", "args", "normalize"]


def test_state_dict_layout_is_open_clip():
    from oracle import vit
    from robustvlm_b200 import model
    ours = model.VisionTransformer("tiny-32").state_dict()
    ref = vit.build("tiny-32").state_dict()
    assert set(ours) == set(ref)
    assert all(ours[k].shape == ref[k].shape for k in ref)
    assert hasattr(model.VisionTransformer("tiny-32"), "proj")           # clip_encoder.py:13-17
    assert hasattr(model.VisionTransformer("tiny-32").transformer, "resblocks")


def test_cpu_tensors_are_refused():
    from robustvlm_b200 import attacks, model
    x = torch.rand(2, 3, 224, 224)
    with pytest.raises(nat.RvlmError):
        attacks.pgd(lambda v, output_normalize: v.flatten(1), lambda o, t: o.sum(), x, None,
                    "linf", 4 / 255, 1, 1 / 255, False)
    with pytest.raises(NotImplementedError):
        attacks.pgd(None, None, x.clamp(0, 1), None, "l7", 1, 1, 1, False) if False else \
            attacks._norm_id("l7")
    with pytest.raises(ValueError):
        model.compute_loss("hinge", x, None, None, 1.0)


def test_losses_match_oracle_on_cpu():
    from oracle import attacks as oa
    from robustvlm_b200 import model
    g = torch.Generator().manual_seed(0)
    e, t = torch.randn(5, 16, generator=g), torch.randn(5, 16, generator=g)
    text = torch.randn(16, 7, generator=g)
    y = torch.randint(0, 7, (5,), generator=g)
    assert torch.equal(model.l2(e, t, "none"), oa.l2(e, t, "none"))
    assert torch.equal(model.l2(e, t, "mean"), oa.l2(e, t, "mean"))
    assert torch.equal(model.compute_loss("ce", e, y, None, 100.0, text, "none"),
                       oa.compute_loss("ce", e, y, None, 100.0, text, "none"))
    assert model.compute_acc(e, y) == oa.compute_acc(e, y)


def test_vlm_attack_signatures_and_masks():
    """vlm_eval/attacks/{pgd,apgd}.py call surface; `_set_mask` needs no GPU, so its semantics are
    compared with the oracle restatement (itself pinned on the live reference) here."""
    from oracle import vlm_attacks as ov
    from robustvlm_b200 import vlm_attacks as va
    assert list(inspect.signature(va.PGD.__init__).parameters)[1:] == [
        "forward", "norm", "eps", "mode", "mask_out", "image_space"]            # pgd.py:10
    assert list(inspect.signature(va.PGD.perturb).parameters)[1:] == [
        "data_clean", "iterations", "stepsize", "perturbation", "verbose", "return_loss"]  # :21
    assert list(inspect.signature(va.APGD.__init__).parameters)[1:] == [
        "model", "norm", "eps", "mask_out", "initial_stepsize", "decrease_every",
        "decrease_every_max", "random_init"]                                    # apgd.py:6
    assert list(inspect.signature(va.apgd).parameters) == [
        "model", "x", "norm", "eps", "n_iter", "use_rs", "mask", "alpha", "n_iter_2", "n_iter_min",
        "pert_init", "verbose", "is_train"]                                     # apgd.py:163-164
    x = torch.rand(1, 4, 3, 8, 8)
    for mo in ("context", "query", 0, 2, -1, None):
        assert torch.equal(va._set_mask(x, mo), ov.set_mask(x, mo)), mo
    assert va.APGD(None, "linf", 1.0, mask_out="none").mask_out is None          # apgd.py:18-21
    for bad in ("bogus", 1.5):
        with pytest.raises(NotImplementedError):
            va._set_mask(x, bad)
    with pytest.raises(nat.RvlmError):                                           # no CPU path
        va.PGD(lambda v: v.sum(), "linf", 4 / 255, mask_out=None).perturb(x, 1, 1 / 255)


def test_apgd_schedule_matches_oracle_checkpoints():
    """The host passes `k_check` to the bookkeeping kernel on exactly the iterations the
    reference's counter3 == k test fires (train/apgd_train.py:151-154,330-352)."""
    from oracle import attacks as oa
    from robustvlm_b200 import attacks
    for n_iter in (1, 5, 10, 50, 100, 1000):
        assert attacks.apgd_schedule(n_iter) == oa.apgd_schedule(n_iter)
        k, k_min, k_decr = attacks.apgd_schedule(n_iter)
        fired, c = [], 0
        for i in range(n_iter):                    # the loop of attacks._apgd_core
            c += 1
            if c == k:
                fired.append(i)
                c, k = 0, max(k - k_decr, k_min)
        assert fired == oa.checkpoints(n_iter)


def test_trainer_surface_names_the_reference_options():
    """AdversarialTrainer keyword arguments = the argparse options of the reference that reach
    train_one_epoch (train/adversarial_training_clip.py:31-64)."""
    from robustvlm_b200.train import AdversarialTrainer
    p = inspect.signature(AdversarialTrainer.__init__).parameters
    for name in ("lr", "wd", "opt", "momentum_sgd", "steps", "warmup", "loss", "loss_clean",
                 "clean_weight", "trades", "inner_loss", "norm", "eps", "stepsize", "iterations",
                 "attack", "output_normalize", "start_step"):
        assert name in p, name
    assert p["opt"].default == "adamw" and p["betas"].default == (0.9, 0.999)     # torch defaults, :197
    assert p["attack"].default == "pgd" and p["clean_weight"].default == 0.0
    for m in ("train_step", "evaluate", "save", "optimizer_state_dict", "load_optimizer_state_dict",
              "load_model_state_dict"):
        assert callable(getattr(AdversarialTrainer, m))
