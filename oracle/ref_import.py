"""Importer for the REAL reference code (TEST INFRASTRUCTURE; authoring container only).

/root/reference is mounted read-only here and does not exist on the GPU box, so
this module is used exclusively by (a) `tests/test_oracle_vs_reference.py`
(skipped when the tree is absent) and (b) `tests/golden/make_golden.py`, which
turns reference outputs into committed fixtures.  Nothing is copied: the
reference modules are imported in place.

Three modules the reference imports are absent offline and are stubbed
(SURVEY.md §8c): `open_clip`, `training.scheduler` (needs `cosine_lr`) and
`einops_exts` (needs `rearrange_many`, `repeat_many`).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RVLM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "train"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


def load():
    """Returns a namespace with the reference's hot-path callables."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _stub("open_clip")
    _stub("training")
    _stub("training.scheduler", cosine_lr=lambda *a, **k: None)
    _stub("einops_exts", rearrange_many=lambda *a, **k: None, repeat_many=lambda *a, **k: None)
    ns = types.SimpleNamespace()
    ns.pgd = importlib.import_module("train.pgd_train").pgd
    ns.apgd_train = importlib.import_module("train.apgd_train").apgd_train
    ns.attack_utils = importlib.import_module("vlm_eval.attacks.utils")
    apgd_base = importlib.import_module("autoattack.autopgd_base")
    ns.APGDAttack = apgd_base.APGDAttack
    ns.APGDAttack_targeted = apgd_base.APGDAttack_targeted
    ns.vlm_pgd = importlib.import_module("vlm_eval.attacks.pgd")     # PGD
    ns.vlm_apgd = importlib.import_module("vlm_eval.attacks.apgd")   # APGD, apgd
    try:
        atc = importlib.import_module("train.adversarial_training_clip")
        ns.ClipVisionModel = atc.ClipVisionModel
        ns.ComputeLossWrapper = atc.ComputeLossWrapper
        ns.compute_loss = atc.compute_loss
        ns.l2 = atc.l2
        ns.ce = atc.ce
    except Exception as e:  # noqa: BLE001  (heavy import chain: wandb, torchvision, ...)
        ns.import_error = e
    return ns
