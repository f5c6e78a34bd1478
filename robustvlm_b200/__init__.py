"""robustvlm_b200 — B200-native adversarial-perturbation hot path of chs20/RobustVLM.

The Python names mirror the reference's call surface; the arithmetic
lives in librvlm_b200.so (include/rvlm.h), built in-tree by __graft_entry__.build().
Importing the package does not need a GPU; calling into it does (no CPU fallback).
"""
from . import _native  # noqa: F401

__all__ = ["attacks", "autoattack", "checkpoint", "engine", "model", "parallel", "train",
           "vlm_attacks"]


def __getattr__(name):
    if name in __all__:
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
