"""ctypes binding of librvlm_b200.so (the C ABI declared in include/rvlm.h).

The product path has no CPU or PyTorch fallback: if the shared library is
missing, or a call returns an error code, we raise.  `torch` is used only for
device memory, streams and `torch.distributed` plumbing.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p, c_size_t

_HERE = os.path.dirname(os.path.abspath(__file__))
# RVLM_LIB: development hook -- load an A/B build of the same library (csrc/Makefile OUTDIR=...)
LIB_PATH = os.environ.get("RVLM_LIB") or os.path.join(_HERE, "_lib", "librvlm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "rvlm.h")

_lib = None


class RvlmError(RuntimeError):
    pass


def declared_symbols(header_path: str = HEADER_PATH) -> list[str]:
    """Every function name the public header declares (used by the CPU tests to
    check the library exports all of them)."""
    txt = open(header_path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rvlm_[a-z0-9_]+)\s*\(", txt)))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RvlmError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback for the product path)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.rvlm_strerror.restype = c_char_p
        _lib.rvlm_strerror.argtypes = [c_int]
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().rvlm_strerror(rc).decode()
        raise RvlmError(f"{what or 'rvlm call'} failed: {rc} ({msg})")


def ptr(t) -> c_void_p:
    """Device (or host) pointer of a torch tensor, or NULL for None."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name: str, *args) -> None:
    fn = getattr(lib(), name)
    fn.restype = c_int
    check(fn(*args), name)


__all__ = [
    "lib", "check", "ptr", "stream_ptr", "call", "declared_symbols", "RvlmError", "LIB_PATH",
    "c_int", "c_float", "c_double", "c_void_p", "c_int64", "c_size_t",
]
