"""ctypes binding of liberl_gp_b200.so (the C ABI declared in include/erl_gp_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is visible, every entry
point raises.  The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# ERL_GP_B200_LIB: another build of the same library (kernel experiments: timing / A-B variants), never a fallback
LIB_PATH = os.environ.get("ERL_GP_B200_LIB") or os.path.join(_HERE, "lib", "liberl_gp_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "erl_gp_b200.h")

STATUS_OK = 0
KERNEL_OU, KERNEL_MATERN32, KERNEL_RBF = 0, 1, 2
KERNELS = {"ou": KERNEL_OU, "matern32": KERNEL_MATERN32, "rbf": KERNEL_RBF}
MAPPING_NONE, MAPPING_IDENTITY, MAPPING_INVERSE, MAPPING_INVERSE_SQRT = -1, 0, 1, 2
MAPPING_EXP, MAPPING_LOG, MAPPING_TANH, MAPPING_SIGMOID = 3, 4, 5, 6


class ErlGpError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: status {status}" + (f" ({detail})" if detail else ""))


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ErlGpError(-1, "load", f"{LIB_PATH} not built; run __graft_entry__.build() — there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.erl_gp_status_string.restype = C.c_char_p
        _lib.erl_gp_context_last_error.restype = C.c_char_p
        _lib.erl_gp_context_last_error.argtypes = [C.c_void_p]
    return _lib


def declared_symbols() -> list[str]:
    """Every function include/erl_gp_b200.h declares (used by the export test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(erl_gp_[a-z0-9_]+)\s*\(", text)))


def check(status: int, where: str, ctx=None):
    if status != STATUS_OK:
        lib = load()
        detail = lib.erl_gp_status_string(status).decode()
        if ctx:
            msg = lib.erl_gp_context_last_error(ctx).decode()
            if msg:
                detail += ": " + msg
        raise ErlGpError(status, where, detail)


def device_count() -> int:
    n = C.c_int(0)
    load().erl_gp_device_count(C.byref(n))
    return n.value
