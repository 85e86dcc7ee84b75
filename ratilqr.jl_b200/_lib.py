"""Loader of libratilqr_b200.so (the CUDA product library).  There is NO CPU fallback: if the
shared library is missing or no CUDA device can be opened, every solver call raises."""
import ctypes
import os

from ._capi import CApi, Multi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RATILQR_B200_LIB") or os.path.join(_HERE, "csrc", "libratilqr_b200.so")  # env override: tuning A/B runs

_dll = None
_default = None


class LibraryMissing(RuntimeError):
    pass


def load_library():
    """dlopen the product library (no CUDA call yet, so this also works on a CPU-only box)."""
    global _dll
    if _dll is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(nvcc, sm_100a). There is no CPU fallback.")
        _dll = ctypes.CDLL(LIB_PATH)
    return _dll


def new_backend(device_id=0):
    """A fresh ctx on `device_id` (one ctx per device / per caller thread)."""
    return CApi(load_library(), "ratilqr_", needs_ctx=True).open(device_id)


def new_multi(device_ids):
    """One process, several GPUs: contexts + NCCL communicators (ratilqr_create_multi)."""
    return Multi(load_library(), list(device_ids))


def default_backend():
    global _default
    if _default is None:
        _default = new_backend(int(os.environ.get("LOCAL_RANK", "0")))
    return _default


def set_default_backend(be):
    """Testing seam: lets the test-suite drive the host logic with another C-ABI provider."""
    global _default
    _default = be
