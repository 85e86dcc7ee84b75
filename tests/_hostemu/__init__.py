"""Loader of the test-only host emulation of the device arithmetic (see hostemu.cpp)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhostemu.so")


def load():
    import ratilqr_b200  # noqa: F401
    from ratilqr_b200._capi import CApi
    src = [os.path.join(_HERE, "hostemu.cpp")] + [os.path.join(_HERE, "../../ratilqr.jl_b200/csrc", f)
                                                  for f in ("rl_core.cuh", "rl_components.cuh", "rl_coop.cuh", "rl_coop2.cuh", "rl_spec.cuh", "rl_user.cuh", "rl_host.hpp")]
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(f) for f in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return CApi(ctypes.CDLL(LIB_PATH), "hostemu_", needs_ctx=False)


def load_variant(name):
    """another build of the same emulation (Makefile target lib<name>.so), e.g. hostemu_coop_rolled"""
    import ratilqr_b200  # noqa: F401
    from ratilqr_b200._capi import CApi
    path = os.path.join(_HERE, f"lib{name}.so")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s", f"lib{name}.so"])
    return CApi(ctypes.CDLL(path), "hostemu_", needs_ctx=False)
