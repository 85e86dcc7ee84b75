"""Python loader of the CPU oracle (TEST INFRASTRUCTURE).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def load():
    """CApi-compatible handle on the oracle (prefix `oracle_`, no ctx)."""
    import ratilqr_b200  # struct definitions and the generic binder live with the C ABI
    from ratilqr_b200._capi import CApi
    if not os.path.exists(LIB_PATH):
        build()
    dll = ctypes.CDLL(LIB_PATH)
    api = CApi(dll, "oracle_", needs_ctx=False)
    api.raw = dll
    return api
