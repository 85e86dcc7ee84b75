"""The kernel arithmetic compiled with -DRL_FUSED=0 (no fused accumulation, IEEE 1/sqrt, one log per stage) follows
the oracle's canonical order exactly: results must be BIT-IDENTICAL, for the thread-per-instance formulation and for
the warp-cooperative one.  (The shipped kernels use RL_FUSED=1 and are compared at 1e-9 elsewhere.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from ratilqr_b200 import workloads as wl
from ratilqr_b200._capi import CApi

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_hostemu")


@pytest.fixture(scope="module")
def unfused():
    subprocess.check_call(["make", "-C", HERE, "-s", "libhostemu_unfused.so"])
    return CApi(ctypes.CDLL(os.path.join(HERE, "libhostemu_unfused.so")), "hostemu_", needs_ctx=False)


@pytest.mark.parametrize("coop", [0, 1])
def test_unfused_build_is_bit_identical_to_oracle(unfused, oracle_be, coop):
    unfused.dll.hostemu_set_coop(coop)
    try:
        for prob, x0, u, th in ((*wl.c1_problem(), [0.0, 0.1, 0.43, 30.7, 40.0]), (*wl.c2_problem(), wl.c2_thetas(96)),
                                (*wl.c3_problem(N=8), [0.0, 0.05])):
            spec = prob.spec()
            a = oracle_be.ileqg_solve_batch(spec, x0, u, th, eps_hist_cap=64)
            b = unfused.ileqg_solve_batch(spec, x0, u, th, eps_hist_cap=64)
            for k in ("status", "iters", "trials", "restarts"):
                assert np.array_equal(a[k], b[k]), k
            ok = a["status"] == 0
            for k in ("value", "x", "l", "L", "eps_hist"):
                assert np.array_equal(a[k][..., ok], b[k][..., ok]), k
    finally:
        unfused.dll.hostemu_set_coop(0)


def test_trig_cache_build_is_bit_identical_to_the_default_build(hostemu_be):
    """-DRL_TRIG_CACHE=1 (opt-in; slower on the power-capped B200, profiles/r02_trig_cache_negative_ab.txt): the unicycle
    rollout stores (sin psi_k, cos psi_k) of every stage in the workspace record and the backward passes linearise from
    them.  Same routine, same argument => the same bits as recomputing; checked through the host emulation."""
    subprocess.check_call(["make", "-C", HERE, "-s", "libhostemu_trigcache.so"])
    tc = CApi(ctypes.CDLL(os.path.join(HERE, "libhostemu_trigcache.so")), "hostemu_", needs_ctx=False)
    for prob, x0, u, th in ((*wl.c2_problem(), wl.c2_thetas(48)), (*wl.c2_problem(N=7), [0.0, 0.5, 3.0])):
        spec = prob.spec()
        a = hostemu_be.ileqg_solve_batch(spec, x0, u, th, eps_hist_cap=64)
        b = tc.ileqg_solve_batch(spec, x0, u, th, eps_hist_cap=64)
        for k in ("status", "iters", "trials", "restarts", "value", "x", "l", "L", "eps_hist", "mu", "d_current"):
            assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_cooperative_formulation_is_bit_identical_to_thread_per_instance(hostemu_be):
    """DESIGN: every output element of the warp-cooperative stage is accumulated by one lane in the order the thread-per-
    instance stage uses, so the two formulations of the SHIPPED (fused) arithmetic agree bit for bit -- structured small
    models (rolled cooperative stage) and a dense one (restructured stage), through the host emulation."""
    cases = [(*wl.c1_problem(), [0.0, 0.1, 0.43, 30.7]), (*wl.c2_problem(N=12), wl.c2_thetas(24))]
    prob, x0, u = wl.c3_problem(N=6)
    cases.append((prob, x0, u, [0.0, 0.02]))
    for prob, x0, u, th in cases:
        spec = prob.spec()
        a = hostemu_be.ileqg_solve_batch(spec, x0, u, th, eps_hist_cap=64)
        hostemu_be.dll.hostemu_set_coop(1)
        try:
            b = hostemu_be.ileqg_solve_batch(spec, x0, u, th, eps_hist_cap=64)
        finally:
            hostemu_be.dll.hostemu_set_coop(0)
        for k in ("status", "iters", "trials", "restarts", "value", "x", "l", "L", "eps_hist", "mu"):
            assert np.array_equal(a[k], b[k], equal_nan=True), k
