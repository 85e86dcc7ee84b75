"""Scheduling never changes results: the heaviest-first slot order taken from the previous call's work profile and the
concurrent sub-fleet blocks of ratilqr_ce_solve_fleet are pure work-placement decisions (rl_capi.cu)."""
import os

import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import workloads as wl


@pytest.mark.gpu
def test_profile_order_does_not_change_results(oracle_be):
    be = R.new_backend(0)  # fresh context: no work profile yet
    try:
        P, K = 96, 6
        prob, cps, x0, u = wl.fleet(P, key=3, N=20)
        spec = prob.spec(cost_params=cps)
        th1 = wl.positive_thetas(P * K, key=1)
        th2 = wl.positive_thetas(P * K, key=2)
        cold = be.ce_costs(spec, x0, u, th2, 0.1, P=P)           # natural order (first call of this context)
        be.ce_costs(spec, x0, u, th1, 0.1, P=P)                  # leaves the profile of another population
        warm = be.ce_costs(spec, x0, u, th2, 0.1, P=P)           # ordered by that profile
        assert np.array_equal(cold[0], warm[0]) and np.array_equal(cold[1], warm[1])
        g = be.ileqg_solve_batch(spec, x0, u, th2, P=P)          # staged after a profile exists: ordered as well
        o = oracle_be.ileqg_solve_batch(spec, x0, u, th2, P=P)
        assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["iters"], o["iters"])
        assert np.max(np.abs(g["x"] - o["x"])) < 1e-9 * max(1.0, np.max(np.abs(o["x"])))
        th1p = th1[:P]                                           # one theta per problem (K = 1): slots follow the problem order
        g1 = be.ileqg_solve_batch(spec, x0, u, th1p, P=P)
        o1 = oracle_be.ileqg_solve_batch(spec, x0, u, th1p, P=P)
        assert np.array_equal(g1["status"], o1["status"]) and np.array_equal(g1["iters"], o1["iters"])
        assert np.max(np.abs(g1["l"] - o1["l"])) < 1e-9 * max(1.0, np.max(np.abs(o1["l"])))
        ref = oracle_be.ce_costs(spec, x0, u, th2, 0.1, P=P)[0]
        fin = np.isfinite(ref)
        assert np.array_equal(fin, np.isfinite(warm[0]))
        assert np.max(np.abs(warm[0][fin] - ref[fin])) < 1e-9 * np.max(np.abs(ref[fin]))
    finally:
        be.close()


@pytest.mark.gpu
@pytest.mark.parametrize("blocks", ["2", "5"])
def test_sub_fleet_blocks_reproduce_the_one_block_solve(blocks):
    """RAT iLQR for a fleet cut into concurrent blocks (own stream, workspace, host thread): identical per-problem results,
    with injected normals and with Philox (streams are indexed by the global problem number)"""
    P = 70  # uneven split
    prob, cps, x0, u = wl.fleet(P, key=11, N=20)
    spec = prob.spec(cost_params=cps)
    z = np.random.default_rng(0).standard_normal((P, 400))
    old = os.environ.get("RATILQR_FLEET_SPLIT")
    res = {}
    try:
        for k in ("1", blocks):
            os.environ["RATILQR_FLEET_SPLIT"] = k
            be = R.new_backend(0)
            try:
                res[k] = (be.ce_solve_fleet(spec, x0, u, 0.1, 1.0, 2.0, num_samples=6, num_elite=2, iter_max=3, z_inject=z),
                          be.ce_solve_fleet(spec, x0, u, 0.1, 1.0, 2.0, num_samples=6, num_elite=2, iter_max=3, seed=5))
            finally:
                be.close()
    finally:
        if old is None:
            os.environ.pop("RATILQR_FLEET_SPLIT", None)
        else:
            os.environ["RATILQR_FLEET_SPLIT"] = old
    for a, b in zip(res["1"], res[blocks]):
        for f in ("theta_opt", "value", "theta_min", "theta_max", "mu", "sigma", "mu_init", "sigma_init", "nz_used", "status",
                  "iters", "x", "l", "L"):
            assert np.array_equal(a[f], b[f]), f


@pytest.mark.gpu
def test_cp_async_staging_is_transparent():
    """the thread-private cp.async staging (two rollout stages in flight over two shared-memory buffers whose contents go
    to registers before the buffer is reused) must not change a single bit: compare with RATILQR_NO_STAGE=1 (plain loads)
    on a batch of several resident waves, i.e. under the memory pressure where a too-early buffer reuse would show"""
    P, K = 160, 1024
    prob, cps, x0, u = wl.fleet(P, key=5)
    spec = prob.spec(cost_params=cps)
    theta = np.concatenate([wl.positive_thetas(K, key=300 + p) for p in range(P)])
    old = os.environ.get("RATILQR_NO_STAGE")
    res = {}
    try:
        for flag in ("0", "1"):
            os.environ["RATILQR_NO_STAGE"] = flag
            be = R.new_backend(0)
            try:
                res[flag] = be.ileqg_solve_batch(spec, x0, u, theta, P=P, want=("l",))
            finally:
                be.close()
    finally:
        if old is None:
            os.environ.pop("RATILQR_NO_STAGE", None)
        else:
            os.environ["RATILQR_NO_STAGE"] = old
    for f in ("status", "iters", "trials", "restarts", "value", "mu", "d_current", "l"):
        assert np.array_equal(res["0"][f], res["1"][f], equal_nan=True) if res["0"][f].dtype.kind == "f" else np.array_equal(res["0"][f], res["1"][f]), f
    assert (res["0"]["status"] == 0).sum() > 0.9 * P * K
