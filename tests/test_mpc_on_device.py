"""ratilqr_mpc_fleet_run (receding-horizon RAT iLQR, every step on the device: SURVEY.md 8f-1) against an ORACLE-CHAINED
reference: repeated oracle_ce_solve (the C++ restatement of solve!, cross_entropy_bilevel_optimization.jl:364-415) with the
persisted mu_init / sigma_init (:66-68), the oracle's one-stage rollout for the true system, injected disturbances and
injected theta-draw normals."""
import ctypes as C

import numpy as np
import pytest

from ratilqr_b200 import workloads as wl
from ratilqr_b200._capi import Spec, make_opts
from ratilqr_b200.mpc import run_fleet_mpc, run_fleet_mpc_on_device
from tests.test_reference_bilevel import OracleCEOpts

dp = C.POINTER(C.c_double)
pytestmark = pytest.mark.gpu


def test_mpc_fleet_on_device_matches_oracle_chain(gpu_be, oracle_be):
    P, steps, N, nz = 8, 5, 20, 600
    prob, cps, x0, u = wl.fleet(P, N=N)
    n, m = 4, 2
    rng = np.random.Generator(np.random.Philox(key=515))
    z = rng.standard_normal((steps, P, nz))
    chol = np.linalg.cholesky(np.asarray(prob.W(0)))
    noise = np.einsum("ij,jtp->itp", chol, rng.standard_normal((n, steps, P)))
    g = run_fleet_mpc_on_device(gpu_be, prob, cps, x0, steps, kl_bound=0.1, noise=noise, z_inject=z)
    assert g["x"].shape == (n, steps + 1, P) and np.array_equal(g["x"][:, 0], x0)
    f = oracle_be.raw.oracle_ce_solve
    f.restype = C.c_int32
    opts = make_opts()
    for p in range(P):
        sp = prob.spec(cost_params=cps[p])
        d = sp.desc()
        step_spec = Spec(sp.model_id, sp.cost_id, n, m, 1, sp.model_params, sp.cost_params, np.eye(n))
        ce = OracleCEOpts(1.0, 2.0, 10, 3, 5, 0.5, 0)   # mu_init / sigma_init are in-out: they persist across the steps
        x = np.ascontiguousarray(x0[:, p])
        plan = np.zeros((m, N))
        for t in range(steps):
            outs = [C.c_double() for _ in range(6)]
            nzu, st = C.c_int64(), C.c_int32()
            xs = np.zeros((n, N + 1), order="F"); l = np.zeros((m, N), order="F"); L = np.zeros((m, n, N), order="F")
            pf = np.ascontiguousarray(plan.ravel(order="F"))
            zt = np.ascontiguousarray(z[t, p])
            rc = f(C.byref(d), C.byref(opts), C.byref(ce), x.ctypes.data_as(dp), pf.ctypes.data_as(dp), C.c_double(0.1),
                   zt.ctypes.data_as(dp), C.c_int64(nz), *[C.byref(o) for o in outs], C.byref(nzu),
                   xs.ctypes.data_as(dp), l.ctypes.data_as(dp), L.ctypes.data_as(dp), C.byref(st))
            assert rc == 0 and st.value == 0
            theta_opt, value = outs[0].value, outs[1].value
            assert np.isclose(g["theta"][t, p], theta_opt, rtol=1e-8), (p, t)
            assert np.isclose(g["value"][t, p], value, rtol=1e-8), (p, t)
            assert np.allclose(g["u"][:, t, p], l[:, 0], rtol=1e-7, atol=1e-9), (p, t)
            xn, sst = oracle_be.rollout_open(step_spec, x, l[:, :1].reshape(m, 1, 1))
            x = np.ascontiguousarray(xn[:, 1, 0] + noise[:, t, p])
            assert np.allclose(g["x"][:, t + 1, p], x, rtol=1e-8, atol=1e-10), (p, t)
            plan = np.concatenate([l[:, 1:], l[:, -1:]], axis=1)
        assert np.isclose(g["mu_init"][p], ce.mu_init) and np.isclose(g["sigma_init"][p], ce.sigma_init)


def test_mpc_fleet_on_device_philox_and_mixture(gpu_be):
    """on-device disturbances (Philox N(0, W) / the true-model mixture): reproducible, every vehicle approaches its goal,
    and statistically the same closed loop as the host-driven loop of mpc.run_fleet_mpc"""
    P, steps = 24, 40
    prob, cps, x0, u = wl.fleet(P, N=20)
    a = run_fleet_mpc_on_device(gpu_be, prob, cps, x0, steps, kl_bound=0.1, noise_seed=5, seed=3)
    b = run_fleet_mpc_on_device(gpu_be, prob, cps, x0, steps, kl_bound=0.1, noise_seed=5, seed=3)
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["theta"], b["theta"])
    goals = cps[:, 5:7]
    d0 = np.linalg.norm(x0[:2].T - goals, axis=1)
    d1 = np.linalg.norm(a["x"][:2, -1].T - goals, axis=1)
    assert np.all(d1 < 0.4 * d0) and np.all(np.isfinite(a["value"])) and np.all(a["theta"] > 0)
    h = run_fleet_mpc(gpu_be, prob, cps, x0, steps, kl_bound=0.1, rng=np.random.default_rng(3))
    dh = np.linalg.norm(h["x"][:2, -1].T - goals, axis=1)
    assert abs(np.mean(d1 / d0) - np.mean(dh / d0)) < 0.1
    mix = dict(weights=[0.7, 0.3], means=np.array([[0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [-0.02, 0.05]]),
               covs=np.stack([np.diag([1e-3, 1e-3, 1e-4, 1e-3]), np.diag([4e-3, 4e-3, 1e-4, 4e-3])], axis=-1))
    c = run_fleet_mpc_on_device(gpu_be, prob, cps, x0, steps, kl_bound=0.1, noise_seed=5, seed=3, true_mixture=mix)
    assert not np.array_equal(c["x"], a["x"]) and np.all(np.isfinite(c["x"]))
