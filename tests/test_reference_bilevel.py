"""test/cross_entropy_bilevel_optimization_test.jl and test/nelder_mead_bilevel_optimization_test.jl of the
reference re-expressed against the host mirror; plus whole-solve parity of the host loops (Python + batched
device fan-out) against the oracle's independent C++ restatement of the same Julia files."""
import ctypes as C
import math

import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import cross_entropy as CE
from ratilqr_b200 import nelder_mead as NM
from ratilqr_b200 import workloads as wl
from ratilqr_b200._capi import IleqgOpts, ProblemDesc, make_opts

dp = C.POINTER(C.c_double)


class OracleCEOpts(C.Structure):
    _fields_ = [("mu_init", C.c_double), ("sigma_init", C.c_double), ("num_samples", C.c_int32), ("num_elite", C.c_int32),
                ("iter_max", C.c_int32), ("lam", C.c_double), ("use_theta_max", C.c_int32)]


class OracleNMOpts(C.Structure):
    _fields_ = [("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double), ("eps", C.c_double), ("lam", C.c_double),
                ("iter_max", C.c_int32), ("theta_high_init", C.c_double), ("theta_low_init", C.c_double),
                ("c_high", C.c_double), ("c_low", C.c_double), ("has_c_high", C.c_int32), ("has_c_low", C.c_int32)]


def _problem():
    prob, x0, u = wl.c1_problem()
    return prob, x0, [u[:, k].copy() for k in range(u.shape[1])]


def test_ce_reference_suite(backend):
    """cross_entropy_bilevel_optimization_test.jl:10-41"""
    prob, x0, u_array = _problem()
    solver = R.CrossEntropyBilevelOptimizationSolver(num_samples=3, backend=backend)
    CE.initialize_(solver)
    theta_array = [0.1, 0.3, 0.43]
    costs = CE.compute_cost(solver, prob, x0, u_array, theta_array, 1.0)          # :27
    costs_test = CE.compute_cost_serial(solver, prob, x0, u_array, theta_array, 1.0)
    assert np.allclose(costs, costs_test)                                         # :30
    # BASELINE.md section 4 anchors (survey-time numpy restatement, not Julia)
    assert np.allclose(costs, [11.002908466254208, 4.33624364124029, 3.3284929065983375], rtol=1e-12)
    th = CE.get_positive_samples(0.0, 1.0, 10, np.random.default_rng(123))        # :32-33
    assert np.all(th > 0.0) and len(th) == 10
    rng = np.random.default_rng(12344)
    theta_opt, x_array, l_array, L_array, c_opt, th_min, th_max = CE.solve_(solver, prob, x0, u_array, rng, kl_bound=1.0)
    assert not math.isinf(c_opt) and not math.isnan(theta_opt)                    # :38-39
    assert 0 < th_min <= th_max
    # kl_bound == 0 reduces to iLQG (:386-389,408)
    t0, _, _, _, v0, a, b = CE.solve_(solver, prob, x0, u_array, rng, kl_bound=0.0)
    assert t0 == 0.0 and a == 0.0 and b == 0.0 and np.isclose(v0, 1.0029075497782471, rtol=1e-12)


def test_nm_reference_suite(backend):
    """nelder_mead_bilevel_optimization_test.jl:11-32"""
    prob, x0, u_array = _problem()
    nm = R.NelderMeadBilevelOptimizationSolver(iter_max=20, eps=1e-3, theta_high_init=10.0, theta_low_init=1e-8, backend=backend)
    theta_opt, x_array, l_array, L_array, c_opt = NM.solve_(nm, prob, x0, u_array, kl_bound=1.0)
    assert not math.isinf(c_opt) and not math.isnan(theta_opt)                    # :25-26
    c_low_init = NM.compute_cost_worker(nm, prob, x0, u_array, nm.theta_low_init, 1.0)
    c_high_init = NM.compute_cost_worker(nm, prob, x0, u_array, nm.theta_high_init, 1.0)
    assert not math.isinf(c_low_init) and not math.isinf(c_high_init)              # :29
    assert c_opt <= c_low_init and c_opt <= c_high_init                            # :30-31
    # SURVEY Appendix B anchor: 5 NM iterations, theta_opt = 29.99999998, c_opt = 1.0367995862438515
    assert nm.iter_current == 5 and np.isclose(theta_opt, 29.99999998, rtol=1e-9) and np.isclose(c_opt, 1.0367995862438515, rtol=1e-9)


def test_nm_speculative_equals_serial(backend):
    prob, x0, u_array = _problem()
    res = []
    for spec in (True, False):
        nm = R.NelderMeadBilevelOptimizationSolver(iter_max=20, eps=1e-3, theta_high_init=10.0, theta_low_init=1e-8,
                                                   backend=backend, speculative=spec)
        out = NM.solve_(nm, prob, x0, u_array, kl_bound=1.0)
        res.append((out[0], out[4], nm.iter_current, nm.n_evals, nm.theta_high, nm.c_high, nm.c_low))
    assert res[0] == res[1]  # identical simplex history, vertices and evaluation count


def _desc(spec):
    return spec.desc()


@pytest.mark.parametrize("kl", [1.0, 0.05])
def test_ce_whole_solve_matches_oracle_restatement(backend, oracle_be, kl):
    """Host CE loop (Python) + batched fan-out vs the oracle's C++ restatement of solve! with the same
    injected standard-normal stream (Julia's rng cannot be reproduced)."""
    prob, x0, u_array = _problem()
    spec = prob.spec()
    z = np.random.Generator(np.random.Philox(key=99)).standard_normal(4000)
    solver = R.CrossEntropyBilevelOptimizationSolver(num_samples=8, num_elite=3, iter_max=4, mu_init=20.0, sigma_init=15.0,
                                                     backend=backend)
    stream = CE.InjectedNormals(z)
    got = CE.solve_(solver, prob, x0, u_array, stream, kl_bound=kl)
    # oracle side
    d = _desc(spec)
    opts = make_opts()
    ce = OracleCEOpts(20.0, 15.0, 8, 3, 4, 0.5, 0)
    n, m, N = spec.n, spec.m, spec.N
    outs = [C.c_double() for _ in range(6)]
    nz = C.c_int64()
    st = C.c_int32()
    x = np.zeros((n, N + 1), order="F"); l = np.zeros((m, N), order="F"); L = np.zeros((m, n, N), order="F")
    u = np.ascontiguousarray(np.stack(u_array, axis=-1).ravel(order="F"))
    f = oracle_be.raw.oracle_ce_solve
    f.restype = C.c_int32
    rc = f(C.byref(d), C.byref(opts), C.byref(ce), x0.ctypes.data_as(dp), u.ctypes.data_as(dp), C.c_double(kl),
           z.ctypes.data_as(dp), C.c_int64(z.size), *[C.byref(o) for o in outs[:6]], C.byref(nz),
           x.ctypes.data_as(dp), l.ctypes.data_as(dp), L.ctypes.data_as(dp), C.byref(st))
    assert rc == 0 and st.value == 0
    theta_opt, value, th_min, th_max, mu, sigma = [o.value for o in outs]
    assert stream.i == nz.value                         # same number of draws consumed => same redraw history
    assert np.isclose(got[0], theta_opt, rtol=1e-9) and np.isclose(got[4], value, rtol=1e-9)
    assert np.isclose(got[5], th_min, rtol=1e-12) and np.isclose(got[6], th_max, rtol=1e-12)
    assert np.isclose(solver.mu, mu, rtol=1e-9) and np.isclose(solver.sigma, sigma, rtol=1e-9)
    assert np.isclose(solver.mu_init, ce.mu_init) and np.isclose(solver.sigma_init, ce.sigma_init)
    assert np.allclose(np.stack(got[1], -1), x, rtol=1e-9, atol=1e-12)
    assert np.allclose(np.stack(got[3], -1), L, rtol=1e-9, atol=1e-12)


def test_nm_whole_solve_matches_oracle_restatement(backend, oracle_be):
    prob, x0, u_array = _problem()
    spec = prob.spec()
    nm = R.NelderMeadBilevelOptimizationSolver(iter_max=20, eps=1e-3, theta_high_init=10.0, theta_low_init=1e-8, backend=backend)
    got = NM.solve_(nm, prob, x0, u_array, kl_bound=1.0)
    d = _desc(spec)
    opts = make_opts()
    o = OracleNMOpts(1.0, 2.0, 0.5, 1e-3, 0.5, 20, 10.0, 1e-8, 0.0, 0.0, 0, 0)
    n, m, N = spec.n, spec.m, spec.N
    th, val = C.c_double(), C.c_double()
    it, ev, st = C.c_int32(), C.c_int32(), C.c_int32()
    x = np.zeros((n, N + 1), order="F"); l = np.zeros((m, N), order="F"); L = np.zeros((m, n, N), order="F")
    u = np.ascontiguousarray(np.stack(u_array, axis=-1).ravel(order="F"))
    f = oracle_be.raw.oracle_nm_solve
    f.restype = C.c_int32
    rc = f(C.byref(d), C.byref(opts), C.byref(o), x0.ctypes.data_as(dp), u.ctypes.data_as(dp), C.c_double(1.0),
           C.byref(th), C.byref(val), C.byref(it), C.byref(ev), x.ctypes.data_as(dp), l.ctypes.data_as(dp),
           L.ctypes.data_as(dp), C.byref(st))
    assert rc == 0 and st.value == 0
    assert nm.iter_current == it.value and nm.n_evals == ev.value == 12  # 12 cost evaluations
    assert np.isclose(got[0], th.value, rtol=1e-12) and np.isclose(got[4], val.value, rtol=1e-9)
    assert np.isclose(nm.theta_high_init, o.theta_high_init) and np.isclose(nm.c_low, o.c_low, rtol=1e-9)
    assert np.allclose(np.stack(got[1], -1), x, rtol=1e-9, atol=1e-12)
    # second solve! on the same solver reuses the stale vertex costs (reference quirk)
    got2 = NM.solve_(nm, prob, x0, u_array, kl_bound=1.0)
    rc = f(C.byref(d), C.byref(opts), C.byref(o), x0.ctypes.data_as(dp), u.ctypes.data_as(dp), C.c_double(1.0),
           C.byref(th), C.byref(val), C.byref(it), C.byref(ev), x.ctypes.data_as(dp), l.ctypes.data_as(dp),
           L.ctypes.data_as(dp), C.byref(st))
    assert np.isclose(got2[0], th.value, rtol=1e-12) and np.isclose(got2[4], val.value, rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("kl,mu0,sg0", [(1.0, 60.0, 30.0), (1.0, 600.0, 300.0), (0.0, 60.0, 30.0)])
def test_ce_fleet_on_device_matches_oracle_per_problem(gpu_be, oracle_be, kl, mu0, sg0):
    """ratilqr_ce_solve_fleet (whole RAT iLQR loop on the device, P problems in lock-step) vs P independent runs of the
    oracle's restatement of solve! with the same injected normal streams."""
    P = 6
    prob, cps, x0, u = wl.fleet(P, N=20)
    spec = prob.spec(cost_params=cps)
    z = np.random.Generator(np.random.Philox(key=17)).standard_normal((P, 3000))
    # (600, 300): most first draws are infeasible => per-problem shrink / redraw rounds of different lengths
    g = gpu_be.ce_solve_fleet(spec, x0, u, kl, mu0, sg0, num_samples=8, num_elite=3, iter_max=3, z_inject=z)
    f = oracle_be.raw.oracle_ce_solve
    f.restype = C.c_int32
    n, m, N = spec.n, spec.m, spec.N
    opts = make_opts()
    uflat = np.ascontiguousarray(u.ravel(order="F"))
    for p in range(P):
        sp = prob.spec(cost_params=cps[p])
        d = sp.desc()
        ce = OracleCEOpts(mu0, sg0, 8, 3, 3, 0.5, 0)
        outs = [C.c_double() for _ in range(6)]
        nz, st = C.c_int64(), C.c_int32()
        x = np.zeros((n, N + 1), order="F"); l = np.zeros((m, N), order="F"); L = np.zeros((m, n, N), order="F")
        x0p = np.ascontiguousarray(x0[:, p])
        rc = f(C.byref(d), C.byref(opts), C.byref(ce), x0p.ctypes.data_as(dp), uflat.ctypes.data_as(dp), C.c_double(kl),
               z[p].ctypes.data_as(dp), C.c_int64(z.shape[1]), *[C.byref(o) for o in outs], C.byref(nz),
               x.ctypes.data_as(dp), l.ctypes.data_as(dp), L.ctypes.data_as(dp), C.byref(st))
        assert rc == 0 and st.value == 0
        theta_opt, value, th_min, th_max, mu, sigma = [o.value for o in outs]
        assert g["nz_used"][p] == nz.value, p            # same draw / redraw history
        assert np.isclose(g["theta_opt"][p], theta_opt, rtol=1e-9) and np.isclose(g["value"][p], value, rtol=1e-9)
        assert np.isclose(g["theta_min"][p], th_min, rtol=1e-12) and np.isclose(g["theta_max"][p], th_max, rtol=1e-12)
        assert np.isclose(g["mu_init"][p], ce.mu_init) and np.isclose(g["sigma_init"][p], ce.sigma_init)
        if kl > 0:
            assert np.isclose(g["mu"][p], mu, rtol=1e-9) and np.isclose(g["sigma"][p], sigma, rtol=1e-9)
        assert np.allclose(g["x"][..., p], x, rtol=1e-9, atol=1e-12) and np.allclose(g["L"][..., p], L, rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_ce_fleet_philox_matches_injected_statistically(gpu_be):
    """north_star: "match statistically (same elite theta distribution) with on-device RNG".  The same problem replicated
    1,536 times: the CE runs driven by on-device Philox and the runs driven by injected host normals are two samples of
    one distribution of (theta_opt, mu, sigma); compare their moments within sampling error."""
    P = 1536
    prob, x0, u = wl.c2_problem(N=20)
    spec = prob.spec()
    x0p = np.tile(x0[:, None], (1, P))
    z = np.random.Generator(np.random.Philox(key=123)).standard_normal((P, 400))
    kw = dict(num_samples=10, num_elite=3, iter_max=3, want=())
    a = gpu_be.ce_solve_fleet(spec, x0p, u, 0.1, 1.0, 2.0, seed=2024, **kw)
    b = gpu_be.ce_solve_fleet(spec, x0p, u, 0.1, 1.0, 2.0, z_inject=z, **kw)
    c = gpu_be.ce_solve_fleet(spec, x0p, u, 0.1, 1.0, 2.0, seed=2024, **kw)
    assert np.array_equal(a["theta_opt"], c["theta_opt"])  # same seed: bit-reproducible
    assert np.std(a["theta_opt"]) > 0 and not np.array_equal(a["theta_opt"], b["theta_opt"])
    for f in ("theta_opt", "mu", "sigma", "value"):
        xa, xb = a[f], b[f]
        se = np.sqrt(xa.var() / P + xb.var() / P)
        assert abs(xa.mean() - xb.mean()) < 5 * se, f
        assert abs(np.log(xa.std() / xb.std())) < 0.15, f
    # the per-problem streams are distinct: problems do not share their draws
    assert np.unique(a["theta_opt"]).size > 0.9 * P


@pytest.mark.gpu
def test_nm_fleet_on_device_matches_oracle_per_problem(gpu_be, oracle_be):
    """ratilqr_nm_solve_fleet (RAT iLQR++ for P problems in lock-step, speculative candidates) vs P independent runs of the
    oracle's restatement of solve!, including the persistence of the vertex costs across a second call."""
    P = 7
    prob, cps, x0, u = wl.fleet(P, N=20)
    spec = prob.spec(cost_params=cps)
    kw = dict(eps=1e-3, iter_max=12, theta_high_init=300.0, theta_low_init=1e-8)  # 300: infeasible => theta_high halves first
    g1 = gpu_be.nm_solve_fleet(spec, x0, u, 0.5, **kw)
    g2 = gpu_be.nm_solve_fleet(spec, x0, u, 0.5, state=g1["state"], **kw)
    f = oracle_be.raw.oracle_nm_solve
    f.restype = C.c_int32
    n, m, N = spec.n, spec.m, spec.N
    opts = make_opts()
    uflat = np.ascontiguousarray(u.ravel(order="F"))
    for p in range(P):
        sp = prob.spec(cost_params=cps[p])
        d = sp.desc()
        o = OracleNMOpts(1.0, 2.0, 0.5, 1e-3, 0.5, 12, 300.0, 1e-8, 0.0, 0.0, 0, 0)
        x0p = np.ascontiguousarray(x0[:, p])
        for g in (g1, g2):
            th, val = C.c_double(), C.c_double()
            it, ev, st = C.c_int32(), C.c_int32(), C.c_int32()
            x = np.zeros((n, N + 1), order="F"); l = np.zeros((m, N), order="F"); L = np.zeros((m, n, N), order="F")
            rc = f(C.byref(d), C.byref(opts), C.byref(o), x0p.ctypes.data_as(dp), uflat.ctypes.data_as(dp), C.c_double(0.5),
                   C.byref(th), C.byref(val), C.byref(it), C.byref(ev), x.ctypes.data_as(dp), l.ctypes.data_as(dp),
                   L.ctypes.data_as(dp), C.byref(st))
            assert rc == 0 and st.value == g["status"][p] == 0
            assert g["nm_iters"][p] == it.value and g["n_evals"][p] == ev.value, (p, g["nm_iters"][p], it.value, g["n_evals"][p], ev.value)
            assert np.isclose(g["theta_opt"][p], th.value, rtol=1e-12) and np.isclose(g["value"][p], val.value, rtol=1e-9)
            assert np.isclose(g["state"]["theta_high_init"][p], o.theta_high_init) and np.isclose(g["state"]["c_low"][p], o.c_low, rtol=1e-9)
            assert np.allclose(g["x"][..., p], x, rtol=1e-9, atol=1e-12) and np.allclose(g["L"][..., p], L, rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_fleet_mpc_drives_every_system_to_its_goal(gpu_be):
    """receding-horizon RAT iLQR for a small fleet: warm-started plans, persisted CE state, noisy true system"""
    from ratilqr_b200.mpc import run_fleet_mpc
    P = 12
    prob, cps, x0, u = wl.fleet(P, N=20)
    out = run_fleet_mpc(gpu_be, prob, cps, x0, steps=45, kl_bound=0.1, rng=np.random.default_rng(3))
    goals = cps[:, 5:7]                                   # xg(0:2) of every problem's cost block
    d0 = np.linalg.norm(x0[:2].T - goals, axis=1)
    d1 = np.linalg.norm(out["x"][:2, -1].T - goals, axis=1)
    assert np.all(d1 < 0.35 * d0) and np.all(np.isfinite(out["value"]))  # every vehicle closed most of its distance
    assert np.all(out["theta"] > 0) and np.all(out["mu_init"] > 0)
    assert np.median(out["ms"][5:]) < np.median(out["ms"][:3]) * 1.5     # warm starts do not make steps slower
