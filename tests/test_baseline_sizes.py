"""Parity at the sizes BASELINE.json's configs state, CUDA library vs the CPU oracle through the C ABI:

  C3  RAT iLQR++ (Nelder-Mead defaults) whole solve on the 12-state quadrotor at T = 40 (the warp-cooperative kernel under
      the NM loop: single problem AND the on-device fleet form), plus a 16-rollout injected-noise Monte Carlo subset
  C4  one PETS iteration on cart-pole at T = 30 with 64 action sequences x 150 particles (5 parameter sets x 30) and
      injected noise; the full 4096 x 150 x 30 population through a size-independent property
  C5  64 RAT iLQR problems at T = 50 with the CE defaults (10 theta x 5 iterations + final solve) vs oracle_ce_solve, plus
      256-sample injected-noise Monte Carlo for 4 of them

Everything here needs the GPU; the oracle side of each case runs in seconds."""
import ctypes as C

import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import nelder_mead as NM
from ratilqr_b200 import workloads as wl
from ratilqr_b200._capi import make_opts
from tests.test_reference_bilevel import OracleCEOpts, OracleNMOpts

dp = C.POINTER(C.c_double)
pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def _oracle_nm(oracle_be, spec, x0, u, kl, o):
    n, m, N = spec.n, spec.m, spec.N
    d = spec.desc()
    opts = make_opts()
    th, val = C.c_double(), C.c_double()
    it, ev, st = C.c_int32(), C.c_int32(), C.c_int32()
    x = np.zeros((n, N + 1), order="F"); l = np.zeros((m, N), order="F"); L = np.zeros((m, n, N), order="F")
    uf = np.ascontiguousarray(u.ravel(order="F"))
    x0f = np.ascontiguousarray(x0)
    f = oracle_be.raw.oracle_nm_solve
    f.restype = C.c_int32
    rc = f(C.byref(d), C.byref(opts), C.byref(o), x0f.ctypes.data_as(dp), uf.ctypes.data_as(dp), C.c_double(kl),
           C.byref(th), C.byref(val), C.byref(it), C.byref(ev), x.ctypes.data_as(dp), l.ctypes.data_as(dp),
           L.ctypes.data_as(dp), C.byref(st))
    assert rc == 0
    return dict(theta_opt=th.value, value=val.value, nm_iters=it.value, n_evals=ev.value, status=st.value, x=x, l=l, L=L)


def test_c3_nm_quadrotor_T40_single_problem(gpu_be, oracle_be):
    """configs[2]: RAT iLQR++ on the quadrotor, T = 40, NM defaults (nelder_mead...jl:112-119), kl = 0.1"""
    prob, x0, u = wl.c3_problem()
    assert prob.N == 40
    spec = prob.spec()
    nm = R.NelderMeadBilevelOptimizationSolver(backend=gpu_be)  # defaults: theta_high_init = 3.0 is infeasible => halvings
    got = NM.solve_(nm, prob, x0, [u[:, k].copy() for k in range(40)], kl_bound=0.1)
    o = OracleNMOpts(1.0, 2.0, 0.5, 1e-2, 0.5, 100, 3.0, 1e-8, 0.0, 0.0, 0, 0)
    ref = _oracle_nm(oracle_be, spec, x0, u, 0.1, o)
    assert ref["status"] == 0
    assert nm.iter_current == ref["nm_iters"] and nm.n_evals == ref["n_evals"]
    assert np.isclose(got[0], ref["theta_opt"], rtol=1e-12) and np.isclose(got[4], ref["value"], rtol=1e-9)
    assert np.isclose(nm.theta_high_init, o.theta_high_init, rtol=1e-15) and np.isclose(nm.c_low, o.c_low, rtol=1e-9)
    assert relerr(np.stack(got[1], -1), ref["x"]) < 1e-9 and relerr(np.stack(got[2], -1), ref["l"]) < 1e-9
    assert relerr(np.stack(got[3], -1), ref["L"]) < 1e-9
    # 16-rollout Monte Carlo subset with an injected noise tensor (the config's 256 rollouts run with Philox)
    rng = np.random.Generator(np.random.Philox(key=16))
    w = np.einsum("ij,jks->iks", np.linalg.cholesky(prob.W(0)), rng.standard_normal((12, 40, 16)))
    g = gpu_be.mc_rollout(spec, ref["x"], ref["l"], ref["L"], 16, noise=w, want_x=True, theta_risk=0.01)
    oo = oracle_be.mc_rollout(spec, ref["x"], ref["l"], ref["L"], 16, noise=w, want_x=True, theta_risk=0.01)
    assert relerr(g["J"], oo["J"]) < 1e-9 and relerr(g["x"], oo["x"]) < 1e-9 and relerr(g["stats"], oo["stats"]) < 1e-9
    # the full 256 rollouts with on-device noise: finite, reproducible, mean within sampling error of the injected run
    p = gpu_be.mc_rollout(spec, ref["x"], ref["l"], ref["L"], 256, seed=3)
    p2 = gpu_be.mc_rollout(spec, ref["x"], ref["l"], ref["L"], 256, seed=3)
    assert np.array_equal(p["J"], p2["J"]) and np.all(np.isfinite(p["J"]))
    assert abs(p["stats"][0, 0] - oo["stats"][0, 0]) < 6 * np.sqrt(p["stats"][0, 1] / 16 + p["stats"][0, 1] / 256)


def test_c3_nm_quadrotor_T40_fleet_on_device(gpu_be, oracle_be):
    """the on-device NM loop (ratilqr_nm_solve_fleet) over the warp-cooperative kernel: 5 quadrotor problems with
    different waypoints, each checked against its own oracle run"""
    prob, x0, u = wl.c3_problem()
    spec0 = prob.spec()
    P = 5
    rng = np.random.Generator(np.random.Philox(key=33))
    cps = np.tile(np.asarray(spec0.cost_params, float), (P, 1))
    cps[:, 5:8] = 1.0 + 0.4 * rng.uniform(-1, 1, (P, 3))  # xg(0:3): the waypoint
    x0s = np.tile(x0[:, None], (1, P))
    x0s[:3] += 0.1 * rng.uniform(-1, 1, (3, P))
    spec = prob.spec(cost_params=cps)
    g = gpu_be.nm_solve_fleet(spec, x0s, u, 0.1)
    for p in range(P):
        o = OracleNMOpts(1.0, 2.0, 0.5, 1e-2, 0.5, 100, 3.0, 1e-8, 0.0, 0.0, 0, 0)
        ref = _oracle_nm(oracle_be, prob.spec(cost_params=cps[p]), x0s[:, p], u, 0.1, o)
        assert ref["status"] == g["status"][p] == 0
        assert g["nm_iters"][p] == ref["nm_iters"] and g["n_evals"][p] == ref["n_evals"], p
        assert np.isclose(g["theta_opt"][p], ref["theta_opt"], rtol=1e-12) and np.isclose(g["value"][p], ref["value"], rtol=1e-9)
        assert relerr(g["x"][..., p], ref["x"]) < 1e-9 and relerr(g["L"][..., p], ref["L"]) < 1e-9


def test_c4_pets_T30_64_sequences_150_particles(gpu_be, oracle_be):
    """configs[3] at its horizon and particle count: N = 30, 150 particles = 5 parameter sets x 30, injected action
    samples and noise; costs, elites and the refitted distribution of one iteration (pets.jl:128-191)"""
    prob, x0 = wl.c4_problem()
    assert prob.N == 30
    spec = prob.spec()
    gen = prob.f_stochastic.gen()
    Cn, Kp, ne = 64, 150, 6
    rng = np.random.Generator(np.random.Philox(key=4096))
    chol = np.linalg.cholesky(prob.f_stochastic.W)
    ctrl = 2.0 * rng.standard_normal((1, 30, Cn))
    noise = np.einsum("ij,jkpc->ikpc", chol, rng.standard_normal((4, 30, Kp, Cn)))
    g = gpu_be.pets_costs(spec, x0, ctrl, Kp, noise=noise, gen=gen)
    o = oracle_be.pets_costs(spec, x0, ctrl, Kp, noise=noise, gen=gen)
    assert relerr(g, o) < 1e-9
    assert np.array_equal(np.argsort(g, kind="stable")[:ne], np.argsort(o, kind="stable")[:ne])  # same elites
    z = rng.standard_normal((1, 30, Cn, 1))
    mu0, Sg0 = np.zeros((1, 30)), np.tile(np.array([[4.0]])[:, :, None], (1, 1, 30))
    oo = oracle_be.pets_solve(spec, x0, mu0, Sg0, Cn, Kp, ne, 1, 0.1, z_inject=z, noise=noise[..., None], gen=gen)
    gg = gpu_be.pets_solve(spec, x0, mu0, Sg0, Cn, Kp, ne, 1, 0.1, z_inject=z, noise=noise[..., None], gen=gen)
    assert relerr(gg[0], oo[0]) < 1e-9 and relerr(gg[1], oo[1]) < 1e-9


def test_c4_pets_full_population_properties(gpu_be):
    """configs[3] in full (4096 sequences x 150 particles x T = 30, Philox noise): the cost of a sequence does not depend
    on which other sequences are in the launch, and is bit-reproducible -- checked on a 64-sequence slice whose particle
    streams are addressed by the same global indices"""
    prob, x0 = wl.c4_problem()
    spec = prob.spec()
    gen = prob.f_stochastic.gen()
    rng = np.random.Generator(np.random.Philox(key=7))
    ctrl = 2.0 * rng.standard_normal((1, 30, 4096))
    a = gpu_be.pets_costs(spec, x0, ctrl, 150, seed=11, gen=gen)
    b = gpu_be.pets_costs(spec, x0, ctrl, 150, seed=11, gen=gen)
    assert a.shape == (4096,) and np.all(np.isfinite(a)) and np.array_equal(a, b)
    c = gpu_be.pets_costs(spec, x0, ctrl[..., :64], 150, seed=11, gen=gen)  # sequences 0..63 keep their stream indices
    assert np.array_equal(a[:64], c)
    d = gpu_be.pets_costs(spec, x0, ctrl, 150, seed=12, gen=gen)
    assert not np.array_equal(a, d) and abs(np.mean(a) - np.mean(d)) < 0.05 * abs(np.mean(a))


def test_c5_fleet_64_problems_T50_ce_defaults(gpu_be, oracle_be):
    """configs[4] on a 64-problem subset: N = 50, CE defaults (cross_entropy...jl:109-114: 10 theta x 5 iterations + the
    final solve), injected normal streams; every problem against its own oracle_ce_solve run; 256-sample injected-noise
    Monte Carlo of the resulting policy for 4 of them"""
    P = 64
    prob, cps, x0, u = wl.fleet(P)
    assert prob.N == 50
    spec = prob.spec(cost_params=cps)
    z = np.random.Generator(np.random.Philox(key=65536)).standard_normal((P, 2000))
    g = gpu_be.ce_solve_fleet(spec, x0, u, 0.1, 1.0, 2.0, z_inject=z)  # num_samples 10, num_elite 3, iter_max 5, lambda .5
    f = oracle_be.raw.oracle_ce_solve
    f.restype = C.c_int32
    n, m, N = spec.n, spec.m, spec.N
    opts = make_opts()
    uflat = np.ascontiguousarray(u.ravel(order="F"))
    for p in range(P):
        sp = prob.spec(cost_params=cps[p])  # keep alive: desc() points into its arrays
        d = sp.desc()
        ce = OracleCEOpts(1.0, 2.0, 10, 3, 5, 0.5, 0)
        outs = [C.c_double() for _ in range(6)]
        nz, st = C.c_int64(), C.c_int32()
        x = np.zeros((n, N + 1), order="F"); l = np.zeros((m, N), order="F"); L = np.zeros((m, n, N), order="F")
        x0p = np.ascontiguousarray(x0[:, p])
        rc = f(C.byref(d), C.byref(opts), C.byref(ce), x0p.ctypes.data_as(dp), uflat.ctypes.data_as(dp), C.c_double(0.1),
               z[p].ctypes.data_as(dp), C.c_int64(z.shape[1]), *[C.byref(o) for o in outs], C.byref(nz),
               x.ctypes.data_as(dp), l.ctypes.data_as(dp), L.ctypes.data_as(dp), C.byref(st))
        assert rc == 0 and st.value == 0
        theta_opt, value, th_min, th_max, mu, sigma = [o.value for o in outs]
        assert g["nz_used"][p] == nz.value, p
        assert np.isclose(g["theta_opt"][p], theta_opt, rtol=1e-9) and np.isclose(g["value"][p], value, rtol=1e-9), p
        assert np.isclose(g["theta_min"][p], th_min, rtol=1e-12) and np.isclose(g["theta_max"][p], th_max, rtol=1e-12)
        assert np.isclose(g["mu"][p], mu, rtol=1e-9) and np.isclose(g["sigma"][p], sigma, rtol=1e-9)
        assert np.isclose(g["mu_init"][p], ce.mu_init) and np.isclose(g["sigma_init"][p], ce.sigma_init)
        assert relerr(g["x"][..., p], x) < 1e-9 and relerr(g["l"][..., p], l) < 1e-9 and relerr(g["L"][..., p], L) < 1e-9, p
        if p < 4:
            rng = np.random.Generator(np.random.Philox(key=256 + p))
            w = np.einsum("ij,jks->iks", np.linalg.cholesky(prob.W(0)), rng.standard_normal((4, 50, 256)))
            gm = gpu_be.mc_rollout(sp, x, l, L, 256, noise=w, theta_risk=1.0)
            om = oracle_be.mc_rollout(sp, x, l, L, 256, noise=w, theta_risk=1.0)
            assert relerr(gm["J"], om["J"]) < 1e-9 and relerr(gm["stats"], om["stats"]) < 1e-9
