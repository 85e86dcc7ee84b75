"""Pins every provider of the C ABI (CPU oracle, g++ build of the kernel arithmetic, CUDA library) to an INDEPENDENT
literal restatement of src/ileqg.jl: tests/ref_literal/ileqg_literal.py follows the Julia file operation by operation
(inv(W) - θS, D = I + (θS)/M by a general solve, (-H)\\G by LU, logdet(W*M), forward-mode AD for the derivatives) in
numpy float64 and in 60-digit mpmath, and shares no code with oracle/ or csrc/.  The fixtures tests/golden/literal_*.npz
hold the 60-digit results (rounded to float64); tests/golden/make_literal_goldens.py regenerates them.

Tolerances: 1e-12 array-relative for the CPU oracle (1e-11 within 3 % of the neurotic-breakdown threshold, where
cond(M) amplifies rounding), 1e-9 for the CUDA library (the north-star tolerance), identical discrete paths
(status, iterations, line-search trials, mu restarts, eps sequence) everywhere."""
import os

import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import workloads as wl

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tol(backend_name, near=False):
    if backend_name == "gpu":
        return 1e-9
    return 1e-11 if near else 1e-12


def _relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


@pytest.fixture(params=["oracle", "hostemu", "hostemu_coop", pytest.param("gpu", marks=pytest.mark.gpu)])
def named_backend(request):
    """(name, provider): the CPU oracle, the g++ builds of the thread-per-instance and of the warp-cooperative kernel
    arithmetic, and the CUDA library"""
    return request.param, request.getfixturevalue(request.param + "_be")


def _riccati_cases():
    z = np.load(os.path.join(GOLD, "literal_riccati.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


def test_literal_f64_agrees_with_mp():
    """the fixture itself: the float64 statement of the reference's formulas is within rounding of the 60-digit one"""
    z, names = _riccati_cases()
    for name in names:
        for tag in ("opt", "eval", "eval_nodl"):
            for k in ("s", "sv", "S", "L"):
                assert _relerr(z[f"{name}/f64/{tag}/{k}"], z[f"{name}/mp/{tag}/{k}"]) < 1e-11, (name, tag, k)
            assert int(z[f"{name}/f64/{tag}/restarts"]) == int(z[f"{name}/mp/{tag}/restarts"])


def test_riccati_passes_match_literal(named_backend):
    """solve_approximate_dp! / solve_approximate_dp (ileqg.jl:341-465) on dense random SPD stage data at all five (n, m),
    theta = 0, mid-range and 97 % of the breakdown threshold, and an indefinite-R case that exercises the mu restarts"""
    bname, be = named_backend
    z, names = _riccati_cases()
    assert len(names) == 20
    for name in names:
        g = lambda k: z[f"{name}/in/{k}"]  # noqa: E731
        lin = {k: g(k)[..., None] for k in ("q", "qv", "Q", "r", "R", "P", "A", "B")}
        theta = float(g("theta"))
        tol = _tol(bname, near=name.endswith("near"))
        # optimising pass
        ref = {k: z[f"{name}/mp/opt/{k}"] for k in ("s", "sv", "S", "L", "dl", "mu", "delta", "restarts")}
        r = be.riccati(lin, g("W"), theta, True)
        assert r["status"][0] == 0, name
        assert int(r["restarts"][0]) == int(ref["restarts"]), name
        assert float(r["mu"][0]) == float(ref["mu"]) and float(r["delta"][0]) == float(ref["delta"]), name
        for k in ("s", "sv", "S", "L", "dl"):
            assert _relerr(r[k][..., 0], ref[k]) < tol, (name, "opt", k, _relerr(r[k][..., 0], ref[k]))
        # evaluating pass with a given policy (L, dl) and mu = 0.5
        ref = {k: z[f"{name}/mp/eval/{k}"] for k in ("s", "sv", "S")}
        r = be.riccati(lin, g("W"), theta, False, L=g("L_eval"), dl=g("dl_eval"), mu=0.5)
        assert r["status"][0] == 0, name
        for k in ("s", "sv", "S"):
            assert _relerr(r[k][..., 0], ref[k]) < tol, (name, "eval", k, _relerr(r[k][..., 0], ref[k]))
        # evaluating pass with dl = nothing (the line-search merit, :523-525)
        ref = {k: z[f"{name}/mp/eval_nodl/{k}"] for k in ("s", "sv", "S")}
        r = be.riccati(lin, g("W"), theta, False, L=g("L_eval"), dl=None, mu=0.0)
        assert r["status"][0] == 0, name
        for k in ("s", "sv", "S"):
            assert _relerr(r[k][..., 0], ref[k]) < tol, (name, "eval_nodl", k, _relerr(r[k][..., 0], ref[k]))


def test_riccati_breakdown_threshold_matches_literal(named_backend):
    """3 % beyond the literal implementation's breakdown threshold every provider reports M not PD"""
    _, be = named_backend
    z, names = _riccati_cases()
    for name in names:
        if not name.endswith("near"):
            continue
        g = lambda k: z[f"{name}/in/{k}"]  # noqa: E731
        lin = {k: g(k)[..., None] for k in ("q", "qv", "Q", "r", "R", "P", "A", "B")}
        tstar = float(g("theta")) / 0.97
        ok = [be.riccati(lin, g("W"), t, True)["status"][0] == 0 and
              be.riccati(lin, g("W"), t, False, L=g("L_eval"), dl=g("dl_eval"), mu=0.5)["status"][0] == 0 and
              be.riccati(lin, g("W"), t, False, L=g("L_eval"), dl=None, mu=0.0)["status"][0] == 0
              for t in (0.999 * tstar, 1.001 * tstar)]
        assert ok == [True, False], (name, ok)


def _problem(tag):
    if tag == "c1":
        return wl.c1_problem()
    if tag == "c2":
        return wl.c2_problem()
    return wl.c3_problem()


@pytest.mark.parametrize("tag", ["c1", "c2", "c3"])
def test_whole_solves_match_literal(named_backend, tag):
    """solve!(::ILEQGSolver) (ileqg.jl:635-659) against the 60-digit literal run: C1 (the reference's shipped problem,
    incl. theta next to / beyond the feasibility boundary), a 10-theta subset of configs[1] (unicycle, N = 50) and
    configs[2]'s quadrotor at N = 40 for three theta"""
    bname, be = named_backend
    z = np.load(os.path.join(GOLD, f"literal_{tag}.npz"))
    prob, x0, u = _problem(tag)
    assert np.array_equal(z["x0"], x0) and np.array_equal(z["u_init"], u)
    thetas = z["thetas"]
    cap = 512
    r = be.ileqg_solve_batch(prob.spec(), x0, u, thetas, eps_hist_cap=cap)
    tol = _tol(bname)
    for i, th in enumerate(thetas):
        g = lambda k: z[f"mp/{i}/{k}"]  # noqa: E731
        assert int(r["status"][i]) == int(g("status")), (tag, th)
        if int(g("status")) != 0:
            assert np.isinf(r["value"][i])
            continue
        assert int(r["iters"][i]) == int(g("iters")), (tag, th, r["iters"][i], g("iters"))
        assert int(r["trials"][i]) == int(g("trials")), (tag, th)
        assert int(r["restarts"][i]) == int(g("restarts")), (tag, th)
        assert float(r["mu"][i]) == float(g("mu"))
        assert abs(r["value"][i] - float(g("value"))) <= tol * abs(float(g("value"))), (tag, th)
        for k in ("x", "l", "L"):
            e = _relerr(r[k][..., i], g(k))
            assert e < tol, (tag, th, k, e)
        assert abs(r["d_current"][i] - float(g("d_current"))) <= max(1e3 * tol * abs(float(g("d_current"))), 1e-13)
        nt = int(g("trials"))
        assert nt <= cap
        eh = g("eps_hist")
        assert np.array_equal(r["eps_hist"][0, :nt, i], eh[:, 0]), (tag, th)  # the eps sequence: exact
        assert np.allclose(r["eps_hist"][1, :nt, i], eh[:, 1], rtol=0, atol=1e3 * tol * abs(float(g("value")))), (tag, th)
        # the float64 statement of the reference walks the same discrete path
        assert int(z[f"f64/{i}/iters"]) == int(g("iters")) and int(z[f"f64/{i}/trials"]) == int(g("trials"))


# ---- live comparison (no fixtures): the float64 literal statement of ileqg.jl vs every CPU provider on random problems ----
_LIT_MODELS = {1: "single_integrator", 2: "power_law", 3: "double_integrator", 4: "pendulum", 5: "cartpole", 6: "unicycle"}


def _literal_solve(spec, x0, u, theta, opts):
    from tests.ref_literal import ileqg_literal as lit
    LA = lit.F64
    n, m, N = spec.n, spec.m, spec.N
    f = lit.make_dynamics(LA, _LIT_MODELS[spec.model_id], list(spec.model_params))
    c, h = lit.make_quadratic_cost(LA, n, m, list(spec.cost_params))
    W = LA.arr(np.asarray(spec.W).reshape(n, n, order="F"))
    prob = lit.Problem(f, c, h, lambda k: W, N, n, m)
    sol = lit.Solver(LA, **opts)
    try:
        x, l, L, v, eh = lit.solve(LA, sol, prob, x0, [u[:, k] for k in range(N)], theta)
    except lit.NotPosDef:
        return None
    return dict(x=np.stack(x, -1), l=np.stack(l, -1), L=np.stack(L, -1), value=float(v), iters=sol.iter_current, trials=len(eh),
                restarts=sol.restarts)


def _random_problem(rng, which):
    def spd(k, scale):
        G = rng.standard_normal((k, k))
        return scale * (G @ G.T / k + 0.3 * np.eye(k))
    if which == "single":
        f, n, m, N = R.SingleIntegrator(0.5), 2, 2, 8
    elif which == "double":
        f, n, m, N = R.DoubleIntegrator(0.2), 4, 2, 10
    elif which == "pendulum":
        f, n, m, N = R.Pendulum(), 2, 1, 12
    elif which == "cartpole":
        f, n, m, N = R.CartPole(), 4, 1, 10
    else:
        f, n, m, N = R.Unicycle(0.1), 4, 2, 10
    # dense Q, R, Qf, a cross term Pc and a time-scaled stage weight: everything QuadraticCost can express
    cost = R.QuadraticCost(n, m, Q=spd(n, 0.5), R=spd(m, 0.3), Qf=spd(n, 2.0), xg=rng.standard_normal(n), Pc=0.05 * rng.standard_normal((n, m)),
                           ws0=1.0, ws1=float(rng.uniform(0.0, 0.2)), c0=0.1, c1=0.01, h0=0.5)
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, R.ConstantCovariance(spd(n, 0.02)), N)
    return prob, 0.3 * rng.standard_normal(n), 0.1 * rng.standard_normal((m, N))


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["single", "double", "pendulum", "cartpole", "unicycle"])
def test_random_problems_match_literal_float64_gpu(which, gpu_be):
    """the same live comparison for the CUDA library (speculative latency kernel: these are small batches)"""
    test_random_problems_match_literal_float64(which, gpu_be, gpu_be)


@pytest.mark.parametrize("which", ["single", "double", "pendulum", "cartpole", "unicycle"])
def test_random_problems_match_literal_float64(which, oracle_be, hostemu_be):
    """dense random quadratic costs (non-diagonal Q, R, cross term, time-scaled weight), dense random W, five registered
    models, theta from 0 to beyond breakdown: oracle and kernel arithmetic (g++ build) against the literal float64
    statement of ileqg.jl evaluated live -- same discrete path, 1e-9 on values / trajectories / gains"""
    rng = np.random.Generator(np.random.Philox(key=77 + ["single", "double", "pendulum", "cartpole", "unicycle"].index(which)))
    opts = dict(iter_max=12)
    checked = feasible = 0
    for _ in range(4):
        prob, x0, u = _random_problem(rng, which)
        spec = prob.spec()
        thetas = np.array([0.0, 0.02, 0.15, 1.0, 40.0])
        res = [be.ileqg_solve_batch(spec, x0, u, thetas, opts=R.make_opts(**opts)) for be in (oracle_be, hostemu_be)]
        for i, th in enumerate(thetas):
            ref = _literal_solve(spec, x0, u, float(th), opts)
            for r in res:
                if ref is None:
                    assert r["status"][i] in (1, 2), (which, th, r["status"][i])
                    continue
                assert r["status"][i] == 0, (which, th)
                assert r["iters"][i] == ref["iters"] and r["trials"][i] == ref["trials"] and r["restarts"][i] == ref["restarts"], (which, th)
                assert abs(r["value"][i] - ref["value"]) <= 1e-9 * abs(ref["value"]) + 1e-12
                for k in ("x", "l", "L"):
                    assert _relerr(r[k][..., i], ref[k]) < 1e-9, (which, th, k)
            checked += 1
            feasible += ref is not None
    assert checked == 20 and feasible >= 5
