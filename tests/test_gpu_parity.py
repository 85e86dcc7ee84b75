"""Parity proper: the CUDA library (through the C ABI) against the CPU oracle on identical seeded
inputs.  Tolerance: 1e-9 relative (BASELINE.json north_star) on values, trajectories and gains,
array-wise (max |a-b| / max |b|); the discrete path (status, iterations, line-search trials,
mu restarts) must be identical.  The same checks run on CPU against the g++ build of the kernel
arithmetic (tests/_hostemu) so that logic errors surface without a GPU."""
import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import workloads as wl

RTOL = 1e-9  # north_star: "within 1e-9 relative on gains, trajectories and costs"


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    assert np.array_equal(fin, np.isfinite(a)), "non-finite pattern differs"
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin])) / max(np.max(np.abs(b[fin])), 1e-300))


@pytest.fixture(params=["hostemu", "hostemu_dynamic", "hostemu_coop", pytest.param("gpu", marks=pytest.mark.gpu)])
def dut(request):
    return request.getfixturevalue(request.param + "_be")


def check_solve(dut, oracle_be, spec, x0, u, theta, P=None, opts=None):
    g = dut.ileqg_solve_batch(spec, x0, u, theta, eps_hist_cap=256, P=P, opts=opts)
    o = oracle_be.ileqg_solve_batch(spec, x0, u, theta, eps_hist_cap=256, P=P, opts=opts)
    for k in ("status", "iters", "trials", "restarts"):
        assert np.array_equal(g[k], o[k]), k
    ok = o["status"] == 0
    assert np.all(np.isinf(g["value"][~ok]))
    for k in ("value", "mu"):
        assert relerr(g[k][ok], o[k][ok]) < RTOL, k
    # d_current = max_k ||l_k - u_new,k|| (ileqg.jl:539) is a difference of controls: its scale is that of l
    l_scale = max(1.0, float(np.max(np.abs(o["l"][..., ok])))) if ok.any() else 1.0
    fin = np.isfinite(o["d_current"][ok])
    assert np.array_equal(fin, np.isfinite(g["d_current"][ok]))
    assert np.all(np.abs(g["d_current"][ok][fin] - o["d_current"][ok][fin]) < RTOL * l_scale), "d_current"
    for k in ("x", "l", "L", "eps_hist"):
        assert relerr(g[k][..., ok], o[k][..., ok]) < RTOL, k
    return g, o


def test_c1_shipped_problem(dut, oracle_be):
    prob, x0, u = wl.c1_problem()
    theta = [0.0, 0.1, 0.3, 0.43, 0.5, 5.0, 25.0, 30.7, 30.9, 40.0]  # includes the neurotic-breakdown boundary
    g, o = check_solve(dut, oracle_be, prob.spec(), x0, u, theta)
    assert list(o["status"][:8]) == [0] * 8 and o["status"][-1] == 1


def test_c2_unicycle_1024_thetas(dut, oracle_be):
    prob, x0, u = wl.c2_problem()
    g, o = check_solve(dut, oracle_be, prob.spec(), x0, u, wl.c2_thetas(1024))
    assert np.all(o["status"] == 0)


def test_fleet_per_problem_inputs(dut, oracle_be):
    prob, cps, x0, u = wl.fleet(6, N=20)
    spec = prob.spec(cost_params=cps)
    theta = wl.positive_thetas(6 * 4, key=11)
    check_solve(dut, oracle_be, spec, x0, u, theta, P=6)


@pytest.mark.parametrize("name", ["double_integrator", "pendulum", "cartpole", "single_integrator_docs"])
def test_other_models(dut, oracle_be, name):
    rng = np.random.default_rng(5)
    if name == "double_integrator":
        f, N = R.DoubleIntegrator(0.1), 30
        cost = R.QuadraticCost(4, 2, Q=0.1 * np.eye(4), R=0.05 * np.eye(2), Qf=np.eye(4), xg=[1.0, -1.0, 0, 0])
        W, x0 = 1e-3 * np.eye(4), np.zeros(4)
    elif name == "pendulum":
        f, N = R.Pendulum(), 40
        cost = R.QuadraticCost(2, 1, Q=np.diag([0.1, 0.01]), R=[[0.01]], Qf=np.diag([5.0, 0.5]), xg=[np.pi, 0.0])
        W, x0 = np.diag([1e-4, 1e-3]), np.array([0.1, 0.0])
    elif name == "cartpole":
        f, N = R.CartPole(), 30
        cost = R.QuadraticCost(4, 1, Q=np.diag([0.1, 1.0, 0.01, 0.01]), R=[[1e-2]], Qf=np.diag([1.0, 10.0, 0.1, 0.1]),
                               xg=[0.0, np.pi, 0.0, 0.0])
        W, x0 = np.diag([1e-6, 1e-6, 1e-4, 1e-4]), np.array([0.0, np.pi - 0.3, 0.0, 0.0])
    else:  # docs example: c = k/2 x'x + k/2 u'u, h = N/2 x'x, W = 0.1 I (optimal_control_problems.jl:48-64)
        f, N = R.SingleIntegrator(1.0), 10
        cost = R.QuadraticCost(2, 2, Q=np.eye(2), R=np.eye(2), Qf=N * np.eye(2), ws0=0.0, ws1=1.0)
        W, x0 = 0.1 * np.eye(2), np.array([1.0, -0.5])
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, R.ConstantCovariance(W), N)
    u = 0.01 * rng.standard_normal((f.m, N))
    tscale = 0.02 if name == "single_integrator_docs" else 0.2
    theta = np.concatenate([[0.0], np.abs(rng.standard_normal(15)) * tscale])
    g, o = check_solve(dut, oracle_be, prob.spec(), x0, u, theta)
    assert (o["status"] == 0).sum() >= 8  # mostly feasible, with some neurotic breakdowns mixed in


@pytest.mark.parametrize("seed", range(8))
def test_random_problems_sweep(dut, oracle_be, seed):
    """seeded random problems: random model of the registered set, dense SPD Q / R / Qf with a cross term Pc, stage-weight
    ramp, dense SPD (sometimes time-varying) W, random x0 / u_init / horizon, theta spanning feasible and infeasible"""
    rng = np.random.default_rng(1000 + seed)

    def spd(k, scale):
        a = rng.standard_normal((k, k))
        return scale * (a @ a.T / k + np.eye(k))

    f = [R.DoubleIntegrator(0.1), R.Pendulum(), R.Unicycle(0.1), R.CartPole(), R.SingleIntegrator(0.5)][seed % 5]
    n, m = f.n, f.m
    N = int(rng.integers(5, 26))
    xg = 0.5 * rng.standard_normal(n)
    cost = R.QuadraticCost(n, m, Q=spd(n, 0.05), R=spd(m, 0.05), Qf=spd(n, 0.5), xg=xg, Pc=0.01 * rng.standard_normal((n, m)),
                           ws0=1.0, ws1=float(rng.choice([0.0, 0.05])), c0=0.3, c1=0.01, h0=0.2)
    if seed % 2:
        Ws = np.stack([spd(n, 1e-3) * (1.0 + 0.05 * k) for k in range(N)])
        Wf = lambda k: Ws[k]
    else:
        Wf = R.ConstantCovariance(spd(n, 1e-3))
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, Wf, N)
    x0 = 0.3 * rng.standard_normal(n) + (np.array([0, 0, 0, 1.0])[:n] if isinstance(f, type(R.Unicycle(0.1))) and n == 4 else 0)
    u = 0.05 * rng.standard_normal((m, N))
    theta = np.concatenate([[0.0], np.abs(rng.standard_normal(11)) * 10.0 ** rng.uniform(-2, 1.5, 11)])
    g, o = check_solve(dut, oracle_be, prob.spec(), x0, u, theta, opts=R.make_opts(iter_max=30))
    assert (o["status"] == 0).any()


def test_quadrotor_small(dut, oracle_be):
    prob, x0, u = wl.c3_problem(N=10)
    check_solve(dut, oracle_be, prob.spec(), x0, u, [0.0, 0.05, 0.5], opts=R.make_opts(iter_max=6))


def test_time_varying_W_and_adaptive_eps(dut, oracle_be):
    prob, x0, u = wl.c2_problem(N=12)
    Ws = np.stack([np.diag([1e-3, 1e-3, 1e-4, 1e-3]) * (1.0 + 0.1 * k) for k in range(12)])
    prob.W = lambda k: Ws[k]
    check_solve(dut, oracle_be, prob.spec(), x0, u, wl.positive_thetas(8, key=3),
                opts=R.make_opts(adaptive_eps_init=True, eps_init=0.5, iter_max=25))


@pytest.mark.parametrize("nm", [(2, 2), (2, 1), (4, 2), (4, 1), (12, 4)])
@pytest.mark.parametrize("theta", [0.0, 0.3])
def test_riccati_passes_random_pd_inputs(dut, oracle_be, nm, theta):
    """both passes on random PD inputs at every registered (n, m)."""
    n, m = nm
    N, B = 7, 5
    rng = np.random.default_rng(100 * n + m)

    def spd(k, scale):
        a = rng.standard_normal((k, k))
        return scale * (a @ a.T / k + np.eye(k))

    lin = dict(q=rng.standard_normal((N + 1, B)), qv=rng.standard_normal((n, N + 1, B)),
               Q=np.stack([np.stack([spd(n, 1.0) for _ in range(N + 1)], -1) for _ in range(B)], -1),
               r=rng.standard_normal((m, N, B)),
               R=np.stack([np.stack([spd(m, 1.0) for _ in range(N)], -1) for _ in range(B)], -1),
               P=0.1 * rng.standard_normal((m, n, N, B)),
               A=np.eye(n)[:, :, None, None] + 0.1 * rng.standard_normal((n, n, N, B)),
               B=rng.standard_normal((n, m, N, B)))
    W = spd(n, 1e-3)
    th = np.full(B, theta)
    go, oo = dut.riccati(lin, W, th, True), oracle_be.riccati(lin, W, th, True)
    assert np.array_equal(go["status"], oo["status"]) and np.all(oo["status"] == 0)
    for k in ("s", "sv", "S", "L", "dl"):
        assert relerr(go[k], oo[k]) < RTOL, k
    ge = dut.riccati(lin, W, th, False, L=go["L"], dl=go["dl"])
    assert np.array_equal(ge["s"], go["s"])  # ileqg_test.jl:130: evaluation reproduces the optimising pass bitwise
    g0, o0 = dut.riccati(lin, W, th, False, L=oo["L"]), oracle_be.riccati(lin, W, th, False, L=oo["L"])
    for k in ("s", "sv", "S"):
        assert relerr(g0[k], o0[k]) < RTOL, k


def test_riccati_mu_restart_path(dut, oracle_be):
    """H not PD => increase mu, restart the sweep (ileqg.jl:372-378): indefinite R forces restarts."""
    n, m, N, B = 4, 2, 6, 3
    rng = np.random.default_rng(9)
    lin = dict(q=np.zeros((N + 1, B)), qv=rng.standard_normal((n, N + 1, B)),
               Q=np.tile(np.eye(n)[:, :, None, None], (1, 1, N + 1, B)), r=rng.standard_normal((m, N, B)),
               R=np.tile(np.diag([-1e-6, 1.0])[:, :, None, None], (1, 1, N, B)), P=np.zeros((m, n, N, B)),
               A=np.tile(np.eye(n)[:, :, None, None], (1, 1, N, B)), B=np.zeros((n, m, N, B)))
    lin["B"][1, 1] = 1.0  # u_0 does not act on the state: H_00 = R_00 + mu
    W = 1e-2 * np.eye(n)
    go = dut.riccati(lin, W, np.zeros(B), True, mu=0.0, delta=2.0)
    oo = oracle_be.riccati(lin, W, np.zeros(B), True, mu=0.0, delta=2.0)
    assert np.all(oo["restarts"] >= 2) and np.array_equal(go["restarts"], oo["restarts"])
    assert np.array_equal(go["mu"], oo["mu"]) and np.array_equal(go["delta"], oo["delta"])
    for k in ("s", "S", "L", "dl"):
        assert relerr(go[k], oo[k]) < RTOL, k


def test_linearize_and_rollouts(dut, oracle_be):
    for prob, x0, u in (wl.c1_problem(), wl.c2_problem(N=15), wl.c3_problem(N=5)):
        spec = prob.spec()
        rng = np.random.default_rng(1)
        B = 4
        us = np.abs(u[:, :, None] + 0.05 * rng.standard_normal(u.shape + (B,)))
        xg, sg = dut.rollout_open(spec, np.tile(x0[:, None], (1, B)), us)
        xo, so = oracle_be.rollout_open(spec, np.tile(x0[:, None], (1, B)), us)
        assert np.array_equal(sg, so) and relerr(xg, xo) < RTOL
        lg, lo = dut.linearize(spec, xo, us), oracle_be.linearize(spec, xo, us)
        for k in ("q", "qv", "Q", "r", "R", "P", "A", "B"):
            assert relerr(lg[k], lo[k]) < RTOL, k
        L = 0.01 * rng.standard_normal((spec.m, spec.n, spec.N, B))
        xbar = xo + 0.0
        xbar[:, 1:, :] += 1e-3 * np.abs(rng.standard_normal(xbar[:, 1:, :].shape))
        a, b = dut.rollout_closed(spec, xbar, us, L), oracle_be.rollout_closed(spec, xbar, us, L)
        assert np.array_equal(a[2], b[2])
        ok = b[2] == 0
        assert relerr(a[0][..., ok], b[0][..., ok]) < RTOL and relerr(a[1][..., ok], b[1][..., ok]) < RTOL
        cg, co = dut.integrate_cost(spec, xo, us), oracle_be.integrate_cost(spec, xo, us)
        assert relerr(cg[0], co[0]) < RTOL


def test_mc_rollout_injected_noise(dut, oracle_be):
    """C2's exactness check: closed-loop MC of 256 samples with an injected noise tensor w[256][50][4]."""
    prob, x0, u = wl.c2_problem()
    spec = prob.spec()
    sol = oracle_be.ileqg_solve_batch(spec, x0, u, [1.0])
    xbar, l, L = sol["x"][..., 0], sol["l"][..., 0], sol["L"][..., 0]
    rng = np.random.Generator(np.random.Philox(key=256))
    chol = np.linalg.cholesky(prob.W(0))
    w = np.einsum("ij,jks->iks", chol, rng.standard_normal((4, 50, 256)))
    g = dut.mc_rollout(spec, xbar, l, L, 256, noise=w, want_x=True)
    o = oracle_be.mc_rollout(spec, xbar, l, L, 256, noise=w, want_x=True)
    assert relerr(g["J"], o["J"]) < RTOL and relerr(g["x"], o["x"]) < RTOL


def test_pets_costs_injected_noise(dut, oracle_be):
    prob, x0 = wl.c4_problem(N=12)
    spec = prob.spec()
    rng = np.random.default_rng(3)
    C, Kp = 9, 10
    ctrl = rng.standard_normal((1, 12, C))
    noise = 1e-2 * rng.standard_normal((4, 12, Kp, C))
    gen = prob.f_stochastic.gen()
    g = dut.pets_costs(spec, x0, ctrl, Kp, noise=noise, gen=gen)
    o = oracle_be.pets_costs(spec, x0, ctrl, Kp, noise=noise, gen=gen)
    assert relerr(g, o) < RTOL
