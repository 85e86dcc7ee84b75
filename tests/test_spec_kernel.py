"""The speculative latency kernel (csrc/rl_spec.cuh: G lanes per instance evaluate several line-search candidates at
once and run the next iteration's optimising pass alongside each candidate's evaluating pass) must reproduce the
one-thread-per-instance kernel BIT FOR BIT: it evaluates the same per-instance arithmetic on the same inputs and only
replays the accept / reject rule of line_search! (ileqg.jl:504-592) in trial order.  CPU: the g++ builds of both state
machines; GPU: the library with RATILQR_SPEC forced on / off."""
import os

import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import workloads as wl


@pytest.fixture(params=[2, 4, 8])
def spec_be(request, hostemu_be):
    hostemu_be.dll.hostemu_set_spec(request.param)
    yield hostemu_be
    hostemu_be.dll.hostemu_set_spec(0)


def _same(a, b, cap):
    for k in ("status", "iters", "trials", "restarts"):
        assert np.array_equal(a[k], b[k]), k
    ok = b["status"] == 0
    for k in ("value", "mu", "d_current"):
        assert np.array_equal(a[k][ok], b[k][ok]), k
    assert np.all(np.isinf(a["value"][~ok]))
    for k in ("x", "l", "L"):
        assert np.array_equal(a[k][..., ok], b[k][..., ok]), k
    if cap:
        for i in np.nonzero(ok)[0]:
            nt = min(int(b["trials"][i]), cap)
            assert np.array_equal(a["eps_hist"][:, :nt, i], b["eps_hist"][:, :nt, i])


def _cases():
    out = []
    prob, x0, u = wl.c1_problem()
    out.append(("c1", prob.spec(), x0, u, np.array([0.0, 0.1, 0.3, 0.43, 0.5, 30.7, 31.0, 100.0]), None, None))
    out.append(("c1_domain", prob.spec(), x0, -0.1 * np.ones((2, 10)), np.array([0.0, 0.5]), None, None))   # negative base
    prob, x0, u = wl.c2_problem()
    out.append(("c2", prob.spec(), x0, u, wl.c2_thetas(24), None, None))
    out.append(("c2_adaptive", prob.spec(), x0, u, wl.c2_thetas(6), R.make_opts(adaptive_eps_init=True, iter_max=30), None))
    out.append(("c2_itermax", prob.spec(), x0, u, wl.c2_thetas(5), R.make_opts(iter_max=3), None))
    out.append(("c2_epsmin", prob.spec(), x0, u, wl.c2_thetas(5), R.make_opts(eps_init=1.0, eps_min=0.3, iter_max=20), None))
    prob, cps, x0s, u = wl.fleet(9, N=20)
    out.append(("fleet", prob.spec(cost_params=cps), x0s, u, wl.positive_thetas(9 * 5, key=11), None, 9))
    f = R.DoubleIntegrator(0.1)  # indefinite R: mu restarts inside the (speculative) optimising passes
    cost = R.QuadraticCost(4, 2, Q=0.5 * np.eye(4), R=np.diag([-0.02, 0.05]), Qf=np.eye(4), xg=[1.0, -1.0, 0, 0])
    p3 = R.FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, R.ConstantCovariance(1e-3 * np.eye(4)), 12)
    out.append(("restarts", p3.spec(), np.zeros(4), np.zeros((2, 12)), np.array([0.0, 0.05, 0.2]), R.make_opts(iter_max=7), None))
    prob, x0, _ = wl.c2_problem(N=1)
    out.append(("N1", prob.spec(), x0, np.zeros((2, 1)), wl.positive_thetas(3, key=1), R.make_opts(iter_max=12), None))
    pend = R.Pendulum()
    cost = R.QuadraticCost(2, 1, Q=np.diag([1.0, 0.1]), R=np.array([[0.01]]), Qf=10 * np.eye(2), xg=[np.pi, 0.0])
    p4 = R.FiniteHorizonRiskSensitiveOptimalControlProblem(pend, cost.c, cost.h, R.ConstantCovariance(1e-3 * np.eye(2)), 30)
    out.append(("pendulum", p4.spec(), np.zeros(2), np.zeros((1, 30)), np.array([0.0, 0.2, 1.0, 50.0]), R.make_opts(iter_max=25), None))
    cp = R.CartPole()
    cost = R.QuadraticCost(4, 1, Q=np.diag([0.1, 1.0, 0.01, 0.01]), R=np.array([[1e-3]]), Qf=10 * np.eye(4), xg=[0, np.pi, 0, 0],
                           Pc=0.01 * np.ones((4, 1)))  # non-diagonal cost: the dense QUADRATIC kernel
    p5 = R.FiniteHorizonRiskSensitiveOptimalControlProblem(cp, cost.c, cost.h, R.ConstantCovariance(1e-4 * np.eye(4)), 25)
    out.append(("cartpole_dense", p5.spec(), np.zeros(4), 0.1 * np.ones((1, 25)), np.array([0.0, 0.3]), R.make_opts(iter_max=15), None))
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_spec_state_machine_bitwise_equals_thread_per_instance(spec_be, hostemu_be, case):
    _, spec, x0, u, th, opts, P = case
    cap = 64
    a = spec_be.ileqg_solve_batch(spec, x0, u, th, opts=opts, eps_hist_cap=cap, P=P)
    spec_be.dll.hostemu_set_spec(0)
    b = hostemu_be.ileqg_solve_batch(spec, x0, u, th, opts=opts, eps_hist_cap=cap, P=P)
    assert np.any(b["status"] == 0) or case[0] == "c1_domain"
    _same(a, b, cap)


def test_spec_case_list_exercises_the_interesting_paths(hostemu_be):
    """the case list above contains rejected trials, forced accepts at eps_min, mu restarts, failures in initialize! and
    DomainErrors -- otherwise the bitwise test would be vacuous"""
    seen = dict(rejected=False, restarts=False, init_fail=False, domain=False, many_trials=False)
    for name, spec, x0, u, th, opts, P in _cases():
        r = hostemu_be.ileqg_solve_batch(spec, x0, u, th, opts=opts, P=P)
        ok = r["status"] == 0
        seen["rejected"] |= bool(np.any(r["trials"][ok] > r["iters"][ok]))
        seen["many_trials"] |= bool(np.any(r["trials"][ok] > 2 * r["iters"][ok]))
        seen["restarts"] |= bool(np.any(r["restarts"] > 0))
        seen["init_fail"] |= bool(np.any(r["status"] == 1))
        seen["domain"] |= bool(np.any(r["status"] == 3))
    assert all(seen.values()), seen


@pytest.mark.gpu
@pytest.mark.parametrize("G", [2, 4, 8])
def test_spec_kernel_gpu_bitwise_equals_thread_per_instance(G):
    """the CUDA library with the speculative kernel forced on (G lanes per instance) vs forced off, fresh contexts"""
    res = {}
    for mode in ("0", str(G)):
        os.environ["RATILQR_SPEC"] = mode
        try:
            be = R.new_backend(0)
            res[mode] = [be.ileqg_solve_batch(spec, x0, u, th, opts=opts, eps_hist_cap=64, P=P) for _, spec, x0, u, th, opts, P in _cases()]
            be.close()
        finally:
            del os.environ["RATILQR_SPEC"]
    for a, b in zip(res[str(G)], res["0"]):
        _same(a, b, 64)


@pytest.mark.gpu
def test_spec_kernel_is_the_default_for_small_batches_and_matches_the_oracle(gpu_be, oracle_be):
    """configs[1] exactly (1 problem x 1024 theta) goes through the speculative kernel by default: identical discrete paths
    and <= 1e-9 against the oracle; the CE fleet loop on a handful of problems uses it too (covered by the fleet tests)"""
    prob, x0, u = wl.c2_problem()
    th = wl.c2_thetas(1024)
    n0 = gpu_be.launch_count()
    g = gpu_be.ileqg_solve_batch(prob.spec(), x0, u, th)
    assert gpu_be.launch_count() - n0 == 2  # k_sort_theta + k_ileqg_solve_spec: no gather kernels (outputs written in host layout)
    o = oracle_be.ileqg_solve_batch(prob.spec(), x0, u, th)
    for k in ("status", "iters", "trials", "restarts"):
        assert np.array_equal(g[k], o[k]), k
    ok = o["status"] == 0
    assert ok.sum() > 900
    assert np.max(np.abs(g["value"][ok] - o["value"][ok]) / np.abs(o["value"][ok])) < 1e-9
    for k in ("x", "l", "L"):
        assert np.max(np.abs(g[k][..., ok] - o[k][..., ok])) / np.max(np.abs(o[k][..., ok])) < 1e-9, k


# ---- the two-warp speculative variant of the warp-cooperative kernel (csrc/rl_coop2.cuh) ---------------------------------
def _coop_cases():
    out = [c for c in _cases() if c[0] in ("c1", "c1_domain", "c2_epsmin", "restarts", "cartpole_dense")]
    prob, x0, u = wl.c3_problem()
    out.append(("c3_quadrotor", prob.spec(), x0, u, np.array([0.0, 0.005, 0.01, 0.02, 0.0234375, 0.05, 3.0]), None, None))
    prob, x0, u = wl.c3_problem(N=7)
    out.append(("c3_short_itermax", prob.spec(), x0, u, np.array([0.0, 0.01]), R.make_opts(iter_max=4, adaptive_eps_init=True), None))
    return out


@pytest.mark.parametrize("case", _coop_cases(), ids=lambda c: c[0])
def test_coop2_bitwise_equals_one_warp_coop(hostemu_be, case):
    _, spec, x0, u, th, opts, P = case
    try:
        hostemu_be.dll.hostemu_set_coop(2)
        a = hostemu_be.ileqg_solve_batch(spec, x0, u, th, opts=opts, eps_hist_cap=64, P=P)
        hostemu_be.dll.hostemu_set_coop(1)
        b = hostemu_be.ileqg_solve_batch(spec, x0, u, th, opts=opts, eps_hist_cap=64, P=P)
    finally:
        hostemu_be.dll.hostemu_set_coop(0)
    _same(a, b, 64)


@pytest.mark.gpu
def test_coop2_kernel_gpu_bitwise_equals_one_warp_coop(oracle_be):
    """the CUDA library: quadrotor solves through the two-warp kernel (default for small batches) vs the one-warp kernel,
    and both against the oracle"""
    prob, x0, u = wl.c3_problem()
    th = np.array([0.0, 0.005, 0.01, 0.02, 0.0234375, 0.05, 3.0])
    res = {}
    for mode in ("0", "1"):
        os.environ["RATILQR_COOP2"] = mode
        try:
            be = R.new_backend(0)
            n0 = be.launch_count()
            res[mode] = be.ileqg_solve_batch(prob.spec(), x0, u, th, eps_hist_cap=64)
            assert be.launch_count() - n0 == 2   # k_sort_theta + one solve kernel (outputs written in host layout, no gather)
            be.close()
        finally:
            del os.environ["RATILQR_COOP2"]
    _same(res["1"], res["0"], 64)
    o = oracle_be.ileqg_solve_batch(prob.spec(), x0, u, th)
    assert np.array_equal(res["1"]["status"], o["status"]) and np.array_equal(res["1"]["iters"], o["iters"])
    ok = o["status"] == 0
    assert np.max(np.abs(res["1"]["value"][ok] - o["value"][ok]) / np.abs(o["value"][ok])) < 1e-9
    assert np.max(np.abs(res["1"]["L"][..., ok] - o["L"][..., ok])) / np.max(np.abs(o["L"][..., ok])) < 1e-9


def _dense_cases():
    prob, x0, u = wl.c3_problem()
    out = [("c3_quadrotor_diag_cost", prob.spec(), x0, u, np.array([0.0, 0.005, 0.01, 0.02, 0.0234375, 0.05, 3.0]), None)]
    Q = np.diag([1.0] * 3 + [0.1] * 3 + [0.1] * 3 + [0.01] * 3)
    Q[0, 1] = Q[1, 0] = 0.05
    cost = R.QuadraticCost(12, 4, Q=Q, R=np.diag([0.01, 1.0, 1.0, 1.0]), Qf=10 * Q, xg=np.r_[1.0, 1.0, 1.0, np.zeros(9)], Pc=1e-3 * np.ones((12, 4)))
    p2 = R.FiniteHorizonRiskSensitiveOptimalControlProblem(R.Quadrotor(0.05), cost.c, cost.h, R.ConstantCovariance(np.asarray(prob.W(0))), 20)
    u2 = np.zeros((4, 20))
    u2[0] = 9.81
    out.append(("quadrotor_dense_cost", p2.spec(), np.zeros(12), u2, np.array([0.005, 0.01, 0.02]), R.make_opts(iter_max=8)))
    cost = R.QuadraticCost(4, 1, Q=np.diag([0.1, 1.0, 0.01, 0.01]), R=np.array([[1e-3]]), Qf=10 * np.eye(4), xg=[0, np.pi, 0, 0], Pc=0.01 * np.ones((4, 1)))
    p5 = R.FiniteHorizonRiskSensitiveOptimalControlProblem(R.CartPole(), cost.c, cost.h, R.ConstantCovariance(1e-4 * np.eye(4)), 25)
    out.append(("cartpole", p5.spec(), np.zeros(4), 0.1 * np.ones((1, 25)), np.array([0.1, 0.3, 2.0]), R.make_opts(iter_max=15)))
    return out


@pytest.mark.parametrize("case", _dense_cases(), ids=lambda c: c[0])
def test_coop_dense_stage_bitwise_equals_rolled_stage(hostemu_be, case):
    """rl::coop_riccati_stage_dense (the lane's dot products advance together, right-looking substitution, full D S+) must
    reproduce the plain element-by-element stage bit for bit: same operations per output element, only their interleaving
    differs.  Reference build: the same emulation compiled with -DRL_COOP_DENSE=0."""
    from tests import _hostemu
    rolled = _hostemu.load_variant("hostemu_coop_rolled")
    _, spec, x0, u, th, opts = case
    try:
        for be in (hostemu_be, rolled):
            be.dll.hostemu_set_coop(1)
        a = hostemu_be.ileqg_solve_batch(spec, x0, u, th, opts=opts, eps_hist_cap=64)
        b = rolled.ileqg_solve_batch(spec, x0, u, th, opts=opts, eps_hist_cap=64)
    finally:
        for be in (hostemu_be, rolled):
            be.dll.hostemu_set_coop(0)
    assert np.any(b["status"] == 0)
    _same(a, b, 64)
