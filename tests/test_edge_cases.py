"""Edge cases the reference's own suite does not exercise ("coverage gaps"), checked against the
oracle: mu restarts inside a full solve (and the reference's consequence: after two increases mu > mu_min, the
convergence test ileqg.jl:642 can never pass and the solve runs to iter_max), N = 1, long horizons, odd batch sizes,
NaN inputs, and argument validation through the C ABI."""
import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import workloads as wl
from tests.test_gpu_parity import check_solve, relerr, RTOL  # noqa: F401


@pytest.fixture(params=["hostemu", "hostemu_coop", pytest.param("gpu", marks=pytest.mark.gpu)])
def dut(request):
    return request.getfixturevalue(request.param + "_be")


def test_mu_restarts_inside_full_solve(dut, oracle_be):
    """indefinite R: H = R + B'DSB is not PD at the late stages => increase_mu_and_delta! restarts (ileqg.jl:372-378)"""
    f = R.DoubleIntegrator(0.1)
    cost = R.QuadraticCost(4, 2, Q=0.5 * np.eye(4), R=np.diag([-0.02, 0.05]), Qf=np.eye(4), xg=[1.0, -1.0, 0, 0])
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, R.ConstantCovariance(1e-3 * np.eye(4)), 12)
    u = np.zeros((2, 12))
    g, o = check_solve(dut, oracle_be, prob.spec(), np.zeros(4), u, [0.0, 0.05, 0.2], opts=R.make_opts(iter_max=7))
    assert np.all(o["restarts"] >= 2) and np.all(o["status"] == 0)
    # mu persists across iterations and is never decreased: once it exceeds mu_min the convergence test (:642) cannot
    # pass any more and the solve can only stop at iter_max
    assert np.all(o["mu"] > 1e-6) and np.all(o["iters"] == 7)


@pytest.mark.parametrize("N", [1, 2, 120])
def test_horizon_extremes(dut, oracle_be, N):
    prob, x0, _ = wl.c2_problem(N=N)
    check_solve(dut, oracle_be, prob.spec(), x0, np.zeros((2, N)), wl.positive_thetas(5, key=N), opts=R.make_opts(iter_max=12))


@pytest.mark.parametrize("B", [1, 31, 33, 65])
def test_odd_batch_sizes(dut, oracle_be, B):
    prob, x0, u = wl.c2_problem(N=10)
    g, o = check_solve(dut, oracle_be, prob.spec(), x0, u, wl.positive_thetas(B, key=B), opts=R.make_opts(iter_max=6))
    assert g["value"].shape == (B,)


def test_nan_inputs_do_not_hang(dut, oracle_be):
    """NaN in x0: Cholesky pivots are NaN => isposdef false => neurotic-breakdown status, like the reference's assert"""
    prob, x0, u = wl.c2_problem(N=8)
    x0 = x0.copy()
    x0[1] = np.nan
    g = dut.ileqg_solve_batch(prob.spec(), x0, u, [0.0, 0.5])
    o = oracle_be.ileqg_solve_batch(prob.spec(), x0, u, [0.0, 0.5])
    assert np.array_equal(g["status"], o["status"]) and np.all(np.isinf(g["value"]) | np.isnan(g["value"]))


def test_iter_max_one_and_eps_history(dut, oracle_be):
    prob, x0, u = wl.c1_problem()
    g, o = check_solve(dut, oracle_be, prob.spec(), x0, u, [0.0, 0.5], opts=R.make_opts(iter_max=1))
    assert np.all(g["iters"] == 1) and np.all(g["trials"] == 1)
    assert g["eps_hist"][0, 0, 0] == 1.0 and g["eps_hist"][1, 0, 0] < 0.0  # (eps, new - cur) of the only trial


@pytest.mark.gpu
def test_argument_validation_through_the_c_abi(gpu_be):
    prob, x0, u = wl.c2_problem(N=5)
    spec = prob.spec()
    with pytest.raises(R.ApiError, match="lambda"):       # @assert 0 < λ < 1  (ileqg.jl:195)
        gpu_be.ileqg_solve_batch(spec, x0, u, [0.1], opts=R.make_opts(lam=1.5))
    with pytest.raises(R.ApiError, match="eps_init"):     # @assert ϵ_init > ϵ_min  (:200)
        gpu_be.ileqg_solve_batch(spec, x0, u, [0.1], opts=R.make_opts(eps_init=1e-7))
    bad = prob.spec()
    bad.W = np.ascontiguousarray((-np.eye(4)).ravel())
    with pytest.raises(R.ApiError, match="positive definite"):
        gpu_be.ileqg_solve_batch(bad, x0, u, [0.1])
    bad = prob.spec()
    bad.model_id = 99
    with pytest.raises(R.ApiError, match="model_id"):
        gpu_be.ileqg_solve_batch(bad, x0, u, [0.1])
    l1 = R.L1ControlCost()
    p2 = R.FiniteHorizonRiskSensitiveOptimalControlProblem(R.SingleIntegrator(), l1.c, l1.h, R.ConstantCovariance(np.eye(2)), 4)
    with pytest.raises(R.ApiError, match="rollout-only"):
        gpu_be.ileqg_solve_batch(p2.spec(), np.zeros(2), np.zeros((2, 4)), [0.0])
    # the context stays usable after errors
    r = gpu_be.ileqg_solve_batch(spec, x0, u, [0.1])
    assert r["status"][0] == 0


def test_unregistered_closures_are_rejected():
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(lambda x, u: x + u, lambda k, x, u: 0.0, lambda x: 0.0,
                                                             lambda k: np.eye(2), 3)
    with pytest.raises(TypeError):
        prob.spec()  # arbitrary closures are not accelerated and there is no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("spec_mode", ["0", "8"])
def test_heading_beyond_the_fast_path_of_sincos(gpu_be, oracle_be, spec_mode):
    """|psi| >= 2^31: the branch-free sincos of the solve kernels (rl_sincos_nb) only RECORDS that the library routine's
    Payne-Hanek slow path is needed; the stage (backward passes) / the step (rollout) is then redone through the library
    routine.  Thread-per-instance kernel (RATILQR_SPEC=0) and speculative latency kernel against the oracle."""
    import os
    N = 20
    cost = wl.unicycle_cost(goal=(5.0, 5.0, 3.0e9, 0.0))
    W = np.diag([1e-2, 1e-2, 1e-3, 1e-2]) * 0.1
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(R.Unicycle(0.1), cost.c, cost.h, R.ConstantCovariance(W), N)
    x0 = np.array([0.0, 0.0, 3.0e9 + 0.3, 1.0])
    os.environ["RATILQR_SPEC"] = spec_mode
    try:
        be = R.new_backend(0)
        g, o = check_solve(be, oracle_be, prob.spec(), x0, np.zeros((2, N)), [0.0, 0.3, 1.0, 2.5], opts=R.make_opts(iter_max=15))
        be.close()
    finally:
        del os.environ["RATILQR_SPEC"]
    assert np.all(o["status"] == 0) and np.all(o["iters"] >= 3)
