"""test/ileqg_test.jl of the reference, re-expressed line by line against the host mirror.
Runs against the CPU oracle and the g++ build of the kernel arithmetic here (-m "not gpu") and
against the CUDA library on the B200 (-m gpu).  Reference line numbers in comments."""
import math

import numpy as np

import ratilqr_b200 as R
from ratilqr_b200 import ileqg as IL


def _prob_linear(cost, N=10):
    return R.FiniteHorizonRiskSensitiveOptimalControlProblem(R.SingleIntegrator(1.0), cost.c, cost.h,
                                                             R.ConstantCovariance(np.eye(2)), N)


def test_ileqg_reference_suite(backend):
    be = backend
    N = 10
    cost = R.QuadraticCost(2, 2, c1=1.0, h0=1.0)  # c(k,x,u) = k ; h(x) = 1.0   (:13-14)
    prob = _prob_linear(cost, N)
    f = prob.f
    u_array = [np.ones(2) for _ in range(N)]
    x_array = R.simulate_dynamics(prob, np.zeros(2), u_array, backend=be)
    assert np.array_equal(x_array[0], np.zeros(2))                                              # :21
    assert all(np.array_equal(x_array[i + 1], f(x_array[i], u_array[i])) for i in range(N))     # :22

    L_array = [np.ones((2, 2)) for _ in range(N)]
    x_new, u_new = R.simulate_dynamics(prob, x_array, u_array, L_array, backend=be)
    assert all(np.array_equal(u_new[i], u_array[i]) for i in range(N))                          # :26
    assert all(np.array_equal(x_new[i], x_array[i]) for i in range(N))                          # :27

    c = R.integrate_cost(prob, x_array, u_array, backend=be)                                    # :30-31
    assert np.isclose(c, sum(prob.c(i, x_array[i], u_array[i]) for i in range(N)) + prob.h(x_array[-1]))

    solver = R.ILEQGSolver(prob, backend=be)
    IL.initialize_(solver, prob, np.zeros(2), u_array, 0.0)
    assert all(np.array_equal(a, b) for a, b in zip(solver.l_array, u_array))                   # :36
    assert all(np.array_equal(L, np.zeros((2, 2))) for L in solver.L_array)                     # :37
    assert all(np.array_equal(a, b) for a, b in zip(solver.x_array, x_array))                   # :38
    assert solver.mu == 0.0 and solver.delta == solver.delta_0                                  # :39-40
    assert solver.d_current == math.inf and solver.iter_current == 0 and solver.eps_history == []  # :41-43
    dp_init = R.solve_approximate_dp(R.approximate_model(prob, u_array, x_array, backend=be),
                                     [np.zeros((2, 2)) for _ in range(N)], theta=0.0, mu=0.0, backend=be)
    assert np.isclose(solver.value_current, dp_init.s_array[0])                                 # :47

    # approximate_model test (:50-63): c = 0.5 x'x + u'u + x'u ; h = 0.5 x'x
    cost2 = R.QuadraticCost(2, 2, Q=np.eye(2), R=2 * np.eye(2), Pc=np.eye(2), Qf=np.eye(2))
    prob.c, prob.h = cost2.c, cost2.h
    ap = R.approximate_model(prob, u_array, x_array, backend=be)
    for i in range(N):
        assert np.isclose(ap.q_array[i], 0.5 * (2 * i ** 2) + 1.0 * 2 + 2 * i)                 # :54
        assert np.allclose(ap.q_vec_array[i], x_array[i] + np.ones(2))                          # :57
        assert np.allclose(ap.r_array[i], x_array[i] + 2.0 * np.ones(2))                        # :60
        assert np.allclose(ap.R_array[i], 2.0 * np.eye(2))                                      # :61
        assert np.allclose(ap.P_array[i], np.eye(2))                                            # :62
        assert np.array_equal(ap.W_array[i], np.eye(2))                                         # :63
    assert np.isclose(ap.q_array[-1], prob.h(x_array[-1]))                                      # :55
    assert np.allclose(ap.q_vec_array[-1], x_array[-1])                                         # :58
    assert all(np.allclose(Q, np.eye(2)) for Q in ap.Q_array)                                   # :59

    cost3 = R.QuadraticCost(2, 2, Q=np.eye(2), R=2 * np.eye(2), Qf=np.eye(2))                   # :65-66
    prob.c, prob.h = cost3.c, cost3.h
    ap = R.approximate_model(prob, u_array, x_array, backend=be)
    dp, dl_new = R.solve_approximate_dp_(solver, ap, False, theta=0.0)
    assert len(dp.s_array) == len(x_array) == len(dp.s_vec_array) == len(dp.S_array)            # :70-73
    assert all(np.array_equal(S, S.T) for S in dp.S_array)                                      # :75
    assert all(np.all(np.linalg.eigvalsh(S) > 0) for S in dp.S_array)                           # :76
    # gains should match the LQR solution (:87-106)
    S_lqr = [None] * (N + 1)
    S_lqr[N] = ap.Q_array[N]
    for i in reversed(range(N)):
        Q_, R_, A_, B_ = ap.Q_array[i], ap.R_array[i], ap.A_array[i], ap.B_array[i]
        Sn = S_lqr[i + 1]
        S_lqr[i] = Q_ + A_.T @ Sn @ A_ - A_.T @ Sn @ B_ @ np.linalg.solve(R_ + B_.T @ Sn @ B_, B_.T @ Sn @ A_)
    for i in range(N):
        R_, A_, B_ = ap.R_array[i], ap.A_array[i], ap.B_array[i]
        L_lqr = -np.linalg.solve(R_ + B_.T @ S_lqr[i + 1] @ B_, B_ @ S_lqr[i + 1] @ A_)
        assert np.allclose(L_lqr, solver.L_array[i], rtol=1e-8)                                 # :105
    # nominal control offsets are 0: u + dl - L x = 0 (:108)
    for i in range(N):
        assert np.linalg.norm(u_array[i] + dl_new[i] - solver.L_array[i] @ x_array[i]) < 1e-8

    dp2, dl_new2 = R.solve_approximate_dp_(solver, ap, False, theta=1e-8)                       # :110
    assert all(np.array_equal(S, S.T) for S in dp2.S_array)
    assert np.isclose(dp.s_array[0], dp2.s_array[0], rtol=1e-5)                                 # :124
    assert all(np.allclose(a, b) for a, b in zip(dl_new, dl_new2))                              # :125

    R.solve_approximate_dp_(solver, ap, False, theta=0.0)
    dp3 = R.solve_approximate_dp(ap, solver.L_array, dl_new, theta=0.0, mu=0.0, backend=be)
    assert dp3.s_array == dp.s_array                                                            # :130 (bitwise)

    R.line_search_(solver, prob, dl_new, 0.0, False)                                            # :133
    assert np.isclose(solver.value_current, dp.s_array[0])                                      # :134

    solver = R.ILEQGSolver(prob, backend=be)                                                    # :137-141
    IL.initialize_(solver, prob, np.zeros(2), u_array, 0.0)
    R.increase_mu_and_delta_(solver)
    assert solver.delta == 4.0 and solver.mu == 1e-6
    solver = R.ILEQGSolver(prob, backend=be)                                                    # :144-148
    IL.initialize_(solver, prob, np.zeros(2), u_array, 0.0)
    R.decrease_mu_and_delta_(solver)
    assert solver.delta == 0.5 and solver.mu == 0.0


def test_ileqg_nonlinear_model(backend):
    """:150-174: f = x.^1.3 + u.^1.5, c = sum(x.^2.5 + u.^2.5), h = 1, W = 0.01 I, N = 10, theta = 0.5"""
    be = backend
    cost = R.PowerLawCost(2.5, 1.0)
    N, theta = 10, 0.5
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(R.PowerLawDynamics(1.3, 1.5), cost.c, cost.h,
                                                             R.ConstantCovariance(0.01 * np.eye(2)), N)
    u_array = [0.1 * np.ones(2) for _ in range(N)]
    solver = R.ILEQGSolver(prob, backend=be)
    IL.initialize_(solver, prob, np.zeros(2), u_array, theta)
    ap = R.approximate_model(prob, solver.l_array, solver.x_array, solver.A_array, solver.B_array, backend=be)
    dp, dl = R.solve_approximate_dp_(solver, ap, False, theta=theta)
    R.line_search_(solver, prob, dl, theta, False)
    assert len(solver.eps_history) == 1                                                         # :168
    assert solver.eps_history[0][0] == 1.0                                                      # :169
    assert solver.eps_history[0][1] < 0.0                                                       # :170
    assert np.isclose(solver.eps_history[0][1], -0.12464762861727641, rtol=1e-9)  # SURVEY Appendix B anchor

    x_array, l_array, L_array, value, eps_hist = IL.solve_(solver, prob, np.zeros(2), u_array, theta=0.0, verbose=False)
    assert all(np.all(np.abs(x) < 1e-4) for x in x_array)                                       # :174
    # SURVEY Appendix B / BASELINE.md section 4 anchors (independent numpy restatement made during the survey)
    assert np.isclose(value, 1.0029075497782471, rtol=1e-12)
    assert solver.iter_current == 4 and len(eps_hist) == 4


def test_whole_solve_equals_host_stepping(backend):
    """solve_ (one persistent kernel) must equal initialize_ + step_ loops built from the component kernels."""
    be = backend
    cost = R.PowerLawCost(2.5, 1.0)
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(R.PowerLawDynamics(1.3, 1.5), cost.c, cost.h,
                                                             R.ConstantCovariance(0.01 * np.eye(2)), 10)
    u_array = [0.1 * np.ones(2) for _ in range(10)]
    for theta in (0.0, 0.43):
        a = R.ILEQGSolver(prob, backend=be)
        xa, la, La, va, ha = IL.solve_(a, prob, np.zeros(2), u_array, theta=theta)
        b = R.ILEQGSolver(prob, backend=be)
        IL.initialize_(b, prob, np.zeros(2), u_array, theta)
        while True:
            IL.step_(b, prob, theta)
            if b.d > b.d_current and b.mu <= b.mu_min:
                break
            if b.iter_current == b.iter_max:
                break
        assert b.iter_current == a.iter_current
        # trajectories and gains never depend on the scalar s, so they agree bit for bit; the value itself differs in
        # the last ulps because the persistent kernel takes ONE log per pass (log of the product of the stage
        # determinants) while the component pass, like the reference, takes one per stage
        assert np.isclose(va, b.value_current, rtol=1e-13)
        assert all(np.array_equal(p, q) for p, q in zip(xa, b.x_array))
        assert all(np.array_equal(p, q) for p, q in zip(La, b.L_array))
        assert len(ha) == len(b.eps_history)
        assert all(e1 == e2 and abs(d1 - d2) <= 1e-13 * max(1.0, abs(va)) for (e1, d1), (e2, d2) in zip(ha, b.eps_history))


def test_neurotic_breakdown_and_domain_error(backend):
    be = backend
    cost = R.PowerLawCost(2.5, 1.0)
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(R.PowerLawDynamics(1.3, 1.5), cost.c, cost.h,
                                                             R.ConstantCovariance(0.01 * np.eye(2)), 10)
    u_array = [0.1 * np.ones(2) for _ in range(10)]
    s = R.ILEQGSolver(prob, backend=be)
    try:
        IL.solve_(s, prob, np.zeros(2), u_array, theta=40.0)  # beyond the breakdown point (~30.78)
        raise SystemError("expected NotPositiveDefinite")
    except R.NotPositiveDefinite:
        pass
    IL.solve_(s, prob, np.zeros(2), u_array, theta=30.7)
    try:
        IL.solve_(s, prob, np.zeros(2), [-0.1 * np.ones(2) for _ in range(10)], theta=0.0)  # (-0.1)^1.5
        raise SystemError("expected DomainError")
    except R.DomainError:
        pass
