"""test/pets_test.jl of the reference re-expressed against the host mirror, plus device-vs-oracle parity of the
PETS pipeline (sample -> rollout -> particle mean -> stable top-k elites -> smoothed refit) with injected randomness."""
import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import pets as PT
from ratilqr_b200 import workloads as wl


def _pets_problem(N=20):
    cost = R.L1ControlCost(1.0)                                     # c = sum(abs.(u)), h = 1.0  (pets_test.jl:16-17)
    fs = R.DeviceStochasticDynamics(R.SingleIntegrator(1.0), W=np.eye(2), noise_kind=1, noise_scale=1.0)  # x+u+rand (:15)
    return R.FiniteHorizonGenerativeOptimalControlProblem(fs, cost.c, cost.h, N)


def test_pets_reference_suite(full_backend):
    be = full_backend
    N = 20
    problem = _pets_problem(N)
    mu_init = [np.zeros(2) for _ in range(N)]
    Sg_init = [np.eye(2) for _ in range(N)]
    s = R.CrossEntropyDirectOptimizationSolver(mu_init, Sg_init, num_control_samples=20, num_trajectory_samples=100,
                                               num_elite=5, iter_max=20, smoothing_factor=0.1, backend=be)
    assert s.N == N and s.iter_current == 0                                                  # :25-26
    assert all(np.array_equal(a, b) for a, b in zip(s.mu_array, mu_init))                    # :27
    assert all(np.array_equal(a, b) for a, b in zip(s.Sigma_array, Sg_init))                 # :28
    s.iter_current = 10
    s.mu_array = [np.ones(2) for _ in range(N)]
    s.Sigma_array = [0.1 * np.eye(2) for _ in range(N)]
    PT.initialize_(s)
    assert s.iter_current == 0 and all(np.array_equal(a, b) for a, b in zip(s.mu_array, mu_init))  # :35-37

    rng = np.random.default_rng(1234)
    seqs = [[rng.random(2) for _ in range(N)] for _ in range(s.num_control_samples)]
    x_init = np.zeros(2)
    noise = PT._noise_for(problem, s, np.random.default_rng(7), s.num_control_samples)
    cost_serial = PT.compute_cost_serial(s, problem, x_init, seqs, None, noise=noise)
    cost_array = PT.compute_cost(s, problem, x_init, seqs, None, noise=noise)
    assert np.array_equal(cost_array, cost_serial)                                           # :48
    assert len(cost_array) == s.num_control_samples                                          # :51
    for ii in range(s.num_control_samples):  # hand-rolled rollout: c does not depend on x (:52-61)
        c = sum(problem.c(tt, None, seqs[ii][tt]) for tt in range(N)) + problem.h(None)
        assert np.isclose(c, cost_array[ii])

    elite = PT.get_elite_samples(s, seqs, cost_array)                                         # :64
    assert len(elite) == s.num_elite
    order = sorted(range(len(cost_array)), key=lambda i: cost_array[i])[:s.num_elite]          # stable
    assert all(all(np.array_equal(a, b) for a, b in zip(e, seqs[i])) for e, i in zip(elite, order))  # :68

    mu_new, Sg_new = PT.compute_new_distribution(s, elite)                                    # :71
    assert len(mu_new) == N and len(Sg_new) == N and all(S.shape == (2, 2) for S in Sg_new)
    a = s.smoothing_factor
    for tt in range(N):
        el = np.stack([e[tt] for e in elite])
        assert np.allclose(mu_new[tt], (1 - a) * el.mean(0) + a * s.mu_array[tt])             # :78-79
        assert np.allclose(Sg_new[tt], (1 - a) * np.diag(el.var(0, ddof=1)) + a * s.Sigma_array[tt])  # :80-81

    PT.step_(s, problem, x_init, np.random.default_rng(1234))                                  # :85-86
    assert s.iter_current == 1
    PT.solve_(s, problem, x_init, np.random.default_rng(1234))                                 # :89
    assert s.iter_current == s.iter_max                                                        # :91


def test_pets_elites_are_stable_and_nan_last(full_backend):
    be = full_backend
    rng = np.random.default_rng(0)
    m, N, Cn = 2, 3, 40
    ctrl = rng.standard_normal((m, N, Cn))
    cost = rng.integers(0, 5, Cn).astype(float)  # many ties
    cost[3] = np.nan
    cost[7] = np.inf
    mu, Sg, idx = be.pets_refit(ctrl, cost, 9, 0.25, np.zeros((m, N)), np.tile(np.eye(m)[:, :, None], (1, 1, N)))
    order = sorted(range(Cn), key=lambda i: (np.isnan(cost[i]), cost[i]))[:9]
    assert list(idx) == order


def test_pets_solve_oracle_runs(oracle_be):
    """the oracle's whole-loop restatement is well behaved on the cart-pole ensemble (inputs of the gpu parity test)"""
    prob, x0 = wl.c4_problem(N=10)
    spec = prob.spec()
    gen = prob.f_stochastic.gen()
    rng = np.random.Generator(np.random.Philox(key=42))
    Cn, Kp, iters, ne = 24, 10, 3, 5
    z = rng.standard_normal((1, 10, Cn, iters))
    noise = np.einsum("ij,jkpct->ikpct", np.linalg.cholesky(prob.f_stochastic.W), rng.standard_normal((4, 10, Kp, Cn, iters)))
    mu0, Sg0 = np.zeros((1, 10)), np.tile(np.array([[4.0]])[:, :, None], (1, 1, 10))
    o = oracle_be.pets_solve(spec, x0, mu0, Sg0, Cn, Kp, ne, iters, 0.1, z_inject=z, noise=noise, gen=gen)
    assert np.all(np.isfinite(o[0])) and np.all(o[1] > 0)


@pytest.mark.gpu
def test_pets_solve_injected_gpu_vs_oracle(gpu_be, oracle_be):
    prob, x0 = wl.c4_problem(N=10)
    spec = prob.spec()
    gen = prob.f_stochastic.gen()
    rng = np.random.Generator(np.random.Philox(key=42))
    Cn, Kp, iters, ne = 24, 10, 3, 5
    z = rng.standard_normal((1, 10, Cn, iters))
    noise = np.einsum("ij,jkpct->ikpct", np.linalg.cholesky(prob.f_stochastic.W), rng.standard_normal((4, 10, Kp, Cn, iters)))
    mu0, Sg0 = np.zeros((1, 10)), np.tile(np.array([[4.0]])[:, :, None], (1, 1, 10))
    o = oracle_be.pets_solve(spec, x0, mu0, Sg0, Cn, Kp, ne, iters, 0.1, z_inject=z, noise=noise, gen=gen)
    g = gpu_be.pets_solve(spec, x0, mu0, Sg0, Cn, Kp, ne, iters, 0.1, z_inject=z, noise=noise, gen=gen)
    assert np.allclose(g[0], o[0], rtol=1e-9, atol=1e-12) and np.allclose(g[1], o[1], rtol=1e-9, atol=1e-14)


@pytest.mark.gpu
def test_pets_philox_statistical(gpu_be):
    """On-device RNG: two seeds give statistically equal elite distributions; same seed is bit-reproducible."""
    prob, x0 = wl.c4_problem(N=10)
    spec = prob.spec()
    gen = prob.f_stochastic.gen()
    mu0, Sg0 = np.zeros((1, 10)), np.tile(np.array([[4.0]])[:, :, None], (1, 1, 10))
    a = gpu_be.pets_solve(spec, x0, mu0, Sg0, 2048, 30, 204, 3, 0.1, seed=1, gen=gen)
    a2 = gpu_be.pets_solve(spec, x0, mu0, Sg0, 2048, 30, 204, 3, 0.1, seed=1, gen=gen)
    b = gpu_be.pets_solve(spec, x0, mu0, Sg0, 2048, 30, 204, 3, 0.1, seed=2, gen=gen)
    assert np.array_equal(a[0], a2[0]) and np.array_equal(a[1], a2[1])
    # three CEM iterations amplify sampling differences: require agreement within one posterior std and
    # the same order of magnitude of the variances (the two runs share nothing but the distribution)
    sd = np.sqrt(np.maximum(a[1][0, 0], b[1][0, 0]))
    assert np.all(np.abs(a[0] - b[0]) < 1.0 * sd + 1e-3)
    assert np.all(a[1] < 4.0 * b[1]) and np.all(b[1] < 4.0 * a[1])


@pytest.mark.gpu
def test_mc_rollout_philox_statistical(gpu_be, oracle_be):
    """Philox mode of the MC kernel: mean cost agrees with an injected-noise oracle run within sampling error."""
    prob, x0, u = wl.c2_problem()
    spec = prob.spec()
    sol = oracle_be.ileqg_solve_batch(spec, x0, u, [1.0])
    xbar, l, L = sol["x"][..., 0], sol["l"][..., 0], sol["L"][..., 0]
    g = gpu_be.mc_rollout(spec, xbar, l, L, 8192, seed=5, theta_risk=1.0)
    w = np.einsum("ij,jks->iks", np.linalg.cholesky(prob.W(0)), np.random.default_rng(0).standard_normal((4, 50, 8192)))
    o = oracle_be.mc_rollout(spec, xbar, l, L, 8192, noise=w, theta_risk=1.0)
    se = np.sqrt(o["stats"][0, 1] / 8192)
    assert abs(g["stats"][0, 0] - o["stats"][0, 0]) < 6 * se
    assert abs(g["stats"][0, 1] / o["stats"][0, 1] - 1) < 0.2
    assert g["stats"][0, 2] >= g["stats"][0, 0]  # entropic risk >= mean (Jensen)
    g2 = gpu_be.mc_rollout(spec, xbar, l, L, 8192, seed=5, theta_risk=1.0)
    assert np.array_equal(g["J"], g2["J"])


@pytest.mark.gpu
@pytest.mark.parametrize("theta_risk", [0.0, 0.5])
def test_mc_stats_of_many_samples_per_problem(gpu_be, theta_risk):
    """>= 4 chunks of 8,192 samples per problem go through the chunked reductions (k_mc_stats_chunked): mean, unbiased
    variance and the entropic risk (1/theta) log E exp(theta J) against numpy on the returned costs"""
    prob, x0, u = wl.c2_problem(N=12)
    spec = prob.spec()
    r = gpu_be.ileqg_solve_batch(spec, x0, u, [0.5])
    xbar, l, L = r["x"][..., 0], r["l"][..., 0], r["L"][..., 0]
    P, S = 2, 40000
    n0 = gpu_be.launch_count()
    g = gpu_be.mc_rollout(spec, np.stack([xbar, xbar], -1), np.stack([l, l], -1), np.stack([L, L], -1), S, seed=11,
                          theta_risk=theta_risk, P=P)
    assert gpu_be.launch_count() - n0 == 1 + (5 if theta_risk > 0 else 3)
    J = g["J"].reshape(P, S)
    assert np.all(np.isfinite(J)) and not np.array_equal(J[0], J[1])
    for p in range(P):
        mean, var, risk = g["stats"][p]
        assert abs(mean - J[p].mean()) < 1e-12 * abs(J[p].mean())
        assert abs(var - J[p].var(ddof=1)) < 1e-10 * J[p].var(ddof=1)
        ref = J[p].mean() if theta_risk == 0 else (np.log(np.mean(np.exp(theta_risk * J[p] - (theta_risk * J[p]).max()))) + (theta_risk * J[p]).max()) / theta_risk
        assert abs(risk - ref) < 1e-12 * abs(ref)
