"""Closed-loop Monte Carlo evaluation and PETS rollouts under the TRUE noise model: the accurate
Gaussian mixture behind `f_stochastic(x, u, rng, use_true_model=true)` (src/optimal_control_problems.jl:85-86,102-115)
versus the single Gaussian the planner assumes.  The mixture is sampled on the device with Philox; the g++ build of the
same sampler (tests/_hostemu) gives the CPU checks and the GPU-vs-CPU parity."""
import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import _capi

# the reference's own example: 0.5 N(0, 0.5 I) + 0.5 N(1, I) on a 2-D single integrator x' = x + u + w
DOCS_MIX = dict(weights=[0.5, 0.5], means=np.array([[0.0, 1.0], [0.0, 1.0]]),
                covs=np.stack([0.5 * np.eye(2), np.eye(2)], axis=-1))


def single_integrator(N=6):
    cost = R.QuadraticCost(2, 2, Q=np.eye(2), R=np.eye(2), Qf=np.eye(2))
    f = R.SingleIntegrator(1.0)
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, R.ConstantCovariance(0.5 * np.eye(2)), N)
    return prob.spec(), N


@pytest.fixture(params=["hostemu", pytest.param("gpu", marks=pytest.mark.gpu)])
def dut(request):
    return request.getfixturevalue(request.param + "_be")


def test_mixture_sampler_moments(dut):
    """zero policy on the single integrator: the increments of the rollout ARE the noise draws"""
    spec, N = single_integrator()
    S = 20000
    xbar, l, L = np.zeros((2, N + 1)), np.zeros((2, N)), np.zeros((2, 2, N))
    r = dut.mc_rollout_true_model(spec, xbar, l, L, S, DOCS_MIX, seed=5, want_x=True)
    w = np.diff(r["x"], axis=1).reshape(2, -1)  # (2, N*S)
    cnt = w.shape[1]
    mean_true = np.array([0.5, 0.5])
    cov_true = 0.75 * np.eye(2) + 0.25 * np.ones((2, 2))
    assert np.all(np.abs(w.mean(axis=1) - mean_true) < 5 * np.sqrt(np.diag(cov_true) / cnt))
    assert np.allclose(np.cov(w), cov_true, atol=0.03)
    # bimodal, not Gaussian: excess kurtosis of the sum coordinate differs from 0 and matches the mixture's
    s = (w[0] + w[1]) / np.sqrt(2)
    z = (s - s.mean()) / s.std()
    m4 = 0.5 * (3 * 0.5 ** 2 + 6 * 0.5 * 0.5 + 0.25) + 0.5 * (3 * 1.0 + 6 * 1.0 * 0.5 + 0.25)  # E[(s-mean)^4], mean shift +-1/sqrt2
    var = 0.75 + 0.25 * 2 * 0.5 * 2  # = cov of s: 0.75 + 0.5
    assert abs(np.mean(z ** 4) - m4 / var ** 2) < 0.08


def test_single_component_mixture_equals_gaussian_model(dut):
    """k = 1, mean 0, cov = W: the same Philox normals, hence the same rollouts as the planner's model, bit for bit"""
    spec, N = single_integrator()
    rng = np.random.default_rng(2)
    xbar = np.cumsum(rng.standard_normal((2, N + 1)), axis=1) * 0.1
    l, L = 0.1 * rng.standard_normal((2, N)), -0.3 * np.tile(np.eye(2)[:, :, None], (1, 1, N))
    mix = dict(weights=[1.0], means=np.zeros((2, 1)), covs=0.5 * np.eye(2)[:, :, None])
    a = dut.mc_rollout_true_model(spec, xbar, l, L, 512, mix, seed=9)
    b = dut.mc_rollout(spec, xbar, l, L, 512, seed=9)
    assert np.array_equal(a["J"], b["J"])


def test_true_model_mc_matches_injected_mixture_noise(dut):
    """E[J] under device-sampled mixture noise == E[J] with host-drawn mixture noise injected (5 sigma)"""
    spec, N = single_integrator()
    rng = np.random.default_rng(4)
    xbar, l = np.zeros((2, N + 1)), np.zeros((2, N))
    L = -0.5 * np.tile(np.eye(2)[:, :, None], (1, 1, N))
    S = 20000
    a = dut.mc_rollout_true_model(spec, xbar, l, L, S, DOCS_MIX, seed=21)["J"]
    comp = rng.random((N, S)) < 0.5
    w = np.where(comp[None], np.sqrt(0.5) * rng.standard_normal((2, N, S)), 1.0 + rng.standard_normal((2, N, S)))
    b = dut.mc_rollout(spec, xbar, l, L, S, noise=w)["J"]
    se = np.sqrt(a.var() / S + b.var() / S)
    assert abs(a.mean() - b.mean()) < 5 * se
    # and the planner's Gaussian model underestimates the cost the true model produces
    g = dut.mc_rollout(spec, xbar, l, L, S, seed=3)["J"]
    assert a.mean() > g.mean() + 10 * se


def test_pets_use_true_model(dut):
    """compute_cost(..., use_true_model): k = 1 mixture == planner's model exactly; the docs mixture costs more"""
    N = 5
    cost = R.QuadraticCost(2, 2, Q=np.eye(2), R=0.1 * np.eye(2), Qf=np.eye(2))
    fs = R.DeviceStochasticDynamics(R.SingleIntegrator(1.0), W=0.5 * np.eye(2), true_mixture=DOCS_MIX)
    prob = R.FiniteHorizonGenerativeOptimalControlProblem(fs, cost.c, cost.h, N)
    rng = np.random.default_rng(0)
    ctr = 0.2 * rng.standard_normal((2, N, 16))
    x0 = np.zeros(2)
    plan = dut.pets_costs(prob.spec(), x0, ctr, 256, seed=7, gen=fs.gen())
    same = dut.pets_costs(prob.spec(), x0, ctr, 256, seed=7,
                          gen=dict(fs.gen(), use_true_model=True,
                                   true_model=dict(weights=[1.0], means=np.zeros((2, 1)), covs=0.5 * np.eye(2)[:, :, None])))
    true = dut.pets_costs(prob.spec(), x0, ctr, 256, seed=7, gen=fs.gen(use_true_model=True))
    assert np.array_equal(plan, same)
    assert np.all(true > plan)
    # host mirror: compute_cost_serial(..., use_true_model=True) with on-device Philox
    s = R.CrossEntropyDirectOptimizationSolver([np.zeros(2)] * N, [np.eye(2)] * N, num_control_samples=16,
                                               num_trajectory_samples=256, backend=dut)
    seqs = [[ctr[:, t, i] for t in range(N)] for i in range(16)]
    assert np.array_equal(R.pets.compute_cost_serial(s, prob, x0, seqs, 7, use_true_model=True), true)


def test_bad_mixture_is_rejected(dut):
    spec, N = single_integrator()
    xbar, l, L = np.zeros((2, N + 1)), np.zeros((2, N)), np.zeros((2, 2, N))
    bad = dict(weights=[0.5, 0.5], means=np.zeros((2, 2)), covs=np.stack([np.eye(2), -np.eye(2)], axis=-1))
    with pytest.raises(_capi.ApiError):
        dut.mc_rollout_true_model(spec, xbar, l, L, 8, bad)
    with pytest.raises(_capi.ApiError):
        dut.mc_rollout_true_model(spec, xbar, l, L, 8, dict(DOCS_MIX, weights=[1.0, -1.0]))


@pytest.mark.gpu
def test_gpu_matches_host_build_of_the_sampler(gpu_be, hostemu_be):
    spec, N = single_integrator()
    rng = np.random.default_rng(8)
    xbar = 0.1 * rng.standard_normal((2, N + 1))
    l, L = 0.1 * rng.standard_normal((2, N)), -0.4 * np.tile(np.eye(2)[:, :, None], (1, 1, N))
    g = gpu_be.mc_rollout_true_model(spec, xbar, l, L, 4096, DOCS_MIX, seed=33, want_x=True)
    h = hostemu_be.mc_rollout_true_model(spec, xbar, l, L, 4096, DOCS_MIX, seed=33, want_x=True)
    assert np.max(np.abs(g["x"] - h["x"])) < 1e-12 and np.max(np.abs(g["J"] - h["J"])) < 1e-11
    assert abs(g["stats"][0, 0] - h["J"].mean()) < 1e-10


@pytest.mark.gpu
def test_fleet_mpc_under_true_mixture_noise(gpu_be):
    """receding-horizon RAT iLQR planned under the Gaussian model, executed against a bimodal true disturbance"""
    from ratilqr_b200 import workloads as wl
    from ratilqr_b200.mpc import run_fleet_mpc
    P = 24
    prob, cps, x0, u = wl.fleet(P, key=4, N=20)
    W = np.asarray(prob.W(0))
    mix = dict(weights=[0.7, 0.3], means=np.stack([np.zeros(4), np.array([0.02, -0.02, 0.0, 0.0])], axis=1),
               covs=np.stack([W, 4.0 * W], axis=-1))
    out = run_fleet_mpc(gpu_be, prob, cps, x0, steps=30, kl_bound=0.1, rng=np.random.default_rng(1), true_mixture=mix)
    goals = cps[:, 5:7]
    d0 = np.linalg.norm(x0[:2].T - goals, axis=1)
    d1 = np.linalg.norm(out["x"][:2, -1].T - goals, axis=1)
    assert np.all(np.isfinite(out["x"])) and np.all(out["theta"] >= 0)
    assert np.median(d1) < 0.5 * np.median(d0)  # every system made clear progress towards its goal despite the model error
