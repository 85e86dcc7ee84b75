"""Synthetic workloads C1-C5 of BASELINE.json / SURVEY.md 8(d), frozen here so that tests, bench.py
and smoke() all build bit-identical inputs.  All randomness comes from numpy's Philox bit
generator with fixed keys."""
import numpy as np

from .models import (CartPole, ConstantCovariance, DeviceStochasticDynamics, PowerLawCost, PowerLawDynamics,
                     QuadraticCost, Quadrotor, Unicycle)
from .problems import FiniteHorizonGenerativeOptimalControlProblem, FiniteHorizonRiskSensitiveOptimalControlProblem


def _rng(key):
    return np.random.Generator(np.random.Philox(key=key))


def positive_thetas(count, mu=1.0, sigma=2.0, key=20201028):
    """first `count` positive draws of mu + sigma*z (get_positive_samples, cross_entropy...:233-246)"""
    rng = _rng(key)
    out = []
    while len(out) < count:
        z = rng.standard_normal(2 * count)
        t = mu + sigma * z
        out.extend(t[t > 0.0].tolist())
    return np.array(out[:count])


def c1_problem():
    """the reference's shipped test problem (test/ileqg_test.jl:151-155, test/cross_entropy..._test.jl:13-21)"""
    cost = PowerLawCost(2.5, 1.0)
    prob = FiniteHorizonRiskSensitiveOptimalControlProblem(PowerLawDynamics(1.3, 1.5), cost.c, cost.h,
                                                           ConstantCovariance(0.01 * np.eye(2)), 10)
    return prob, np.zeros(2), 0.1 * np.ones((2, 10))


def unicycle_cost(goal=(5.0, 5.0, 0.0, 0.0), scale=0.01):
    """Goal-tracking quadratic cost.  SURVEY.md 8(d) proposed Q = diag(1,1,.1,.1), R = .1 I, Qf = 10 Q;
    with that scaling iLEQG is feasible only for theta < 0.099, i.e. 94% of the theta ~ N(1,2) population
    of C2 dies in initialize! (neurotic breakdown) and the batch would measure nothing.  The same cost
    scaled by 0.01 moves the breakdown to theta ~ 9.9 while leaving the optimal policy unchanged."""
    Q = scale * np.diag([1.0, 1.0, 0.1, 0.1])
    return QuadraticCost(4, 2, Q=Q, R=scale * np.diag([0.1, 0.1]), Qf=10.0 * Q, xg=np.asarray(goal, float))


def c2_problem(N=50, dt=0.1):
    """batched iLEQG, 4-state unicycle, T=50 (BASELINE.json configs[1])"""
    cost = unicycle_cost()
    W = np.diag([1e-2, 1e-2, 1e-3, 1e-2]) * dt
    prob = FiniteHorizonRiskSensitiveOptimalControlProblem(Unicycle(dt), cost.c, cost.h, ConstantCovariance(W), N)
    x0 = np.array([0.0, 0.0, 0.0, 1.0])
    return prob, x0, np.zeros((2, N))


def c2_thetas(count=1024):
    return positive_thetas(count)


def fleet(P, key=7, N=50, dt=0.1):
    """C5-style fleet: P independent unicycle problems, x0 and goal uniform in a box.
    Returns (problem, cost_params (P, ncp), x0 (4, P), u_init (2, N))."""
    prob, _, u = c2_problem(N, dt)
    rng = _rng(key)
    x0 = np.zeros((4, P))
    x0[0] = rng.uniform(-1.0, 1.0, P)
    x0[1] = rng.uniform(-1.0, 1.0, P)
    x0[2] = rng.uniform(-0.5, 0.5, P)
    x0[3] = rng.uniform(0.5, 1.5, P)
    goals = np.zeros((P, 4))
    goals[:, 0] = rng.uniform(3.0, 6.0, P)
    goals[:, 1] = rng.uniform(3.0, 6.0, P)
    cost = unicycle_cost()
    cps = np.stack([cost.params(xg=g) for g in goals])
    return prob, cps, x0, u


def c3_problem(N=40, dt=0.05):
    """RAT iLQR++ on the 12-state quadrotor: hover-to-waypoint quadratic cost."""
    f = Quadrotor(dt)
    Q = np.diag([1.0] * 3 + [0.1] * 3 + [0.1] * 3 + [0.01] * 3)
    goal = np.zeros(12)
    goal[:3] = [1.0, 1.0, 1.0]
    cost = QuadraticCost(12, 4, Q=Q, R=np.diag([0.01, 1.0, 1.0, 1.0]), Qf=10.0 * Q, xg=goal)
    W = np.diag([1e-4] * 3 + [1e-4] * 3 + [1e-3] * 3 + [1e-3] * 3)
    prob = FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, ConstantCovariance(W), N)
    u = np.zeros((4, N))
    u[0] = f.params[1] * f.params[2]  # hover thrust m*g
    return prob, np.zeros(12), u


def c4_problem(N=30, dt=0.02, n_ensemble=5):
    """PETS CEM on cart-pole: 5-member parameter ensemble, additive Gaussian noise."""
    f = CartPole(dt)
    Q = np.diag([0.1, 1.0, 0.01, 0.01])
    goal = np.array([0.0, np.pi, 0.0, 0.0])  # upright (theta measured from the downward vertical)
    cost = QuadraticCost(4, 1, Q=Q, R=np.array([[1e-3]]), Qf=10.0 * Q, xg=goal)
    W = np.diag([1e-6, 1e-6, 1e-4, 1e-4])
    rng = _rng(4)
    ens = np.tile(f.params, (n_ensemble, 1))
    ens[:, 1] *= 1.0 + 0.05 * rng.standard_normal(n_ensemble)  # cart mass
    ens[:, 2] *= 1.0 + 0.05 * rng.standard_normal(n_ensemble)  # pole mass
    fs = DeviceStochasticDynamics(f, W=W, noise_kind=0, ensemble_params=ens)
    prob = FiniteHorizonGenerativeOptimalControlProblem(fs, cost.c, cost.h, N)
    return prob, np.zeros(4)
