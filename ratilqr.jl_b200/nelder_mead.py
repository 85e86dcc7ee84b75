"""Host mirror of src/nelder_mead_bilevel_optimization.jl (RAT iLQR++: 1-D Nelder-Mead over θ).

The reference evaluates the candidate θ of one NM step one after another (2-3 iLEQG solves per
step, :195-240).  Evaluations are pure functions of θ, so `step_` evaluates ALL candidates of a
step speculatively in one batched launch (θ_r, θ_e, both possible θ_c and both shrink points)
and then replays the reference's decision tree on the results: the visited vertices and returned
values are identical to the serial order.
"""
import math
import unicodedata

import numpy as np

from . import _lib
from .ileqg import ILEQGSolver, _stack, solve_ as ileqg_solve_


class NelderMeadBilevelOptimizationSolver:  # :71-128
    def __init__(self, backend=None, speculative=True, **kw):
        kw = {k.replace("ϵ", "eps").replace("ε", "eps"): v for k, v in kw.items()}
        al = {"μ_min_ileqg": "mu_min_ileqg", "Δ_0_ileqg": "delta_0_ileqg", "λ_ileqg": "lam_ileqg", "α": "alpha",
              "β": "beta", "γ": "gamma", "λ": "lam", "θ_high_init": "theta_high_init", "θ_low_init": "theta_low_init"}
        al = {unicodedata.normalize("NFKC", k): v for k, v in al.items()}
        kw = {al.get(unicodedata.normalize("NFKC", k), k): v for k, v in kw.items()}
        o = dict(mu_min_ileqg=1e-6, delta_0_ileqg=2.0, lam_ileqg=0.5, d_ileqg=1e-2, iter_max_ileqg=100,
                 adaptive_eps_init_ileqg=False, eps_init_ileqg=1.0, eps_min_ileqg=1e-6, f_returns_jacobian=False,
                 alpha=1.0, beta=2.0, gamma=0.5, eps=1e-2, lam=0.5, iter_max=100, theta_high_init=3.0,
                 theta_low_init=1e-8)
        unknown = set(kw) - set(o)
        if unknown:
            raise TypeError(f"unknown keyword arguments {sorted(unknown)}")
        o.update(kw)
        self.__dict__.update(o)
        self.theta_high, self.theta_low = self.theta_high_init, self.theta_low_init
        self.iter_current = 0
        self.c_high = self.c_low = None  # persist across solve_ calls, like the reference (:98-99)
        self.backend, self.speculative = backend, speculative
        self.n_evals = 0

    def _be(self):
        return self.backend or _lib.default_backend()

    def ileqg_kwargs(self):
        return dict(mu_min=self.mu_min_ileqg, delta_0=self.delta_0_ileqg, lam=self.lam_ileqg, d=self.d_ileqg,
                    iter_max=self.iter_max_ileqg, adaptive_eps_init=self.adaptive_eps_init_ileqg,
                    eps_init=self.eps_init_ileqg, eps_min=self.eps_min_ileqg,
                    f_returns_jacobian=self.f_returns_jacobian)


def _costs(nm, problem, x, u_array, thetas, kl_bound):
    ileqg = ILEQGSolver(problem, **nm.ileqg_kwargs())
    cost, _ = nm._be().ce_costs(problem.spec(), np.asarray(x, float), _stack(u_array),
                                np.asarray(thetas, float), kl_bound, opts=ileqg.opts())
    return cost


def compute_cost_worker(nm, problem, x, u_array, theta, kl_bound):  # :134-158
    nm.n_evals += 1
    return float(_costs(nm, problem, x, u_array, [theta], kl_bound)[0])


def initialize_(nm):  # :164-168
    nm.iter_current = 0
    nm.theta_low = nm.theta_low_init
    nm.theta_high = nm.theta_high_init


def step_(nm, problem, x, u_array, kl_bound, verbose=False):  # step! :174-252
    nm.iter_current += 1
    if nm.c_high < nm.c_low:
        nm.theta_low, nm.theta_high = nm.theta_high, nm.theta_low
        nm.c_low, nm.c_high = nm.c_high, nm.c_low
    th_m, lo = nm.theta_low, nm.theta_low_init
    th_r = max(lo, th_m + nm.alpha * (th_m - nm.theta_high))
    cache = {}
    if nm.speculative:
        # every θ the decision tree below can ask for, evaluated in one launch
        th_e = max(lo, th_m + nm.beta * (th_r - th_m))
        th_c1 = max(lo, th_m + nm.gamma * (th_r - th_m))          # contraction if θ_high <- θ_r
        th_c2 = max(lo, th_m + nm.gamma * (nm.theta_high - th_m))  # contraction if θ_high kept
        th_s1 = (th_r + nm.theta_low) / 2                          # shrink points
        th_s2 = (nm.theta_high + nm.theta_low) / 2
        cand = [th_r, th_e, th_c1, th_c2, th_s1, th_s2]
        for t, c in zip(cand, _costs(nm, problem, x, u_array, cand, kl_bound)):
            cache[t] = float(c)

    def cost_at(t):
        nm.n_evals += 1
        if t in cache:
            return cache[t]
        return float(_costs(nm, problem, x, u_array, [t], kl_bound)[0])

    c_r = cost_at(th_r)
    if c_r < nm.c_low:
        th_e = max(lo, th_m + nm.beta * (th_r - th_m))
        c_e = cost_at(th_e)
        if c_e < c_r:
            nm.theta_high, nm.c_high = th_e, c_e
        else:
            nm.theta_high, nm.c_high = th_r, c_r
    else:
        if c_r < nm.c_high:
            nm.theta_high, nm.c_high = th_r, c_r
        th_c = max(lo, th_m + nm.gamma * (nm.theta_high - th_m))
        c_c = cost_at(th_c)
        if c_c > nm.c_high:
            nm.theta_high = (nm.theta_high + nm.theta_low) / 2
            nm.c_high = cost_at(nm.theta_high)
        else:
            nm.theta_high, nm.c_high = th_c, c_c


def solve_(nm, problem, x_0, u_array, verbose=False, **kw):
    """solve! (:276-352) -> (θ_opt, x_array, l_array, L_array, value)."""
    kl_bound = float(kw["kl_bound"])
    assert kl_bound >= 0, "KL Divergence Bound must be non-negative"
    initialize_(nm)
    if kl_bound > 0:
        if nm.c_high is None:
            while True:
                nm.c_high = compute_cost_worker(nm, problem, x_0, u_array, nm.theta_high, kl_bound)
                if not math.isinf(nm.c_high):
                    break
                nm.theta_high *= nm.lam
                nm.theta_high_init *= nm.lam
        if nm.c_low is None:
            while True:
                nm.c_low = compute_cost_worker(nm, problem, x_0, u_array, nm.theta_low, kl_bound)
                if not math.isinf(nm.c_low):
                    break
                nm.theta_low *= nm.lam
                nm.theta_low_init *= nm.lam
        while True:
            step_(nm, problem, x_0, u_array, kl_bound, verbose)
            c_mean = (nm.c_low + nm.c_high) / 2
            stdev = math.sqrt(0.5 * ((nm.c_high - c_mean) ** 2 + (nm.c_low - c_mean) ** 2))
            if stdev < nm.eps:
                break
            if nm.iter_current == nm.iter_max:
                break
        theta_opt = nm.theta_low
    else:
        theta_opt = 0.0
    ileqg = ILEQGSolver(problem, backend=nm._be(), **nm.ileqg_kwargs())
    x_array, l_array, L_array, value, _ = ileqg_solve_(ileqg, problem, x_0, u_array, theta=theta_opt, verbose=False)
    if kl_bound > 0:
        return theta_opt, x_array, l_array, L_array, value + kl_bound / theta_opt
    return theta_opt, x_array, l_array, L_array, value


def solve_fleet_(nm, problem, x0, u_init, kl_bound, cost_params=None, state=None, want=("x", "l", "L")):
    """solve! for a FLEET of independent problems in one call (additive API, ratilqr_nm_solve_fleet): the six candidate
    theta of every problem's step are solved in one launch and the decision tree is replayed per problem on the device.
    `state` (from a previous call's result) carries theta_*_init and the vertex costs across calls."""
    spec = problem.spec(cost_params=cost_params)
    return nm._be().nm_solve_fleet(spec, x0, u_init, kl_bound, state=state, alpha=nm.alpha, beta=nm.beta, gamma=nm.gamma,
                                   eps=nm.eps, lam=nm.lam, iter_max=nm.iter_max, theta_high_init=nm.theta_high_init,
                                   theta_low_init=nm.theta_low_init, opts=ILEQGSolver(problem, **nm.ileqg_kwargs()).opts(), want=want)
