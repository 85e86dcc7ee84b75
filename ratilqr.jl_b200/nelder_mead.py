"""Host mirror of src/nelder_mead_bilevel_optimization.jl (RAT iLQR++: 1-D Nelder-Mead over θ).

The reference evaluates the candidate θ of one NM step one after another (2-3 iLEQG solves per
step, :195-240), and every evaluation is a chain of sequential iLEQG passes: a single problem is
pure latency.  Evaluations are pure functions of θ, so the solver here keeps a memo θ -> solve and
fills it SPECULATIVELY, a whole batch per launch, before replaying the reference's decision tree on
it -- the visited vertices, the evaluation count and the returned values are those of the serial
order:
  * vertex search (:283-304): θ_high·λ^j for j = 0..7 together with θ_low in one launch instead of
    one launch per halving;
  * step! (:174-252): all candidates of the step (θ_r, θ_e, both possible θ_c, both shrink points)
    AND, for each of those six possible new vertices and either ordering of the resulting simplex,
    the six candidates of the FOLLOWING step -- two NM steps per launch (<= 78 solves, one warp each);
  * the final solve at θ_opt = θ_low (:325-346) is a vertex already solved: taken from the memo
    (trajectories are retained with the costs).
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


class _Memo:
    """θ -> (cost, solve) for one solve! call: evaluations are pure, so a θ is solved at most once"""

    def __init__(self, nm, problem, x, u_array, kl_bound):
        self.nm, self.problem, self.x, self.u, self.kl = nm, problem, np.asarray(x, float), _stack(u_array), kl_bound
        self.opts = ILEQGSolver(problem, **nm.ileqg_kwargs()).opts()
        self.spec = problem.spec()
        self.cost, self.sol = {}, {}

    def fill(self, thetas):
        todo = sorted({float(t) for t in thetas if float(t) not in self.cost and float(t) > 0.0})
        if not todo:
            return
        r = self.nm._be().ileqg_solve_batch(self.spec, self.x, self.u, np.array(todo), opts=self.opts)
        for i, t in enumerate(todo):
            ok = r["status"][i] == 0
            self.cost[t] = float(r["value"][i] + self.kl / t) if ok else math.inf   # :153, any exception => Inf (:151-156)
            if ok:
                self.sol[t] = (r["x"][..., i].copy(), r["l"][..., i].copy(), r["L"][..., i].copy(), float(r["value"][i]))

    def __call__(self, theta):  # compute_cost_worker (:134-158)
        self.nm.n_evals += 1
        t = float(theta)
        if t not in self.cost:
            self.fill([t])
        return self.cost.get(t, math.inf)


def compute_cost_worker(nm, problem, x, u_array, theta, kl_bound):  # :134-158
    nm.n_evals += 1
    return float(_costs(nm, problem, x, u_array, [theta], kl_bound)[0])


def _step_candidates(nm, th_low, th_high):
    """every θ the decision tree of step! (:195-240) can ask for, given the ORDERED simplex (θ_low, θ_high)"""
    lo, th_m = nm.theta_low_init, th_low
    th_r = max(lo, th_m + nm.alpha * (th_m - th_high))
    return [th_r, max(lo, th_m + nm.beta * (th_r - th_m)), max(lo, th_m + nm.gamma * (th_r - th_m)),
            max(lo, th_m + nm.gamma * (th_high - th_m)), (th_r + th_low) / 2, (th_high + th_low) / 2]


def initialize_(nm):  # :164-168
    nm.iter_current = 0
    nm.theta_low = nm.theta_low_init
    nm.theta_high = nm.theta_high_init


def step_(nm, problem, x, u_array, kl_bound, verbose=False, memo=None):  # step! :174-252
    nm.iter_current += 1
    if nm.c_high < nm.c_low:
        nm.theta_low, nm.theta_high = nm.theta_high, nm.theta_low
        nm.c_low, nm.c_high = nm.c_high, nm.c_low
    th_m, lo = nm.theta_low, nm.theta_low_init
    th_r = max(lo, th_m + nm.alpha * (th_m - nm.theta_high))
    if memo is None:
        memo = _Memo(nm, problem, x, u_array, kl_bound)
    if nm.speculative:
        cand = _step_candidates(nm, nm.theta_low, nm.theta_high)
        if any(float(t) not in memo.cost for t in cand):
            # this step's candidates are not all known: solve them and, in the same launch, the candidates of the NEXT step
            # for every vertex this step can produce and either ordering of the resulting simplex
            nxt = []
            for v in cand:
                nxt += _step_candidates(nm, nm.theta_low, v) + _step_candidates(nm, v, nm.theta_low)
            memo.fill(cand + nxt)
    cost_at = memo

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
    memo = _Memo(nm, problem, x_0, u_array, kl_bound)
    if kl_bound > 0:
        if nm.c_high is None:
            while True:
                if nm.speculative and float(nm.theta_high) not in memo.cost:  # the next 8 halvings (and θ_low) in one launch
                    memo.fill([nm.theta_high * nm.lam ** j for j in range(8)] + ([nm.theta_low] if nm.c_low is None else []))
                nm.c_high = memo(nm.theta_high)
                if not math.isinf(nm.c_high):
                    break
                nm.theta_high *= nm.lam
                nm.theta_high_init *= nm.lam
        if nm.c_low is None:
            while True:
                if nm.speculative and float(nm.theta_low) not in memo.cost:
                    memo.fill([nm.theta_low * nm.lam ** j for j in range(8)])
                nm.c_low = memo(nm.theta_low)
                if not math.isinf(nm.c_low):
                    break
                nm.theta_low *= nm.lam
                nm.theta_low_init *= nm.lam
        while True:
            step_(nm, problem, x_0, u_array, kl_bound, verbose, memo)
            c_mean = (nm.c_low + nm.c_high) / 2
            stdev = math.sqrt(0.5 * ((nm.c_high - c_mean) ** 2 + (nm.c_low - c_mean) ** 2))
            if stdev < nm.eps:
                break
            if nm.iter_current == nm.iter_max:
                break
        theta_opt = nm.theta_low
    else:
        theta_opt = 0.0
    if kl_bound > 0 and float(theta_opt) in memo.sol:  # the final solve (:334-346) repeats a vertex evaluation: same inputs, same result
        xs, ls, Ls, value = memo.sol[float(theta_opt)]
        N = problem.N
        return (theta_opt, [xs[:, k].copy() for k in range(N + 1)], [ls[:, k].copy() for k in range(N)],
                [Ls[:, :, k].copy() for k in range(N)], value + kl_bound / theta_opt)
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
