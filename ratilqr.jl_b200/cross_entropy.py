"""Host mirror of src/cross_entropy_bilevel_optimization.jl (RAT iLQR: CEM over θ).

The θ fan-out -- `compute_cost` (:173-195), which the reference spreads over Distributed worker
processes -- is one batched CUDA launch (`ratilqr_ce_costs`).  The tiny feasibility / redraw
state machine of `step!` (:252-335) stays on the host, exactly as it stays on the Julia master.
`rng` is anything with `.standard_normal()` (numpy Generator, or InjectedNormals for exactness
tests: Julia's MersenneTwister stream cannot be reproduced).
"""
import math

import numpy as np

from . import _lib
from .ileqg import ILEQGSolver, _stack
from .ileqg import solve_ as ileqg_solve_


class InjectedNormals:
    """A fixed stream of standard-normal draws (replaces `rng` for bit-reproducible tests)."""

    def __init__(self, z):
        self.z, self.i = np.asarray(z, dtype=np.float64), 0

    def standard_normal(self, size=None):
        k = 1 if size is None else int(np.prod(size))
        if self.i + k > self.z.size:
            raise IndexError("injected normal stream exhausted")
        out = self.z[self.i:self.i + k]
        self.i += k
        return float(out[0]) if size is None else out.reshape(size)


class CrossEntropyBilevelOptimizationSolver:  # :70-127
    def __init__(self, backend=None, **kw):
        kw = {k.replace("ϵ", "eps").replace("ε", "eps"): v for k, v in kw.items()}
        al = {"μ_min_ileqg": "mu_min_ileqg", "Δ_0_ileqg": "delta_0_ileqg", "λ_ileqg": "lam_ileqg",
              "adaptive_eps_init_ileqg": "adaptive_eps_init_ileqg", "μ_init": "mu_init", "σ_init": "sigma_init",
              "λ": "lam", "use_θ_max": "use_theta_max"}
        import unicodedata
        al = {unicodedata.normalize("NFKC", k): v for k, v in al.items()}
        kw = {al.get(unicodedata.normalize("NFKC", k), k): v for k, v in kw.items()}
        o = dict(mu_min_ileqg=1e-6, delta_0_ileqg=2.0, lam_ileqg=0.5, d_ileqg=1e-2, iter_max_ileqg=100,
                 adaptive_eps_init_ileqg=False, eps_init_ileqg=1.0, eps_min_ileqg=1e-6, mu_init=1.0, sigma_init=2.0,
                 num_samples=10, num_elite=3, iter_max=5, lam=0.5, f_returns_jacobian=False, use_theta_max=False)
        unknown = set(kw) - set(o)
        if unknown:
            raise TypeError(f"unknown keyword arguments {sorted(unknown)}")
        o.update(kw)
        self.__dict__.update(o)
        self.mu, self.sigma = self.mu_init, self.sigma_init
        self.theta_max, self.theta_min = 0.0, math.inf
        self.iter_current = 0
        self.backend = backend

    def _be(self):
        return self.backend or _lib.default_backend()

    def ileqg_kwargs(self):
        return dict(mu_min=self.mu_min_ileqg, delta_0=self.delta_0_ileqg, lam=self.lam_ileqg, d=self.d_ileqg,
                    iter_max=self.iter_max_ileqg, adaptive_eps_init=self.adaptive_eps_init_ileqg,
                    eps_init=self.eps_init_ileqg, eps_min=self.eps_min_ileqg,
                    f_returns_jacobian=self.f_returns_jacobian)


def initialize_(ce_solver):  # :133-138
    ce_solver.iter_current = 0
    ce_solver.mu, ce_solver.sigma = ce_solver.mu_init, ce_solver.sigma_init
    ce_solver.theta_max = 0.0
    ce_solver.theta_min = math.inf


def compute_cost(ce_solver, problem, x, u_array, theta_array, kl_bound):
    """compute_cost (:173-195): every θ sample solved concurrently on the GPU."""
    ileqg = ILEQGSolver(problem, **ce_solver.ileqg_kwargs())  # runs the constructor's @asserts
    cost, _ = ce_solver._be().ce_costs(problem.spec(), np.asarray(x, float), _stack(u_array),
                                       np.asarray(theta_array, float), kl_bound, opts=ileqg.opts())
    return cost


def compute_cost_serial(ce_solver, problem, x, u_array, theta_array, kl_bound):
    """compute_cost_serial (:198-227). Same device path; keeps the reference's length assertion."""
    assert len(theta_array) == ce_solver.num_samples
    return compute_cost(ce_solver, problem, x, u_array, theta_array, kl_bound)


def get_positive_samples(mu, sigma, num_samples, rng):  # :233-246
    out = []
    while True:
        t = mu + sigma * rng.standard_normal()
        if t > 0.0:
            out.append(t)
        if len(out) >= num_samples:
            break
    return np.array(out)


def step_(ce, problem, x, u_array, kl_bound, rng, verbose=False, serial=False):  # step! :252-335
    ce.iter_current += 1
    while True:
        if ce.iter_current == 1:
            theta_array = get_positive_samples(ce.mu_init, ce.sigma_init, ce.num_samples, rng)
        else:
            theta_array = get_positive_samples(ce.mu, ce.sigma, ce.num_samples, rng)
        fn = compute_cost_serial if serial else compute_cost
        costs = fn(ce, problem, x, u_array, theta_array, kl_bound)
        if verbose:
            print(costs)
        num_inf = int(np.sum(np.isinf(costs)))
        num_valid = ce.num_samples - num_inf
        thr = max(ce.num_elite, ce.num_samples * ce.lam)
        if ce.iter_current == 1 and num_valid < thr:
            ce.mu_init *= ce.lam
            ce.sigma_init *= ce.lam
        elif ce.iter_current == 1 and num_valid == ce.num_samples:
            ce.mu_init /= ce.lam
            ce.sigma_init /= ce.lam
            break
        elif num_valid >= thr:
            break
    for t, c in zip(theta_array, costs):  # :314-324 (if / elseif quirk kept)
        if math.isinf(c):
            continue
        if t < ce.theta_min:
            ce.theta_min = float(t)
        elif t > ce.theta_max:
            ce.theta_max = float(t)
    order = sorted(range(len(costs)), key=lambda i: (math.isnan(costs[i]), costs[i]))  # stable, NaN last
    elite = np.array([theta_array[i] for i in order[:ce.num_elite]])
    mu_new = float(np.sum(elite) / ce.num_elite)
    sigma_new = float(np.sqrt(np.sum((elite - mu_new) ** 2) / ce.num_elite))
    ce.mu, ce.sigma = mu_new, sigma_new
    ce.last_theta_array, ce.last_costs = theta_array, costs


def solve_(ce, problem, x_0, u_array, rng, verbose=False, serial=False, **kw):
    """solve! (:364-415) -> (θ_opt, x_array, l_array, L_array, value, θ_min, θ_max)."""
    kl_bound = float(kw["kl_bound"])
    assert kl_bound >= 0, "KL Divergence Bound must be non-negative"
    initialize_(ce)
    theta_min = theta_max = 0.0
    if kl_bound > 0:
        while ce.iter_current < ce.iter_max:
            step_(ce, problem, x_0, u_array, kl_bound, rng, verbose, serial)
        theta_min, theta_max = ce.theta_min, ce.theta_max
        theta_opt = theta_max if ce.use_theta_max else ce.mu
    else:
        theta_opt = 0.0
    guard = 0
    while True:
        try:
            ileqg = ILEQGSolver(problem, backend=ce._be(), **ce.ileqg_kwargs())
            x_array, l_array, L_array, value, _ = ileqg_solve_(ileqg, problem, x_0, u_array, theta=theta_opt, verbose=False)
            if kl_bound > 0:
                # Julia: kl_bound/0.0 == Inf (the retry rule can drive θ_opt to 0.0, :410-413); Python would raise
                kl_term = kl_bound / theta_opt if theta_opt > 0 else math.inf
                return theta_opt, x_array, l_array, L_array, value + kl_term, theta_min, theta_max
            return theta_opt, x_array, l_array, L_array, value, 0.0, 0.0
        except (AssertionError, ValueError, RuntimeError):
            if verbose:
                print(f"θ_opt == {theta_opt} resulted in neurotic breakdown. Re-trying with {max(0.0, theta_opt - ce.sigma)}")
            theta_opt = max(0.0, theta_opt - ce.sigma)
            guard += 1
            if guard > 10000:
                raise


def solve_on_device_(ce, problem, x_0, u_array, rng, verbose=False, **kw):
    """solve! (:364-415) with the WHOLE CE loop on the device (ratilqr_ce_solve): draws, batched iLEQG solves, feasibility /
    redraw logic, elite selection and refit, final solve with the retry rule.  The θ draws are the caller's rng stream:
    standard normals are taken from a copy of `rng` and injected (rand(rng, Normal(μ, σ)) = μ + σ·randn(rng), :235-237), then
    `rng` is advanced by the number the solve consumed -- the same samples, and the same generator state afterwards, as the
    host loop solve_().  Returns the reference's 7-tuple and persists μ_init / σ_init in `ce` like the reference (:66-68)."""
    import copy
    kl_bound = float(kw["kl_bound"])
    assert kl_bound >= 0, "KL Divergence Bound must be non-negative"
    initialize_(ce)
    spec = problem.spec()
    opts = ILEQGSolver(problem, **ce.ileqg_kwargs()).opts()
    nz = max(4096, 64 * ce.num_samples * ce.iter_max)
    while True:
        z = copy.deepcopy(rng).standard_normal(nz)
        try:
            r = ce._be().ce_solve(spec, np.asarray(x_0, float), np.stack(u_array, axis=-1), kl_bound, ce.mu_init, ce.sigma_init,
                                  num_samples=ce.num_samples, num_elite=ce.num_elite, iter_max=ce.iter_max, lam=ce.lam,
                                  use_theta_max=ce.use_theta_max, z_inject=z, opts=opts)
            break
        except Exception as e:  # injected stream exhausted by a long redraw phase: retry with a longer one
            if "exhausted" not in str(e) or nz >= 1 << 24:
                raise
            nz *= 4
    if r["nz_used"]:
        rng.standard_normal(r["nz_used"])
    ce.mu_init, ce.sigma_init, ce.mu, ce.sigma = r["mu_init"], r["sigma_init"], r["mu"], r["sigma"]
    if kl_bound > 0:
        ce.theta_min, ce.theta_max, ce.iter_current = r["theta_min"], r["theta_max"], ce.iter_max
    N = problem.N
    xa = [r["x"][:, k].copy() for k in range(N + 1)]
    la = [r["l"][:, k].copy() for k in range(N)]
    La = [r["L"][:, :, k].copy() for k in range(N)]
    return r["theta_opt"], xa, la, La, r["value"], r["theta_min"], r["theta_max"]


def solve_fleet_(ce, problem, x0, u_init, kl_bound, cost_params=None, rng_seed=0, z_inject=None, want=("x", "l", "L")):
    """solve! for a FLEET of independent problems in one call (additive API): x0 (n, P), per-problem cost parameter blocks
    (P, ncp).  The whole CE loop runs on the device (ratilqr_ce_solve_fleet).  Returns a dict of per-problem arrays and
    updates nothing on `ce` except reading its options (mu_init / sigma_init come back in the dict, per problem)."""
    ILEQGSolver(problem, **ce.ileqg_kwargs())  # constructor assertions
    spec = problem.spec(cost_params=cost_params)
    return ce._be().ce_solve_fleet(spec, x0, u_init, kl_bound, ce.mu_init, ce.sigma_init, num_samples=ce.num_samples,
                                   num_elite=ce.num_elite, iter_max=ce.iter_max, lam=ce.lam, use_theta_max=ce.use_theta_max,
                                   z_inject=z_inject, seed=rng_seed, opts=ILEQGSolver(problem, **ce.ileqg_kwargs()).opts(), want=want)
