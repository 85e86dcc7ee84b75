"""Receding-horizon RAT iLQR for a fleet of independent systems (the caller the reference does not
ship -- its `solve!` is a single planning step, F7).  Every control step plans all P problems at once on the device
(`ratilqr_ce_solve_fleet`), applies the first control of each plan to the true (noisy) system, shifts the plan and warm
starts the next step with it; the CE distribution parameters mu_init / sigma_init persist per problem exactly as the
reference persists them across `solve!` calls (cross_entropy_bilevel_optimization.jl:66-68, 297-301)."""
import time

import numpy as np

from .models import dynamics_numpy


def run_fleet_mpc(be, problem, cost_params, x0, steps, kl_bound=0.1, rng=None, noise_chol=None, num_samples=10, num_elite=3,
                  iter_max=5, mu_init=1.0, sigma_init=2.0, seed=0, opts=None):
    """x0 (n, P) -> dict(x (n, steps+1, P), u (m, steps, P), theta (steps, P), value (steps, P), ms (steps,)).
    The true system is the registered model plus additive noise w ~ N(0, noise_chol noise_chol') (default: the planner's W)."""
    spec = problem.spec(cost_params=cost_params)
    n, m, N = spec.n, spec.m, spec.N
    x0 = np.asarray(x0, dtype=np.float64).reshape(n, -1)
    P = x0.shape[1]
    rng = rng or np.random.default_rng(0)
    if noise_chol is None:
        noise_chol = np.linalg.cholesky(np.asarray(problem.W(0), float))
    xs = np.zeros((n, steps + 1, P)); xs[:, 0] = x0
    us = np.zeros((m, steps, P)); thetas = np.zeros((steps, P)); values = np.zeros((steps, P)); ms = np.zeros(steps)
    plan = np.zeros((m, N, P))                      # warm start: shifted previous plan
    mu_i = np.full(P, float(mu_init)); sg_i = np.full(P, float(sigma_init))
    model_id, mp = problem.f.model_id, problem.f.params
    for t in range(steps):
        t0 = time.perf_counter()
        r = be.ce_solve_fleet(spec, xs[:, t], plan, kl_bound, mu_i, sg_i, num_samples=num_samples, num_elite=num_elite,
                              iter_max=iter_max, seed=seed + t, opts=opts, want=("l",))
        ms[t] = (time.perf_counter() - t0) * 1e3
        mu_i, sg_i = r["mu_init"], r["sigma_init"]  # persisted CE state
        l = r["l"]                                  # (m, N, P): nominal controls of the new plans
        us[:, t] = l[:, 0]
        thetas[t], values[t] = r["theta_opt"], r["value"]
        w = noise_chol @ rng.standard_normal((n, P))
        for p in range(P):                          # true system step (host; the fleet sizes of interest plan on the GPU)
            xs[:, t + 1, p] = dynamics_numpy(model_id, mp, xs[:, t, p], us[:, t, p]) + w[:, p]
        plan = np.concatenate([l[:, 1:], l[:, -1:]], axis=1)  # shift, repeat the last control
    return dict(x=xs, u=us, theta=thetas, value=values, ms=ms, mu_init=mu_i, sigma_init=sg_i)
