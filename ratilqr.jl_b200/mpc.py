"""Receding-horizon RAT iLQR for a fleet of independent systems (the caller the reference does not
ship -- its `solve!` is a single planning step, F7).  Every control step plans all P problems at once on the device
(`ratilqr_ce_solve_fleet`), applies the first control of each plan to the true (noisy) system, shifts the plan and warm
starts the next step with it; the CE distribution parameters mu_init / sigma_init persist per problem exactly as the
reference persists them across `solve!` calls (cross_entropy_bilevel_optimization.jl:66-68, 297-301)."""
import time

import numpy as np

from ._capi import Spec


def _mixture_draws(mix, rng, n, P):
    """P draws of the true-model Gaussian mixture dict(weights, means (n, k), covs (n, n, k)) -> (n, P)"""
    w = np.asarray(mix["weights"], float)
    comp = np.minimum(np.searchsorted(np.cumsum(w / w.sum()), rng.random(P), side="right"), w.size - 1)
    means = np.asarray(mix["means"], float).reshape(n, -1)
    chols = np.stack([np.linalg.cholesky(np.asarray(mix["covs"], float).reshape(n, n, -1)[:, :, c]) for c in range(w.size)])
    z = rng.standard_normal((n, P))
    return means[:, comp] + np.einsum("pij,jp->ip", chols[comp], z)


def run_fleet_mpc(be, problem, cost_params, x0, steps, kl_bound=0.1, rng=None, noise_chol=None, num_samples=10, num_elite=3,
                  iter_max=5, mu_init=1.0, sigma_init=2.0, seed=0, opts=None, true_mixture=None):
    """x0 (n, P) -> dict(x (n, steps+1, P), u (m, steps, P), theta (steps, P), value (steps, P), ms (steps,)).
    The true system is the problem's device model (registered or user-supplied) plus additive noise: w ~ N(0, noise_chol
    noise_chol') (default: the planner's W), or draws of `true_mixture` (the accurate GMM the planner does not know,
    optimal_control_problems.jl:105-109).  The true step of all P systems is one 1-stage device rollout."""
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
    step_spec = Spec(spec.model_id, spec.cost_id, n, m, 1, spec.model_params, spec.cost_params.reshape(spec.cost_params_count, -1)[0],
                     np.eye(n))  # one stage of the true dynamics for the whole fleet (the cost block is not used by a rollout)
    for t in range(steps):
        t0 = time.perf_counter()
        r = be.ce_solve_fleet(spec, xs[:, t], plan, kl_bound, mu_i, sg_i, num_samples=num_samples, num_elite=num_elite,
                              iter_max=iter_max, seed=seed + t, opts=opts, want=("l",))
        ms[t] = (time.perf_counter() - t0) * 1e3
        mu_i, sg_i = r["mu_init"], r["sigma_init"]  # persisted CE state
        l = r["l"]                                  # (m, N, P): nominal controls of the new plans
        us[:, t] = l[:, 0]
        thetas[t], values[t] = r["theta_opt"], r["value"]
        w = _mixture_draws(true_mixture, rng, n, P) if true_mixture is not None else noise_chol @ rng.standard_normal((n, P))
        xn, st = be.rollout_open(step_spec, xs[:, t], us[:, t].reshape(m, 1, P))
        if np.any(st != 0):
            raise ArithmeticError(f"the true system left its domain for {int(np.sum(st != 0))} problems at step {t}")
        xs[:, t + 1] = xn[:, 1] + w
        plan = np.concatenate([l[:, 1:], l[:, -1:]], axis=1)  # shift, repeat the last control
    return dict(x=xs, u=us, theta=thetas, value=values, ms=ms, mu_init=mu_i, sigma_init=sg_i)


def run_fleet_mpc_on_device(be, problem, cost_params, x0, steps, kl_bound=0.1, u_init=None, noise=None, noise_seed=0,
                            true_mixture=None, num_samples=10, num_elite=3, iter_max=5, mu_init=1.0, sigma_init=2.0, seed=0,
                            z_inject=None, opts=None):
    """The same receding-horizon loop with EVERY step on the device (`ratilqr_mpc_fleet_run`): planning (CE loop + final
    solve), the true-system step x+ = f(x, l_0) + w, the disturbance draw (Philox N(0, W), the true-model mixture, or an
    injected tensor `noise` (n, steps, P)) and the plan shift; states, plans and noise never visit the host.
    Returns dict(x (n, steps+1, P), u (m, steps, P), theta (steps, P), value (steps, P), ms (steps,), mu_init, sigma_init)."""
    spec = problem.spec(cost_params=cost_params)
    if u_init is None:
        u_init = np.zeros((spec.m, spec.N))
    return be.mpc_fleet_run(spec, x0, u_init, steps, kl_bound, mu_init, sigma_init, num_samples=num_samples, num_elite=num_elite,
                            iter_max=iter_max, z_inject=z_inject, seed=seed, noise=noise, noise_seed=noise_seed,
                            true_mixture=true_mixture, opts=opts)
