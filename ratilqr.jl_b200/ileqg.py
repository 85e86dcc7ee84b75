"""Host mirror of src/ileqg.jl: same functions, same argument meaning, same error behaviour.

`solve_` (Julia `solve!`, ileqg.jl:635-659) runs the whole solve in ONE persistent CUDA kernel
through `ratilqr_ileqg_solve_batch`.  The finer-grained functions the reference exports
(`initialize!`, `approximate_model`, `solve_approximate_dp!`, `solve_approximate_dp`,
`line_search!`, `step!`, the rollouts) are mirrored on top of the component kernels so the
reference's own test-suite can be re-expressed line by line (tests/test_reference_ileqg.py).

Naming: Julia's `f!` becomes `f_`; Greek keyword arguments are accepted as written in the
reference (`μ_min`, `Δ_0`, `λ`, `ϵ_init`, `adaptive_ϵ_init`, `ϵ_min`) or in ASCII.
Trajectories are lists of 1-D arrays / 2-D arrays, like Julia's Vector{Vector} / Vector{Matrix}.
"""
import math
import unicodedata

import numpy as np

from . import _lib
from ._capi import make_opts
from .models import DomainError

_ALIASES = {"μ_min": "mu_min", "Δ_0": "delta_0", "λ": "lam", "lambda_": "lam", "ϵ_init": "eps_init",
            "ε_init": "eps_init", "adaptive_ϵ_init": "adaptive_eps_init", "adaptive_ε_init": "adaptive_eps_init",
            "ϵ_min": "eps_min", "ε_min": "eps_min", "θ": "theta", "μ": "mu"}
_ALIASES = {unicodedata.normalize("NFKC", k): v for k, v in _ALIASES.items()}


def _ascii_kw(kw):
    return {_ALIASES.get(unicodedata.normalize("NFKC", k), k): v for k, v in kw.items()}


class NotPositiveDefinite(AssertionError):
    """`@assert isposdef(M)` failed (ileqg.jl:366,440): neurotic breakdown."""


_STATUS_EXC = {1: NotPositiveDefinite, 2: NotPositiveDefinite, 3: DomainError, 4: RuntimeError, 5: RuntimeError}
_STATUS_MSG = {1: "M = inv(W) - θ*S is not PSD in initialize!", 2: "M = inv(W) - θ*S is not PSD",
               3: "DomainError in the model (negative base of a real power)",
               4: "line search cannot terminate (the reference loops forever here, ileqg.jl:526-535)",
               5: "regularisation μ overflowed"}


def raise_for_status(st):
    if st != 0:
        raise _STATUS_EXC.get(int(st), RuntimeError)(_STATUS_MSG.get(int(st), f"status {st}"))


def _cols(a):  # (d, T) array -> list of T vectors
    return [np.array(a[:, i]) for i in range(a.shape[1])]


def _mats(a):  # (r, c, T) -> list of T matrices
    return [np.array(a[:, :, i]) for i in range(a.shape[2])]


def _stack(v):  # list of vectors/matrices -> array with the list index last
    return np.stack([np.asarray(e, dtype=np.float64) for e in v], axis=-1)


def _W_arg(W_array):
    """W_array of an ApproximationResult -> one n x n matrix when W(k) is constant, else n x n x N (stage k uses W(k),
    ileqg.jl:364,438)."""
    W0 = W_array[0]
    if all(np.array_equal(W0, w) for w in W_array[1:]):
        return W0
    return _stack(W_array)


class ILEQGSolver:  # ileqg.jl:164-208
    def __init__(self, problem, backend=None, **kw):
        kw = _ascii_kw(kw)
        o = dict(mu_min=1e-6, delta_0=2.0, lam=0.5, d=1e-2, iter_max=100, eps_init=1.0,
                 adaptive_eps_init=False, eps_min=1e-6, f_returns_jacobian=False)
        unknown = set(kw) - set(o)
        if unknown:
            raise TypeError(f"unknown keyword arguments {sorted(unknown)}")
        o.update(kw)
        assert 0 < o["lam"] < 1, "λ has to be in (0, 1)"
        assert o["d"] > 0, "d > 0 is necessary"
        assert o["mu_min"] > 0, "μ_min > 0 is necessary"
        assert o["delta_0"] > 0, "Δ_0 > 0 is necessary"
        assert 0 < o["eps_init"] <= 1, "ϵ_init has to be in (0, 1]"
        assert o["eps_init"] > o["eps_min"], "ϵ_init > ϵ_min is necessary"
        assert 0 < o["eps_min"] < 1, "ϵ_min has to be in (0, 1)"
        self.mu_min, self.mu = o["mu_min"], o["mu_min"]
        self.delta_0, self.delta = o["delta_0"], o["delta_0"]
        self.lam, self.d, self.iter_max = o["lam"], o["d"], int(o["iter_max"])
        self.eps_init_auto, self.eps_init, self.eps_min = bool(o["adaptive_eps_init"]), o["eps_init"], o["eps_min"]
        self.f_returns_jacobian = bool(o["f_returns_jacobian"])
        self.x_array = [None] * (problem.N + 1)
        self.l_array = [None] * problem.N
        self.L_array = [None] * problem.N
        self.A_array = self.B_array = None
        self.value_current, self.iter_current, self.d_current = math.inf, 0, math.inf
        self.eps_history = []
        self.eps_init_init = o["eps_init"]
        self.backend = backend

    def opts(self):
        return make_opts(mu_min=self.mu_min, delta_0=self.delta_0, lam=self.lam, d=self.d, iter_max=self.iter_max,
                         adaptive_eps_init=self.eps_init_auto, eps_init=self.eps_init_init, eps_min=self.eps_min,
                         f_returns_jacobian=self.f_returns_jacobian)

    def _be(self):
        return self.backend or _lib.default_backend()


# ---- rollouts and cost (ileqg.jl:18-124) ---------------------------------------------------
def simulate_dynamics(problem, a, b, c=None, rng=None, f_returns_jacobian=False, backend=None):
    """The four `simulate_dynamics` overloads of ileqg.jl:18-109, selected like Julia's dispatch:
    (problem, x_0, u_array[, rng]) open loop; (problem, x_array, l_array, L_array[, rng]) closed loop.
    The noisy overloads draw w ~ N(0, W(k)) on the host and inject it (device: ratilqr_mc_rollout)."""
    be = backend or _lib.default_backend()
    spec = problem.spec()
    n, m, N = spec.n, spec.m, spec.N
    if c is not None and not isinstance(c, np.random.Generator):
        x_array, l_array, L_array = a, b, c
        assert N == len(l_array) and N == len(L_array)
        if rng is not None:
            noise = _draw_noise(problem, rng, n, N)
            r = be.mc_rollout(spec, _stack(x_array), _stack(l_array), _stack(L_array), 1, noise=noise, want_x=True)
            xs = r["x"][:, :, 0]
            us = [np.asarray(l_array[k]) + np.asarray(L_array[k]) @ (xs[:, k] - np.asarray(x_array[k])) for k in range(N)]
            return _cols(xs), us
        xn, un, st = be.rollout_closed(spec, _stack(x_array), _stack(l_array), _stack(L_array))
        raise_for_status(st[0])
        if f_returns_jacobian:
            lin = be.linearize(spec, xn, un)
            return _cols(xn[:, :, 0]), _cols(un[:, :, 0]), _mats(lin["A"][..., 0]), _mats(lin["B"][..., 0])
        return _cols(xn[:, :, 0]), _cols(un[:, :, 0])
    if isinstance(c, np.random.Generator):
        rng = c
    x_0, u_array = a, b
    assert N == len(u_array)
    if rng is not None:  # ileqg.jl:44-55: open loop + noise == closed loop around anything with L = 0
        noise = _draw_noise(problem, rng, n, N)
        xbar = np.zeros((n, N + 1))
        xbar[:, 0] = x_0
        r = be.mc_rollout(spec, xbar, _stack(u_array), np.zeros((m, n, N)), 1, noise=noise, want_x=True)
        return _cols(r["x"][:, :, 0])
    x, st = be.rollout_open(spec, np.asarray(x_0, float), _stack(u_array))
    raise_for_status(st[0])
    if f_returns_jacobian:
        lin = be.linearize(spec, x, _stack(u_array))
        return _cols(x[:, :, 0]), _mats(lin["A"][..., 0]), _mats(lin["B"][..., 0])
    return _cols(x[:, :, 0])


def _draw_noise(problem, rng, n, N):
    # rand(rng, MvNormal(zeros(n), W(k))) = chol_lower(W(k)) * randn(rng, n)   (ileqg.jl:51,104)
    w = np.zeros((n, N))
    for k in range(N):
        w[:, k] = np.linalg.cholesky(np.asarray(problem.W(k), float)) @ rng.standard_normal(n)
    return w


def integrate_cost(problem, x_array, u_array, backend=None):  # ileqg.jl:115-124
    be = backend or _lib.default_backend()
    assert problem.N == len(u_array) and problem.N + 1 == len(x_array)
    cost, st = be.integrate_cost(problem.spec(), _stack(x_array), _stack(u_array))
    raise_for_status(st[0])
    return float(cost[0])


# ---- approximation and DP (ileqg.jl:242-465) -------------------------------------------------
class ApproximationResult:  # ileqg.jl:242-252
    def __init__(self, lin, W_array):
        self._lin = lin
        self.q_array = list(lin["q"][:, 0])
        self.q_vec_array = _cols(lin["qv"][:, :, 0])
        self.Q_array = _mats(lin["Q"][..., 0])
        self.r_array = _cols(lin["r"][:, :, 0])
        self.R_array = _mats(lin["R"][..., 0])
        self.P_array = _mats(lin["P"][..., 0])
        self.A_array = _mats(lin["A"][..., 0])
        self.B_array = _mats(lin["B"][..., 0])
        self.W_array = W_array


class DynamicProgrammingResult:  # ileqg.jl:328-335
    def __init__(self, r):
        self.s_array = list(r["s"][:, 0])
        self.s_vec_array = _cols(r["sv"][:, :, 0])
        self.S_array = _mats(r["S"][..., 0])


def approximate_model(problem, u_array, x_array, A_array_input=None, B_array_input=None, backend=None):
    """ileqg.jl:258-322. User-supplied Jacobians are accepted and substituted (:302-311)."""
    be = backend or _lib.default_backend()
    assert problem.N == len(u_array)
    lin = be.linearize(problem.spec(), _stack(x_array), _stack(u_array))
    raise_for_status(lin["status"][0])
    if A_array_input is not None:
        assert len(A_array_input) == problem.N
        lin["A"][..., 0] = _stack(A_array_input)
    if B_array_input is not None:
        assert len(B_array_input) == problem.N
        lin["B"][..., 0] = _stack(B_array_input)
    return ApproximationResult(lin, [np.asarray(problem.W(k), float) for k in range(problem.N)])


def solve_approximate_dp_(ileqg, approx_result, verbose=False, backend=None, **kw):
    """solve_approximate_dp! (ileqg.jl:341-406): optimises L (stored into ileqg.L_array) and dl."""
    theta = _ascii_kw(kw)["theta"]
    be = backend or ileqg._be()
    r = be.riccati(approx_result._lin, _W_arg(approx_result.W_array), theta, True, mu=ileqg.mu, delta=ileqg.delta,
                   mu_min=ileqg.mu_min, delta_0=ileqg.delta_0)
    if r["status"][0] != 0:
        raise NotPositiveDefinite(_STATUS_MSG[2])
    ileqg.mu, ileqg.delta = float(r["mu"][0]), float(r["delta"][0])
    ileqg.L_array = _mats(r["L"][..., 0])
    return DynamicProgrammingResult(r), _cols(r["dl"][:, :, 0])


def solve_approximate_dp(approx_result, L_array, dl_array=None, backend=None, **kw):
    """solve_approximate_dp (ileqg.jl:412-465): evaluates the policy (L, dl) at fixed μ."""
    kw = _ascii_kw(kw)
    be = backend or _lib.default_backend()
    N = len(approx_result.W_array)
    assert N == len(L_array)
    if dl_array is not None:
        assert N == len(dl_array)
    r = be.riccati(approx_result._lin, _W_arg(approx_result.W_array), kw["theta"], False,
                   L=_stack(L_array)[..., None], dl=None if dl_array is None else _stack(dl_array)[..., None],
                   mu=kw["mu"])
    if r["status"][0] != 0:
        raise NotPositiveDefinite(_STATUS_MSG[2])
    return DynamicProgrammingResult(r)


def increase_mu_and_delta_(ileqg):  # increase_μ_and_Δ!  ileqg.jl:471-474
    ileqg.delta = max(ileqg.delta_0, ileqg.delta * ileqg.delta_0)
    ileqg.mu = max(ileqg.mu_min, ileqg.mu * ileqg.delta)


def decrease_mu_and_delta_(ileqg):  # decrease_μ_and_Δ!  ileqg.jl:480-488
    ileqg.delta = min(1 / ileqg.delta_0, ileqg.delta / ileqg.delta_0)
    cand = ileqg.mu * ileqg.delta
    ileqg.mu = cand if cand >= ileqg.mu_min else 0.0


def initialize_(ileqg, problem, x_0, u_array, theta, backend=None):
    """initialize! (ileqg.jl:214-236)."""
    be = backend or ileqg._be()
    ileqg.mu, ileqg.delta = 0.0, ileqg.delta_0
    ileqg.d_current = math.inf
    ileqg.iter_current = 0
    ileqg.eps_init = ileqg.eps_init_init
    ileqg.eps_history = []
    if ileqg.f_returns_jacobian:
        ileqg.x_array, ileqg.A_array, ileqg.B_array = simulate_dynamics(problem, x_0, u_array,
                                                                        f_returns_jacobian=True, backend=be)
    else:
        ileqg.x_array = simulate_dynamics(problem, x_0, u_array, backend=be)
        ileqg.A_array = ileqg.B_array = None
    ileqg.l_array = [np.array(u, dtype=np.float64) for u in u_array]
    n, m = len(ileqg.x_array[0]), len(ileqg.l_array[0])
    ileqg.L_array = [np.zeros((m, n)) for _ in range(problem.N)]
    approx = approximate_model(problem, ileqg.l_array, ileqg.x_array, backend=be)
    try:
        dp = solve_approximate_dp(approx, ileqg.L_array, theta=theta, mu=ileqg.mu, backend=be)
    except NotPositiveDefinite:
        raise NotPositiveDefinite(_STATUS_MSG[1])
    ileqg.value_current = dp.s_array[0]


def _isapprox(a, b):  # Base.isapprox defaults
    if a == b:
        return True
    if not (math.isfinite(a) and math.isfinite(b)):
        return False
    return abs(a - b) <= math.sqrt(np.finfo(float).eps) * max(abs(a), abs(b))


def line_search_(ileqg, problem, dl_array_new, theta, verbose=False, backend=None):
    """line_search! (ileqg.jl:494-592), host loop over the component kernels."""
    be = backend or ileqg._be()
    cur = ileqg.value_current
    eps = ileqg.eps_init
    count = 0
    while True:
        count += 1
        if eps == 0.0:
            raise RuntimeError(_STATUS_MSG[4])
        l_new = [l + eps * dl for l, dl in zip(ileqg.l_array, dl_array_new)]
        x_new, u_new = simulate_dynamics(problem, ileqg.x_array, l_new, ileqg.L_array, backend=be)
        approx_new = approximate_model(problem, u_new, x_new, backend=be)
        try:
            dp_new = solve_approximate_dp(approx_new, ileqg.L_array, theta=theta, mu=ileqg.mu, backend=be)
        except NotPositiveDefinite:
            eps *= ileqg.lam
            continue
        new = dp_new.s_array[0]
        ileqg.eps_history.append((eps, new - cur))

        def accept():
            ileqg.d_current = max(float(np.sqrt(np.sum((l - u) ** 2))) for l, u in zip(ileqg.l_array, u_new))
            ileqg.value_current, ileqg.x_array, ileqg.l_array = new, x_new, u_new
            if ileqg.f_returns_jacobian:
                ileqg.A_array, ileqg.B_array = approx_new.A_array, approx_new.B_array
            else:
                ileqg.A_array = ileqg.B_array = None

        if _isapprox(new, cur) or new < cur:
            accept()
            break
        eps *= ileqg.lam
        if eps < ileqg.eps_min:
            accept()
            break
    if ileqg.eps_init_auto:
        if count == 1:
            ileqg.eps_init = min(ileqg.eps_init_init, eps / ileqg.lam)
        else:
            while eps < ileqg.eps_min:
                eps = eps / ileqg.lam
            ileqg.eps_init = eps


def step_(ileqg, problem, theta, verbose=False, backend=None):  # step! ileqg.jl:598-613
    be = backend or ileqg._be()
    ileqg.iter_current += 1
    approx = approximate_model(problem, ileqg.l_array, ileqg.x_array, ileqg.A_array, ileqg.B_array, backend=be)
    _, dl_array = solve_approximate_dp_(ileqg, approx, verbose, theta=theta, backend=be)
    line_search_(ileqg, problem, dl_array, theta, verbose, backend=be)


def solve_(ileqg, problem, x_0, u_array, verbose=False, backend=None, **kw):
    """solve! (ileqg.jl:635-659): one persistent-kernel launch. Returns
    (x_array, l_array, L_array, value, ϵ_history); raises where the reference throws."""
    theta = float(_ascii_kw(kw)["theta"])
    be = backend or ileqg._be()
    spec = problem.spec()
    # ϵ_history holds one entry per merit evaluation: up to ceil(log(ϵ_min/ϵ_init)/log(λ)) + 1 trials per iteration (:537)
    per_iter = int(math.ceil(math.log(ileqg.eps_min / ileqg.eps_init_init) / math.log(ileqg.lam))) + 2
    cap = max(16, ileqg.iter_max * per_iter)
    r = be.ileqg_solve_batch(spec, np.asarray(x_0, float), _stack(u_array), [theta], opts=ileqg.opts(),
                             eps_hist_cap=cap)
    raise_for_status(r["status"][0])
    ileqg.x_array, ileqg.l_array = _cols(r["x"][:, :, 0]), _cols(r["l"][:, :, 0])
    ileqg.L_array = _mats(r["L"][..., 0])
    ileqg.value_current = float(r["value"][0])
    ileqg.iter_current, ileqg.d_current, ileqg.mu = int(r["iters"][0]), float(r["d_current"][0]), float(r["mu"][0])
    nt = int(r["trials"][0])
    if nt > cap:
        raise RuntimeError(f"ϵ_history overflow: {nt} line-search trials > capacity {cap}")
    ileqg.eps_history = [(float(r["eps_hist"][0, i, 0]), float(r["eps_hist"][1, i, 0])) for i in range(nt)]
    if verbose:
        print(f"ILEQG finished: iterations {ileqg.iter_current}, value {ileqg.value_current:.6g}, d == {ileqg.d_current:.3g}")
    return (list(ileqg.x_array), list(ileqg.l_array), list(ileqg.L_array), ileqg.value_current, list(ileqg.eps_history))
