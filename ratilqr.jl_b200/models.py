"""Registered device models and cost families (host-side callables).

The reference problem structs hold arbitrary Julia closures `f, c, h, W`
(optimal_control_problems.jl:67-73).  Closures cannot run on the GPU, so the accelerated
path covers a registered set.  Each object below is a *callable* with the
reference's call signature -- `f(x, u, f_returns_jacobian=false)`, `c(k, x, u)`, `h(x)`,
`W(k)` -- so user code written against the reference keeps working on the host, and it
carries `(model_id, params)` so the solvers can dispatch to the CUDA library.

The numpy bodies here are a third, independent statement of the registry equations
(DESIGN.md "Model registry"); tests differentiate them by complex step to validate the
analytic device Jacobians.
"""
import numpy as np

MODEL_SINGLE_INTEGRATOR = 1
MODEL_POWER_LAW = 2
MODEL_DOUBLE_INTEGRATOR = 3
MODEL_PENDULUM = 4
MODEL_CARTPOLE = 5
MODEL_UNICYCLE = 6
MODEL_QUADROTOR = 7

COST_QUADRATIC = 1
COST_POWER_LAW = 2
COST_L1_CONTROL = 3

_DIMS = {  # model_id -> (n, m, n_params, default params)
    MODEL_SINGLE_INTEGRATOR: (2, 2, 1, [1.0]),
    MODEL_POWER_LAW: (2, 2, 2, [1.3, 1.5]),
    MODEL_DOUBLE_INTEGRATOR: (4, 2, 1, [0.1]),
    MODEL_PENDULUM: (2, 1, 5, [0.05, 9.81, 1.0, 1.0, 0.1]),
    MODEL_CARTPOLE: (4, 1, 5, [0.02, 1.0, 0.1, 0.5, 9.81]),
    MODEL_UNICYCLE: (4, 2, 1, [0.1]),
    MODEL_QUADROTOR: (12, 4, 6, [0.05, 1.0, 9.81, 0.01, 0.01, 0.02]),
}


class DomainError(ValueError):
    """Julia's DomainError: negative base of a real power."""


def _rpow(x, p):
    if np.isrealobj(x) and np.any(np.asarray(x) < 0):
        raise DomainError("negative base of a real power")
    return x ** p


def dynamics_numpy(model_id, p, x, u):
    """x_next = f(x, u); generic in the element type so complex-step differentiation works."""
    x = np.asarray(x)
    u = np.asarray(u)
    dt = p[0]
    if model_id == MODEL_SINGLE_INTEGRATOR:
        return x + dt * u
    if model_id == MODEL_POWER_LAW:
        return _rpow(x, p[0]) + _rpow(u, p[1])
    if model_id == MODEL_DOUBLE_INTEGRATOR:
        return np.array([x[0] + dt * x[2], x[1] + dt * x[3], x[2] + dt * u[0], x[3] + dt * u[1]])
    if model_id == MODEL_PENDULUM:
        _, g, ln, mass, damp = p
        inertia = mass * ln * ln
        alpha = (u[0] - damp * x[1] - (mass * g * ln) * np.sin(x[0])) / inertia
        return np.array([x[0] + dt * x[1], x[1] + dt * alpha])
    if model_id == MODEL_CARTPOLE:
        _, mc, mp, ln, g = p
        s, c = np.sin(x[1]), np.cos(x[1])
        den = mc + mp * (s * s)
        thd2 = x[3] * x[3]
        acc = (u[0] + mp * s * (ln * thd2 + g * c)) / den
        thacc = (-(u[0] * c) - (mp * ln) * thd2 * c * s - ((mc + mp) * g) * s) / (ln * den)
        return np.array([x[0] + dt * x[2], x[1] + dt * x[3], x[2] + dt * acc, x[3] + dt * thacc])
    if model_id == MODEL_UNICYCLE:
        s, c = np.sin(x[2]), np.cos(x[2])
        return np.array([x[0] + dt * (x[3] * c), x[1] + dt * (x[3] * s), x[2] + dt * u[1], x[3] + dt * u[0]])
    if model_id == MODEL_QUADROTOR:
        _, mass, g, Ix, Iy, Iz = p
        sph, cph = np.sin(x[3]), np.cos(x[3])
        sth, cth = np.sin(x[4]), np.cos(x[4])
        sps, cps = np.sin(x[5]), np.cos(x[5])
        tth = sth / cth
        wp, wq, wr = x[9], x[10], x[11]
        qr = wq * sph + wr * cph
        dphi = wp + qr * tth
        dth = wq * cph - wr * sph
        dpsi = qr / cth
        tm = u[0] / mass
        ax = tm * (cph * sth * cps + sph * sps)
        ay = tm * (cph * sth * sps - sph * cps)
        az = tm * (cph * cth) - g
        dwp = (u[1] + (Iy - Iz) * (wq * wr)) / Ix
        dwq = (u[2] + (Iz - Ix) * (wp * wr)) / Iy
        dwr = (u[3] + (Ix - Iy) * (wp * wq)) / Iz
        d = [x[6], x[7], x[8], dphi, dth, dpsi, ax, ay, az, dwp, dwq, dwr]
        return np.array([x[i] + dt * d[i] for i in range(12)])
    raise ValueError(f"unknown model id {model_id}")


def jacobians_complex_step(model_id, p, x, u, h=1e-30):
    """df/dx, df/du by complex step (exact to rounding for analytic f)."""
    x = np.asarray(x, dtype=float)
    u = np.asarray(u, dtype=float)
    n, m = x.size, u.size
    A = np.zeros((n, n))
    B = np.zeros((n, m))
    for j in range(n):
        xc = x.astype(complex)
        xc[j] += 1j * h
        A[:, j] = np.imag(dynamics_numpy(model_id, p, xc, u.astype(complex))) / h
    for j in range(m):
        uc = u.astype(complex)
        uc[j] += 1j * h
        B[:, j] = np.imag(dynamics_numpy(model_id, p, x.astype(complex), uc)) / h
    return A, B


class DeviceDynamics:
    """`f(x, u, f_returns_jacobian=false)` of a registered model (callable, like a Julia Function)."""

    def __init__(self, model_id, params=None):
        n, m, npar, default = _DIMS[model_id]
        self.model_id, self.n, self.m = model_id, n, m
        self.params = np.asarray(default if params is None else params, dtype=np.float64)
        assert self.params.size == npar, f"model {model_id} takes {npar} parameters"

    def __call__(self, x, u, f_returns_jacobian=False):
        mid = getattr(self, "host_model_id", self.model_id)  # (a registered model bound to a user cost keeps its own id here)
        xn = dynamics_numpy(mid, self.params, np.asarray(x, float), np.asarray(u, float))
        if f_returns_jacobian:
            A, B = jacobians_complex_step(mid, self.params, x, u)
            return xn, A, B
        return xn


def SingleIntegrator(dt=1.0):
    return DeviceDynamics(MODEL_SINGLE_INTEGRATOR, [dt])


def PowerLawDynamics(a=1.3, b=1.5):
    return DeviceDynamics(MODEL_POWER_LAW, [a, b])


def DoubleIntegrator(dt=0.1):
    return DeviceDynamics(MODEL_DOUBLE_INTEGRATOR, [dt])


def Pendulum(dt=0.05, g=9.81, length=1.0, mass=1.0, damping=0.1):
    return DeviceDynamics(MODEL_PENDULUM, [dt, g, length, mass, damping])


def CartPole(dt=0.02, m_cart=1.0, m_pole=0.1, length=0.5, g=9.81):
    return DeviceDynamics(MODEL_CARTPOLE, [dt, m_cart, m_pole, length, g])


def Unicycle(dt=0.1):
    return DeviceDynamics(MODEL_UNICYCLE, [dt])


def Quadrotor(dt=0.05, mass=1.0, g=9.81, Ixx=0.01, Iyy=0.01, Izz=0.02):
    return DeviceDynamics(MODEL_QUADROTOR, [dt, mass, g, Ixx, Iyy, Izz])


# ----------------------------------------------------------------------------------------
# costs
# ----------------------------------------------------------------------------------------
class _StageCost:
    def __init__(self, cost):
        self.cost = cost

    def __call__(self, k, x, u):
        return self.cost.stage(k, np.asarray(x, float), np.asarray(u, float))


class _TerminalCost:
    def __init__(self, cost):
        self.cost = cost

    def __call__(self, x):
        return self.cost.terminal(np.asarray(x, float))


class DeviceCost:
    """A registered cost family; `.c` and `.h` are the callables to put in the problem struct."""

    cost_id = 0

    def __init__(self):
        self.c = _StageCost(self)
        self.h = _TerminalCost(self)


class QuadraticCost(DeviceCost):
    """c(k,x,u) = (ws0+ws1 k)(1/2 dx'Q dx + 1/2 u'R u + dx'Pc u) + c0 + c1 k ; h(x) = 1/2 dx'Qf dx + h0.

    Covers the docs example `k/2 x'x + k/2 u'u`, `N/2 x'x` (optimal_control_problems.jl:59-61),
    `c = k`, `h = 1` (test/ileqg_test.jl:13-14) and `1/2 x'x + u'u + x'u` (:53).
    """

    cost_id = COST_QUADRATIC

    def __init__(self, n, m, Q=None, R=None, Qf=None, xg=None, Pc=None, ws0=1.0, ws1=0.0, c0=0.0, c1=0.0, h0=0.0):
        super().__init__()
        self.n, self.m = n, m
        z = np.zeros
        self.Q = z((n, n)) if Q is None else np.asarray(Q, float).reshape(n, n)
        self.R = z((m, m)) if R is None else np.asarray(R, float).reshape(m, m)
        self.Qf = z((n, n)) if Qf is None else np.asarray(Qf, float).reshape(n, n)
        self.Pc = z((n, m)) if Pc is None else np.asarray(Pc, float).reshape(n, m)
        self.xg = z(n) if xg is None else np.asarray(xg, float).reshape(n)
        for M in (self.Q, self.R, self.Qf):
            assert np.array_equal(M, M.T), "Q, R, Qf must be symmetric"
        self.ws0, self.ws1, self.c0, self.c1, self.h0 = float(ws0), float(ws1), float(c0), float(c1), float(h0)

    def stage(self, k, x, u):
        dx = x - self.xg
        w = self.ws0 + self.ws1 * k
        return w * (0.5 * dx @ self.Q @ dx + 0.5 * u @ self.R @ u + dx @ self.Pc @ u) + self.c0 + self.c1 * k

    def terminal(self, x):
        dx = x - self.xg
        return 0.5 * dx @ self.Qf @ dx + self.h0

    def params(self, xg=None):
        xg = self.xg if xg is None else np.asarray(xg, float)
        return np.concatenate([[self.ws0, self.ws1, self.c0, self.c1, self.h0], xg,
                               self.Q.ravel(order="F"), self.R.ravel(order="F"),
                               self.Pc.ravel(order="F"), self.Qf.ravel(order="F")])


class PowerLawCost(DeviceCost):
    """c = sum(x.^p + u.^p), h = h0 (test/ileqg_test.jl:152-153)."""

    cost_id = COST_POWER_LAW

    def __init__(self, p=2.5, h0=1.0):
        super().__init__()
        self.p, self.h0 = float(p), float(h0)

    def stage(self, k, x, u):
        return float(np.sum(_rpow(x, self.p) + _rpow(u, self.p)))

    def terminal(self, x):
        return self.h0

    def params(self):
        return np.array([self.p, self.h0])


class L1ControlCost(DeviceCost):
    """c = sum(abs.(u)), h = h0 (test/pets_test.jl:16-17). Rollout-only (PETS)."""

    cost_id = COST_L1_CONTROL

    def __init__(self, h0=1.0):
        super().__init__()
        self.h0 = float(h0)

    def stage(self, k, x, u):
        return float(np.sum(np.abs(u)))

    def terminal(self, x):
        return self.h0

    def params(self):
        return np.array([self.h0])


# ----------------------------------------------------------------------------------------
# user-extensible device models: CUDA C++ snippets compiled at run time (NVRTC)
# ----------------------------------------------------------------------------------------
class UserDynamics(DeviceDynamics):
    """Dynamics given as a CUDA C++ snippet  `template <class T> void dynamics(const double* p, const T* x,
    const T* u, T* xn)`  (include/ratilqr.h, "user-extensible device models").  `py` is an optional host callable
    `py(p, x, u) -> xn` so that the object is still callable on the CPU like a Julia closure; the solvers never use it."""

    def __init__(self, n, m, src, params=(), py=None, a_kind=None, b_kind=None):
        self.model_id, self.n, self.m, self.src, self.py = None, int(n), int(m), src, py
        self.a_kind, self.b_kind = a_kind, b_kind  # optional declared structure of df/dx (n, n), df/du (n, m): 0 / 1 / 2
        self.params = np.asarray(params, dtype=np.float64).reshape(-1)
        assert self.params.size <= 8, "user dynamics take at most 8 parameters"

    def __call__(self, x, u, f_returns_jacobian=False):
        if self.py is None:
            raise TypeError("this UserDynamics has no host callable (pass py=...)")
        x, u = np.asarray(x, float), np.asarray(u, float)
        xn = np.asarray(self.py(self.params, x, u), float)
        if not f_returns_jacobian:
            return xn
        h, n, m = 1e-30, self.n, self.m  # complex-step Jacobians, like the registered models
        A, B = np.zeros((n, n)), np.zeros((n, m))
        for j in range(n):
            xc = x.astype(complex); xc[j] += 1j * h
            A[:, j] = np.imag(self.py(self.params, xc, u.astype(complex))) / h
        for j in range(m):
            uc = u.astype(complex); uc[j] += 1j * h
            B[:, j] = np.imag(self.py(self.params, x.astype(complex), uc)) / h
        return xn, A, B


class UserCost(DeviceCost):
    """Cost given as a CUDA C++ snippet defining `stage_cost<T>(cp, k, x, u)` and `terminal_cost<T>(cp, x)`;
    `cp` is the parameter vector.  `stage_py(cp, k, x, u)` / `terminal_py(cp, x)` are optional host callables."""

    cost_id = 100  # RATILQR_COST_USER

    def __init__(self, src, params=(), stage_py=None, terminal_py=None, q_kind=None, r_kind=None, p_kind=None):
        super().__init__()
        self.src, self.stage_py, self.terminal_py = src, stage_py, terminal_py
        self.q_kind, self.r_kind, self.p_kind = q_kind, r_kind, p_kind  # optional structure of cxx (n, n), cuu (m, m), cux (m, n): 0 / 2
        self._params = np.asarray(params, dtype=np.float64).reshape(-1)

    def stage(self, k, x, u):
        if self.stage_py is None:
            raise TypeError("this UserCost has no host callable (pass stage_py=...)")
        return float(self.stage_py(self._params, k, x, u))

    def terminal(self, x):
        if self.terminal_py is None:
            raise TypeError("this UserCost has no host callable (pass terminal_py=...)")
        return float(self.terminal_py(self._params, x))

    def params(self):
        return self._params.copy()


def register_user_model(backend, dynamics, cost):
    """Compile + load a (dynamics, cost) pair of which at least one is user-supplied; returns the dynamics object
    to put into the problem struct (its model_id now names the compiled pair in `backend`'s context; use it together
    with `cost.c`, `cost.h`).  `dynamics`: UserDynamics or a registered DeviceDynamics; `cost`: UserCost or a
    registered DeviceCost."""
    ud, uc = isinstance(dynamics, UserDynamics), isinstance(cost, UserCost)
    if not (ud or uc):
        raise TypeError("neither the dynamics nor the cost is user-supplied: nothing to compile")
    mid = backend.user_model_register(
        dynamics.n, dynamics.m,
        dynamics_src=dynamics.src if ud else None, base_model_id=0 if ud else dynamics.model_id,
        n_model_params=dynamics.params.size,
        cost_src=cost.src if uc else None, base_cost_id=0 if uc else cost.cost_id,
        n_cost_params=cost.params().size if uc else 0,
        a_kind=dynamics.a_kind if ud else None, b_kind=dynamics.b_kind if ud else None,
        q_kind=cost.q_kind if uc else None, r_kind=cost.r_kind if uc else None, p_kind=cost.p_kind if uc else None)
    if ud:
        bound = UserDynamics(dynamics.n, dynamics.m, dynamics.src, dynamics.params, dynamics.py, dynamics.a_kind, dynamics.b_kind)
    else:
        bound = DeviceDynamics(dynamics.model_id, dynamics.params)
        bound.host_model_id = dynamics.model_id
    bound.model_id = mid
    return bound


class ConstantCovariance:
    """`W(k)` returning the same PD matrix for every k."""

    def __init__(self, W):
        self.W = np.asarray(W, dtype=np.float64)

    def __call__(self, k):
        return self.W


class DeviceStochasticDynamics:
    """`f_stochastic(x, u, rng, use_true_model=false)` = registered f + additive noise
    (optimal_control_problems.jl:82-87).  noise_kind 0: N(0, W); 1: uniform[0,1)*scale
    (test/pets_test.jl:15).  An ensemble is a list of parameter sets; the particle index
    selects the member."""

    def __init__(self, dynamics, W=None, noise_kind=0, noise_scale=1.0, ensemble_params=None, true_mixture=None):
        self.dynamics, self.noise_kind, self.noise_scale = dynamics, int(noise_kind), float(noise_scale)
        n = dynamics.n
        self.W = np.eye(n) if W is None else np.asarray(W, float)
        self.ensemble_params = None if ensemble_params is None else np.asarray(ensemble_params, float)
        # the accurate model behind `use_true_model=true` (optimal_control_problems.jl:105-109): a Gaussian mixture
        # dict(weights (k,), means (n, k), covs (n, n, k)); None -> the planner's model is also the true one
        self.true_mixture = true_mixture

    def __call__(self, x, u, rng, use_true_model=False):
        xn = self.dynamics(x, u)
        if use_true_model and self.true_mixture is not None:
            w = np.asarray(self.true_mixture["weights"], float)
            c = int(np.searchsorted(np.cumsum(w / w.sum()), rng.random(), side="right"))
            c = min(c, w.size - 1)
            mu = np.asarray(self.true_mixture["means"], float).reshape(xn.size, -1)[:, c]
            cov = np.asarray(self.true_mixture["covs"], float).reshape(xn.size, xn.size, -1)[:, :, c]
            return xn + mu + np.linalg.cholesky(cov) @ rng.standard_normal(xn.size)
        if self.noise_kind == 1:
            return xn + self.noise_scale * rng.random(xn.size)
        return xn + np.linalg.cholesky(self.W) @ rng.standard_normal(xn.size)

    def gen(self, use_true_model=False):
        g = dict(noise_kind=self.noise_kind, noise_scale=self.noise_scale, n_ensemble=1)
        if self.ensemble_params is not None:
            g["n_ensemble"] = self.ensemble_params.shape[0]
            g["ensemble_params"] = np.ascontiguousarray(self.ensemble_params).reshape(-1)
        if use_true_model and self.true_mixture is not None:
            g["use_true_model"], g["true_model"] = True, self.true_mixture
        return g
