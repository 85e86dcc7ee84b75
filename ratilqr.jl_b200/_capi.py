"""ctypes view of the C ABI declared in include/ratilqr.h.

`CApi(cdll, prefix)` binds the entry points `<prefix>ileqg_solve_batch`, ... of a shared
library and wraps them in numpy-friendly methods.  The product binds prefix ``ratilqr_`` of
libratilqr_b200.so (see _lib.py).  The same class is reused by the test-suite to drive the
CPU oracle (prefix ``oracle_``) so that one harness feeds both sides identical inputs.

All arrays are Fortran-ordered (column-major, instance index slowest), i.e. exactly what the
Julia shim passes (julia/src/RATiLQRB200.jl).
"""
import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class ProblemDesc(C.Structure):  # ratilqr_problem_desc
    _fields_ = [
        ("model_id", C.c_int32), ("cost_id", C.c_int32),
        ("n", C.c_int32), ("m", C.c_int32), ("N", C.c_int32),
        ("model_params", c_double_p), ("n_model_params", C.c_int32),
        ("cost_params", c_double_p), ("n_cost_params", C.c_int32),
        ("cost_params_count", C.c_int32),
        ("W", c_double_p), ("W_time_varying", C.c_int32),
    ]


class IleqgOpts(C.Structure):  # ratilqr_ileqg_opts  (ileqg.jl:165-175)
    _fields_ = [
        ("mu_min", C.c_double), ("delta_0", C.c_double), ("lam", C.c_double), ("d", C.c_double),
        ("iter_max", C.c_int32), ("adaptive_eps_init", C.c_int32),
        ("eps_init", C.c_double), ("eps_min", C.c_double), ("f_returns_jacobian", C.c_int32),
    ]


class BatchIn(C.Structure):  # ratilqr_batch_in
    _fields_ = [
        ("P", C.c_int32), ("K", C.c_int32),
        ("x0", c_double_p), ("x0_count", C.c_int32),
        ("u_init", c_double_p), ("u_count", C.c_int32),
        ("theta", c_double_p),
    ]


class IleqgOut(C.Structure):  # ratilqr_ileqg_out
    _fields_ = [
        ("x", c_double_p), ("l", c_double_p), ("L", c_double_p), ("value", c_double_p),
        ("status", c_int32_p), ("iters", c_int32_p), ("trials", c_int32_p), ("restarts", c_int32_p),
        ("mu", c_double_p), ("d_current", c_double_p),
        ("eps_hist", c_double_p), ("eps_hist_cap", C.c_int32),
    ]


class NoiseMixture(C.Structure):  # ratilqr_noise_mixture
    _fields_ = [("n_components", C.c_int32), ("weights", c_double_p), ("means", c_double_p), ("covs", c_double_p)]


class GenerativeDesc(C.Structure):  # ratilqr_generative_desc
    _fields_ = [
        ("noise_kind", C.c_int32), ("noise_scale", C.c_double),
        ("n_ensemble", C.c_int32), ("ensemble_params", c_double_p),
        ("true_model", C.POINTER(NoiseMixture)), ("use_true_model", C.c_int32),
    ]


class CeOpts(C.Structure):  # ratilqr_ce_opts
    _fields_ = [("num_samples", C.c_int32), ("num_elite", C.c_int32), ("iter_max", C.c_int32), ("lam", C.c_double),
                ("use_theta_max", C.c_int32)]


class MpcOpts(C.Structure):  # ratilqr_mpc_opts
    _fields_ = [("steps", C.c_int32), ("noise", c_double_p), ("noise_seed", C.c_uint64), ("true_noise", C.POINTER(NoiseMixture))]


class NmOpts(C.Structure):  # ratilqr_nm_opts
    _fields_ = [("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double), ("eps", C.c_double), ("lam", C.c_double),
                ("iter_max", C.c_int32)]


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


class ApiError(RuntimeError):
    pass


ILEQG_DEFAULTS = dict(mu_min=1e-6, delta_0=2.0, lam=0.5, d=1e-2, iter_max=100,
                      adaptive_eps_init=False, eps_init=1.0, eps_min=1e-6, f_returns_jacobian=False)


def make_opts(**kw):
    o = dict(ILEQG_DEFAULTS)
    o.update(kw)
    return IleqgOpts(o["mu_min"], o["delta_0"], o["lam"], o["d"], int(o["iter_max"]),
                     int(bool(o["adaptive_eps_init"])), o["eps_init"], o["eps_min"],
                     int(bool(o["f_returns_jacobian"])))


class UserModelDesc(C.Structure):  # ratilqr_user_model_desc
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("dynamics_src", C.c_char_p), ("base_model_id", C.c_int32),
                ("n_model_params", C.c_int32), ("cost_src", C.c_char_p), ("base_cost_id", C.c_int32),
                ("n_cost_params", C.c_int32), ("a_kind", C.POINTER(C.c_int8)), ("b_kind", C.POINTER(C.c_int8)),
                ("q_kind", C.POINTER(C.c_int8)), ("r_kind", C.POINTER(C.c_int8)), ("p_kind", C.POINTER(C.c_int8))]


MODEL_USER_BASE, COST_USER = 1000, 100


class Spec:
    """Plain description of a registered problem: what ratilqr_problem_desc carries."""

    def __init__(self, model_id, cost_id, n, m, N, model_params, cost_params, W, cost_params_count=1):
        self.model_id, self.cost_id, self.n, self.m, self.N = int(model_id), int(cost_id), int(n), int(m), int(N)
        self.model_params = _f64(model_params)
        cp = np.asarray(cost_params, dtype=np.float64)
        if cp.ndim == 2:  # (P, n_cost_params): one parameter block per problem
            cost_params_count = cp.shape[0]
        self.cost_params_count = int(cost_params_count)
        self.cost_params = np.ascontiguousarray(cp.reshape(-1))
        self.n_cost_params = self.cost_params.size // self.cost_params_count
        W = np.asarray(W, dtype=np.float64)
        self.W_time_varying = int(W.ndim == 3)
        # time-varying W is given as (N, n, n); each block stored column-major
        self.W = np.ascontiguousarray(np.stack([w.ravel(order="F") for w in W]).ravel()) if W.ndim == 3 else _f64(W)

    def desc(self):
        d = ProblemDesc(self.model_id, self.cost_id, self.n, self.m, self.N,
                        _dp(self.model_params), self.model_params.size,
                        _dp(self.cost_params), self.n_cost_params, self.cost_params_count,
                        _dp(self.W), self.W_time_varying)
        return d


class CApi:
    def __init__(self, cdll, prefix, needs_ctx):
        self.dll, self.prefix, self.needs_ctx = cdll, prefix, needs_ctx
        self.ctx = C.c_void_p(None)
        self._bind()
        if needs_ctx:
            self._create = getattr(cdll, prefix + "create")
            self._create.restype = C.c_int32
            self._create.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
            self._destroy = getattr(cdll, prefix + "destroy")
            self._destroy.argtypes = [C.c_void_p]
            self._last_error = getattr(cdll, prefix + "last_error")
            self._last_error.restype = C.c_char_p
            self._last_error.argtypes = [C.c_void_p]

    # -- binding --------------------------------------------------------------------------
    def _fn(self, name, argtypes, restype=C.c_int32):
        try:
            f = getattr(self.dll, self.prefix + name)
        except AttributeError:  # partial providers (the test-only host emulation) lack some entry points
            def missing(*a, **k):
                raise ApiError(f"{self.prefix}{name} is not exported by this library")
            return missing
        f.restype = restype
        f.argtypes = argtypes
        return f

    def _bind(self):
        vp = C.c_void_p
        PD, IO, BI, OUT, GD = (C.POINTER(ProblemDesc), C.POINTER(IleqgOpts), C.POINTER(BatchIn),
                               C.POINTER(IleqgOut), C.POINTER(GenerativeDesc))
        dp, ip, i32, f64 = c_double_p, c_int32_p, C.c_int32, C.c_double
        self.f_solve = self._fn("ileqg_solve_batch", [vp, PD, IO, BI, OUT])
        self.f_ce_costs = self._fn("ce_costs", [vp, PD, IO, BI, f64, dp, ip])
        self.f_open = self._fn("rollout_open_batch", [vp, PD, i32, dp, dp, dp, ip])
        self.f_closed = self._fn("rollout_closed_batch", [vp, PD, i32, dp, dp, dp, dp, dp, ip])
        self.f_cost = self._fn("integrate_cost_batch", [vp, PD, i32, dp, dp, dp, ip])
        self.f_lin = self._fn("linearize_batch", [vp, PD, i32, dp, dp] + [dp] * 8 + [ip])
        self.f_ric = self._fn("riccati_batch_tv", [vp, i32, i32, i32, i32, i32] + [dp] * 8 + [dp, i32, dp, f64, f64, dp, dp,
                                                                                             dp, dp, dp, dp, dp, ip, ip])
        self.f_mc = self._fn("mc_rollout", [vp, PD, i32, dp, dp, dp, i32, dp, C.c_uint64, f64, dp, dp, dp])
        self.f_mc_true = self._fn("mc_rollout_true_model", [vp, PD, i32, dp, dp, dp, i32, C.POINTER(NoiseMixture), C.c_uint64,
                                                            f64, dp, dp, dp])
        self.f_pets_costs = self._fn("pets_costs", [vp, PD, GD, dp, dp, i32, i32, dp, C.c_uint64, dp])
        self.f_pets_refit = self._fn("pets_refit", [vp, i32, i32, i32, i32, f64, dp, dp, dp, dp, ip])
        self.f_ce_fleet = self._fn("ce_solve_fleet", [vp, PD, IO, C.POINTER(CeOpts), i32, dp, i32, dp, i32, f64, dp, C.c_int64,
                                                      C.c_uint64, dp, dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int64),
                                                      ip, OUT])
        self.f_nm_fleet = self._fn("nm_solve_fleet", [vp, PD, IO, C.POINTER(NmOpts), i32, dp, i32, dp, i32, f64, dp, dp, dp, dp, ip,
                                                      dp, dp, ip, ip, OUT])
        self.f_mpc = self._fn("mpc_fleet_run", [vp, PD, IO, C.POINTER(CeOpts), C.POINTER(MpcOpts), i32, dp, dp, i32, f64, dp, C.c_int64,
                                                C.c_uint64, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_float), C.POINTER(i32)])
        u8p = C.POINTER(C.c_uint8)
        self.f_uid = self._fn("nccl_unique_id", [u8p])
        self.f_attach = self._fn("attach_comm", [vp, u8p, i32, i32])
        self.f_ce_costs_sh = self._fn("ce_costs_sharded", [vp, PD, IO, dp, dp, dp, i32, f64, dp, ip])
        self.f_pets_costs_sh = self._fn("pets_costs_sharded", [vp, PD, GD, dp, dp, i32, i32, dp, C.c_uint64, dp])
        # (the oracle exports oracle_ce_solve / oracle_nm_solve with its OWN, older argument lists -- the tests call those
        #  directly -- so the single-problem entry points are bound for the product library only)
        one = (lambda name, at: self._fn(name, at)) if self.needs_ctx else (lambda name, at: self._fn("__absent__" + name, at))
        self.f_ce_one = one("ce_solve", [vp, PD, IO, C.POINTER(CeOpts), dp, dp, f64, dp, C.c_int64, C.c_uint64, dp, dp, dp, dp,
                                              dp, dp, dp, dp, C.POINTER(C.c_int64), C.POINTER(i32), OUT])
        self.f_ce_one_sh = one("ce_solve_sharded", [vp, PD, IO, C.POINTER(CeOpts), dp, dp, f64, dp, C.c_int64, C.c_uint64, dp, dp, dp,
                                                    dp, dp, dp, dp, dp, C.POINTER(C.c_int64), C.POINTER(i32), OUT])
        self.f_nm_one = one("nm_solve", [vp, PD, IO, C.POINTER(NmOpts), dp, dp, f64, dp, dp, dp, dp, ip, dp, dp, ip, ip, OUT])
        self.f_stage = self._fn("ileqg_stage", [vp, PD, IO, BI])
        self.f_run = self._fn("ileqg_run", [vp, i32, C.POINTER(C.c_float)])
        self.f_fetch = self._fn("ileqg_fetch", [vp, OUT])
        self.f_probe = self._fn("fp64_peak_probe", [vp, dp, C.POINTER(C.c_float)])
        self.f_probe_sus = self._fn("fp64_peak_probe_sustained", [vp, f64, dp])
        self.f_launches = self._fn("launch_count", [vp], restype=C.c_int64)
        self.f_pets_solve = self._fn("pets_solve", [vp, PD, GD, dp, i32, i32, i32, i32, f64, dp, dp, C.c_uint64, dp, dp])
        UM = C.POINTER(UserModelDesc)
        self.f_um_check = self._fn("user_model_check", [UM, C.c_char_p, C.c_int64])
        self.f_um_register = self._fn("user_model_register", [vp, UM, ip, C.c_char_p, C.c_int64])

    # -- user-extensible device models (NVRTC) ------------------------------------------------
    @staticmethod
    def _um_desc(n, m, dynamics_src, base_model_id, n_model_params, cost_src, base_cost_id, n_cost_params, kinds):
        enc = lambda t: None if t is None else t.encode()
        keep = []

        def kp(name, rows, cols):  # (rows, cols) table of 0 / 1 / 2 -> column-major int8 buffer
            k = kinds.get(name)
            if k is None:
                return None
            a = np.asfortranarray(np.asarray(k, dtype=np.int8).reshape(rows, cols))
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int8))
        d = UserModelDesc(int(n), int(m), enc(dynamics_src), int(base_model_id), int(n_model_params), enc(cost_src),
                          int(base_cost_id), int(n_cost_params), kp("a_kind", n, n), kp("b_kind", n, m), kp("q_kind", n, n),
                          kp("r_kind", m, m), kp("p_kind", m, n))
        return d, keep

    def user_model_check(self, n, m, dynamics_src=None, cost_src=None, base_model_id=0, base_cost_id=0,
                         n_model_params=0, n_cost_params=0, **kinds):
        """compile only (no GPU needed) -> (status, compiler log).  kinds: a_kind (n, n), b_kind (n, m), q_kind (n, n),
        r_kind (m, m), p_kind (m, n) tables of 0 / 1 / 2 declaring the structure of the derivatives (default dense)."""
        d, keep = self._um_desc(n, m, dynamics_src, base_model_id, n_model_params, cost_src, base_cost_id, n_cost_params, kinds)
        log = C.create_string_buffer(1 << 16)
        rc = self.f_um_check(C.byref(d), log, len(log))
        return rc, log.value.decode(errors="replace")

    def user_model_register(self, n, m, dynamics_src=None, cost_src=None, base_model_id=0, base_cost_id=0,
                            n_model_params=0, n_cost_params=0, **kinds):
        """compile + load into this context -> model id to put into Spec.model_id (kinds as in user_model_check; a
        declaration that the dual-number derivatives contradict at the probe points is refused)"""
        d, keep = self._um_desc(n, m, dynamics_src, base_model_id, n_model_params, cost_src, base_cost_id, n_cost_params, kinds)
        log = C.create_string_buffer(1 << 16)
        mid = C.c_int32(0)
        rc = self.f_um_register(self.ctx, C.byref(d), C.byref(mid), log, len(log))
        if rc != 0:
            raise ApiError(f"user_model_register failed ({rc}): {log.value.decode(errors='replace')[:4000]}")
        return int(mid.value)

    def open(self, device_id=0):
        if self.needs_ctx and not self.ctx:
            rc = self._create(C.byref(self.ctx), device_id)
            if rc != 0:
                raise ApiError(f"{self.prefix}create failed with code {rc}")
        return self

    def close(self):
        if self.needs_ctx and self.ctx:
            self._destroy(self.ctx)
            self.ctx = C.c_void_p(None)

    def _check(self, rc, what):
        if rc != 0:
            msg = ""
            if self.needs_ctx and self.ctx:
                m = self._last_error(self.ctx)
                msg = m.decode() if m else ""
            raise ApiError(f"{self.prefix}{what} failed with code {rc}: {msg}")

    # -- wrappers --------------------------------------------------------------------------
    @staticmethod
    def _batch(spec, x0, u_init, theta, P=None):
        n, m, N = spec.n, spec.m, spec.N
        theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
        x0 = np.asarray(x0, dtype=np.float64)
        u_init = np.asarray(u_init, dtype=np.float64)
        x0_count = 1 if x0.ndim == 1 else x0.shape[-1]
        u_count = 1 if u_init.ndim == 2 else u_init.shape[-1]
        if P is None:
            P = max(x0_count, u_count, spec.cost_params_count)
        assert theta.size % P == 0, "theta count must be a multiple of the problem count"
        K = theta.size // P
        assert x0.shape[0] == n and u_init.shape[0] == m and u_init.shape[1] == N
        x0f, uf = _f64(x0), _f64(u_init)
        bi = BatchIn(P, K, _dp(x0f), x0_count, _dp(uf), u_count, _dp(theta))
        return bi, (x0f, uf, theta)

    def ileqg_solve_batch(self, spec, x0, u_init, theta, opts=None, want=("x", "l", "L"), eps_hist_cap=0, P=None):
        """solve!(::ILEQGSolver, ...) for a batch (ileqg.jl:635-659). Returns a dict of arrays."""
        opts = opts or make_opts()
        n, m, N = spec.n, spec.m, spec.N
        bi, keep = self._batch(spec, x0, u_init, theta, P)
        B = bi.P * bi.K
        res = dict(value=np.empty(B), status=np.empty(B, np.int32), iters=np.empty(B, np.int32),
                   trials=np.empty(B, np.int32), restarts=np.empty(B, np.int32), mu=np.empty(B),
                   d_current=np.empty(B))
        if "x" in want:
            res["x"] = np.zeros((n, N + 1, B), order="F")
        if "l" in want:
            res["l"] = np.zeros((m, N, B), order="F")
        if "L" in want:
            res["L"] = np.zeros((m, n, N, B), order="F")
        if eps_hist_cap:
            res["eps_hist"] = np.zeros((2, eps_hist_cap, B), order="F")
        out = IleqgOut(_dp(res.get("x")), _dp(res.get("l")), _dp(res.get("L")), _dp(res["value"]),
                       _ip(res["status"]), _ip(res["iters"]), _ip(res["trials"]), _ip(res["restarts"]),
                       _dp(res["mu"]), _dp(res["d_current"]), _dp(res.get("eps_hist")), eps_hist_cap)
        d = spec.desc()
        self._check(self.f_solve(self.ctx, C.byref(d), C.byref(opts), C.byref(bi), C.byref(out)), "ileqg_solve_batch")
        return res

    def ce_solve_fleet(self, spec, x0, u_init, kl_bound, mu_init, sigma_init, num_samples=10, num_elite=3, iter_max=5,
                       lam=0.5, use_theta_max=False, z_inject=None, seed=0, opts=None, want=("x", "l", "L")):
        """solve!(::CrossEntropyBilevelOptimizationSolver) for P problems at once (cross_entropy...:364-415).
        x0 (n, P); mu_init/sigma_init scalars or (P,); z_inject (P, nz) standard normals or None -> Philox(seed)."""
        opts = opts or make_opts()
        n, m, N = spec.n, spec.m, spec.N
        x0 = np.asarray(x0, dtype=np.float64).reshape(n, -1)
        P = max(x0.shape[1], spec.cost_params_count)
        u_init = np.asarray(u_init, dtype=np.float64)
        u_count = 1 if u_init.ndim == 2 else u_init.shape[-1]
        x0f, uf = _f64(x0), _f64(u_init)
        mu_i = np.ascontiguousarray(np.broadcast_to(np.asarray(mu_init, np.float64), (P,))).copy()
        sg_i = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma_init, np.float64), (P,))).copy()
        zf, nz = None, 0
        if z_inject is not None:
            z = np.ascontiguousarray(np.asarray(z_inject, dtype=np.float64).reshape(P, -1))
            zf, nz = z.reshape(-1), z.shape[1]
        res = dict(theta_opt=np.zeros(P), value=np.zeros(P), theta_min=np.zeros(P), theta_max=np.zeros(P), mu=np.zeros(P),
                   sigma=np.zeros(P), nz_used=np.zeros(P, np.int64), status=np.zeros(P, np.int32), iters=np.zeros(P, np.int32))
        if "x" in want:
            res["x"] = np.zeros((n, N + 1, P), order="F")
        if "l" in want:
            res["l"] = np.zeros((m, N, P), order="F")
        if "L" in want:
            res["L"] = np.zeros((m, n, N, P), order="F")
        out = IleqgOut(_dp(res.get("x")), _dp(res.get("l")), _dp(res.get("L")), None, _ip(res["status"]), _ip(res["iters"]),
                       None, None, None, None, None, 0)
        ce = CeOpts(int(num_samples), int(num_elite), int(iter_max), float(lam), int(bool(use_theta_max)))
        rounds = C.c_int32(0)
        d = spec.desc()
        self._check(self.f_ce_fleet(self.ctx, C.byref(d), C.byref(opts), C.byref(ce), P, _dp(x0f), x0.shape[1], _dp(uf), u_count,
                                    float(kl_bound), _dp(zf), nz, int(seed), _dp(mu_i), _dp(sg_i), _dp(res["theta_opt"]),
                                    _dp(res["value"]), _dp(res["theta_min"]), _dp(res["theta_max"]), _dp(res["mu"]),
                                    _dp(res["sigma"]), res["nz_used"].ctypes.data_as(C.POINTER(C.c_int64)), C.byref(rounds),
                                    C.byref(out)), "ce_solve_fleet")
        res["mu_init"], res["sigma_init"], res["rounds"] = mu_i, sg_i, int(rounds.value)
        return res

    # ---- multi-GPU: NCCL communicator owned by the ctx (include/ratilqr.h "multi-GPU") ------------------------------
    def nccl_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._check(self.f_uid(buf), "nccl_unique_id")
        return bytes(buf)

    def attach_comm(self, uid, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._check(self.f_attach(self.ctx, buf, int(rank), int(world)), "attach_comm")
        self.comm_rank, self.comm_world = int(rank), int(world)

    def ce_costs_sharded(self, spec, x0, u_init, theta, kl_bound, opts=None):
        """collective: compute_cost of one problem's theta population, block-sharded over the ranks of the attached
        communicator; the all-gather runs inside the library on device buffers.  Every rank gets (cost, status)."""
        opts = opts or make_opts()
        theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
        x0f, uf = _f64(x0), _f64(u_init)
        cost, status = np.empty(theta.size), np.empty(theta.size, np.int32)
        d = spec.desc()
        self._check(self.f_ce_costs_sh(self.ctx, C.byref(d), C.byref(opts), _dp(x0f), _dp(uf), _dp(theta), theta.size,
                                       float(kl_bound), _dp(cost), _ip(status)), "ce_costs_sharded")
        return cost, status

    def pets_costs_sharded(self, spec, x0, controls, particles, noise=None, seed=0, gen=None):
        """collective: PETS compute_cost with the action sequences block-sharded over the ranks"""
        n, m, N = spec.n, spec.m, spec.N
        cf = _f64(controls)
        Cn = cf.size // (m * N)
        nf = None if noise is None else _f64(noise)
        cost = np.zeros(Cn)
        g, keep = self._gen(gen)
        x0f = _f64(x0)
        d = spec.desc()
        self._check(self.f_pets_costs_sh(self.ctx, C.byref(d), C.byref(g), _dp(x0f), _dp(cf), Cn, int(particles), _dp(nf), int(seed),
                                         _dp(cost)), "pets_costs_sharded")
        del keep
        return cost

    def ce_solve(self, spec, x0, u_init, kl_bound, mu_init, sigma_init, num_samples=10, num_elite=3, iter_max=5, lam=0.5,
                 use_theta_max=False, z_inject=None, seed=0, opts=None, sharded=False):
        """solve!(::CrossEntropyBilevelOptimizationSolver) for ONE problem, whole loop on the device (ratilqr_ce_solve);
        sharded: collective over the attached communicator (ratilqr_ce_solve_sharded)."""
        opts = opts or make_opts()
        n, m, N = spec.n, spec.m, spec.N
        x0f, uf = _f64(x0), _f64(u_init)
        zf, nz = None, 0
        if z_inject is not None:
            zf = np.ascontiguousarray(np.asarray(z_inject, dtype=np.float64).ravel())
            nz = zf.size
        sc = {k: np.zeros(1) for k in ("theta_opt", "value", "theta_min", "theta_max", "mu", "sigma")}
        mu_i, sg_i = np.array([float(mu_init)]), np.array([float(sigma_init)])
        nzu, st, it = np.zeros(1, np.int64), np.zeros(1, np.int32), np.zeros(1, np.int32)
        x, l, L = np.zeros((n, N + 1), order="F"), np.zeros((m, N), order="F"), np.zeros((m, n, N), order="F")
        out = IleqgOut(_dp(x), _dp(l), _dp(L), None, _ip(st), _ip(it), None, None, None, None, None, 0)
        ce = CeOpts(int(num_samples), int(num_elite), int(iter_max), float(lam), int(bool(use_theta_max)))
        rounds = C.c_int32(0)
        d = spec.desc()
        fn = self.f_ce_one_sh if sharded else self.f_ce_one
        self._check(fn(self.ctx, C.byref(d), C.byref(opts), C.byref(ce), _dp(x0f), _dp(uf), float(kl_bound), _dp(zf), nz,
                                  int(seed), _dp(mu_i), _dp(sg_i), _dp(sc["theta_opt"]), _dp(sc["value"]), _dp(sc["theta_min"]),
                                  _dp(sc["theta_max"]), _dp(sc["mu"]), _dp(sc["sigma"]), nzu.ctypes.data_as(C.POINTER(C.c_int64)),
       C.byref(rounds), C.byref(out)), "ce_solve")
        res = {k: float(v[0]) for k, v in sc.items()}
        res.update(mu_init=float(mu_i[0]), sigma_init=float(sg_i[0]), nz_used=int(nzu[0]), rounds=int(rounds.value),
                   status=int(st[0]), iters=int(it[0]), x=x, l=l, L=L)
        return res

    def mpc_fleet_run(self, spec, x0, u_init, steps, kl_bound, mu_init, sigma_init, num_samples=10, num_elite=3, iter_max=5,
                      lam=0.5, use_theta_max=False, z_inject=None, seed=0, noise=None, noise_seed=0, true_mixture=None, opts=None):
        """Receding-horizon RAT iLQR for a fleet, `steps` MPC steps on the device (ratilqr_mpc_fleet_run).
        x0 (n, P); u_init (m, N) or (m, N, P); noise (n, steps, P) injected disturbances or None -> Philox;
        z_inject (steps, P, nz) standard normals for the theta draws or None -> Philox."""
        opts = opts or make_opts()
        n, m, N = spec.n, spec.m, spec.N
        x0 = np.asarray(x0, dtype=np.float64).reshape(n, -1)
        P = x0.shape[1]
        u_init = np.asarray(u_init, dtype=np.float64)
        u_count = 1 if u_init.ndim == 2 else u_init.shape[-1]
        x0f, uf = _f64(x0), _f64(u_init)
        mu_i = np.ascontiguousarray(np.broadcast_to(np.asarray(mu_init, np.float64), (P,))).copy()
        sg_i = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma_init, np.float64), (P,))).copy()
        zf, nz = None, 0
        if z_inject is not None:
            z = np.ascontiguousarray(np.asarray(z_inject, dtype=np.float64).reshape(steps, P, -1))
            zf, nz = z.reshape(-1), z.shape[2]
        nf = None if noise is None else _f64(np.asarray(noise, dtype=np.float64).reshape(n, steps, P))
        keep = None
        mo = MpcOpts(int(steps), _dp(nf), int(noise_seed), None)
        if true_mixture is not None:
            mix, keep = self._mixture(true_mixture)
            mo.true_noise = C.pointer(mix)
        ce = CeOpts(int(num_samples), int(num_elite), int(iter_max), float(lam), int(bool(use_theta_max)))
        xt = np.zeros((n, steps + 1, P), order="F")
        ut = np.zeros((m, steps, P), order="F")
        tt, vt = np.zeros((steps, P), order="F"), np.zeros((steps, P), order="F")
        ms = np.zeros(steps, np.float32)
        rounds = C.c_int32(0)
        d = spec.desc()
        self._check(self.f_mpc(self.ctx, C.byref(d), C.byref(opts), C.byref(ce), C.byref(mo), P, _dp(x0f), _dp(uf), u_count,
                               float(kl_bound), _dp(zf), nz, int(seed), _dp(mu_i), _dp(sg_i), _dp(xt), _dp(ut), _dp(tt), _dp(vt),
                               ms.ctypes.data_as(C.POINTER(C.c_float)), C.byref(rounds)), "mpc_fleet_run")
        del keep
        return dict(x=xt, u=ut, theta=tt, value=vt, ms=ms.astype(float), mu_init=mu_i, sigma_init=sg_i, rounds=int(rounds.value))

    def nm_solve(self, spec, x0, u_init, kl_bound, state=None, alpha=1.0, beta=2.0, gamma=0.5, eps=1e-2, lam=0.5, iter_max=100,
                 theta_high_init=3.0, theta_low_init=1e-8, opts=None):
        """solve!(::NelderMeadBilevelOptimizationSolver) for ONE problem, whole loop on the device (ratilqr_nm_solve)."""
        opts = opts or make_opts()
        n, m, N = spec.n, spec.m, spec.N
        x0f, uf = _f64(x0), _f64(u_init)
        if state is None:
            state = dict(theta_high_init=np.array([float(theta_high_init)]), theta_low_init=np.array([float(theta_low_init)]),
                         c_high=np.zeros(1), c_low=np.zeros(1), has_c=np.zeros(2, np.int32))
        stt = {k: np.ascontiguousarray(v).copy() for k, v in state.items()}
        th, val = np.zeros(1), np.zeros(1)
        it, ev, st = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32)
        x, l, L = np.zeros((n, N + 1), order="F"), np.zeros((m, N), order="F"), np.zeros((m, n, N), order="F")
        out = IleqgOut(_dp(x), _dp(l), _dp(L), None, _ip(st), None, None, None, None, None, None, 0)
        nm = NmOpts(alpha, beta, gamma, eps, lam, int(iter_max))
        d = spec.desc()
        self._check(self.f_nm_one(self.ctx, C.byref(d), C.byref(opts), C.byref(nm), _dp(x0f), _dp(uf), float(kl_bound),
                                  _dp(stt["theta_high_init"]), _dp(stt["theta_low_init"]), _dp(stt["c_high"]), _dp(stt["c_low"]),
                                  _ip(stt["has_c"]), _dp(th), _dp(val), _ip(it), _ip(ev), C.byref(out)), "nm_solve")
        return dict(theta_opt=float(th[0]), value=float(val[0]), nm_iters=int(it[0]), n_evals=int(ev[0]), status=int(st[0]),
                    x=x, l=l, L=L, state=stt)

    def nm_solve_fleet(self, spec, x0, u_init, kl_bound, state=None, alpha=1.0, beta=2.0, gamma=0.5, eps=1e-2, lam=0.5,
                       iter_max=100, theta_high_init=3.0, theta_low_init=1e-8, opts=None, want=("x", "l", "L")):
        """solve!(::NelderMeadBilevelOptimizationSolver) for P problems at once (nelder_mead...:276-352).
        `state` (returned by a previous call) carries theta_*_init, c_high, c_low, has_c across calls like the Julia struct."""
        opts = opts or make_opts()
        n, m, N = spec.n, spec.m, spec.N
        x0 = np.asarray(x0, dtype=np.float64).reshape(n, -1)
        P = max(x0.shape[1], spec.cost_params_count)
        u_init = np.asarray(u_init, dtype=np.float64)
        u_count = 1 if u_init.ndim == 2 else u_init.shape[-1]
        x0f, uf = _f64(x0), _f64(u_init)
        if state is None:
            state = dict(theta_high_init=np.full(P, float(theta_high_init)), theta_low_init=np.full(P, float(theta_low_init)),
                         c_high=np.zeros(P), c_low=np.zeros(P), has_c=np.zeros((P, 2), np.int32))
        st = {k: np.ascontiguousarray(v).copy() for k, v in state.items()}
        res = dict(theta_opt=np.zeros(P), value=np.zeros(P), nm_iters=np.zeros(P, np.int32), n_evals=np.zeros(P, np.int32),
                   status=np.zeros(P, np.int32))
        if "x" in want:
            res["x"] = np.zeros((n, N + 1, P), order="F")
        if "l" in want:
            res["l"] = np.zeros((m, N, P), order="F")
        if "L" in want:
            res["L"] = np.zeros((m, n, N, P), order="F")
        out = IleqgOut(_dp(res.get("x")), _dp(res.get("l")), _dp(res.get("L")), None, _ip(res["status"]), None, None, None,
                       None, None, None, 0)
        nm = NmOpts(alpha, beta, gamma, eps, lam, int(iter_max))
        d = spec.desc()
        self._check(self.f_nm_fleet(self.ctx, C.byref(d), C.byref(opts), C.byref(nm), P, _dp(x0f), x0.shape[1], _dp(uf), u_count,
                                    float(kl_bound), _dp(st["theta_high_init"]), _dp(st["theta_low_init"]), _dp(st["c_high"]),
                                    _dp(st["c_low"]), _ip(st["has_c"]), _dp(res["theta_opt"]), _dp(res["value"]),
                                    _ip(res["nm_iters"]), _ip(res["n_evals"]), C.byref(out)), "nm_solve_fleet")
        res["state"] = st
        return res

    # device-resident variant (throughput measurement): stage once, run many, fetch
    def stage(self, spec, x0, u_init, theta, opts=None, P=None):
        opts = opts or make_opts()
        bi, keep = self._batch(spec, x0, u_init, theta, P)
        d = spec.desc()
        self._check(self.f_stage(self.ctx, C.byref(d), C.byref(opts), C.byref(bi)), "ileqg_stage")
        self._staged = (spec.n, spec.m, spec.N, bi.P * bi.K)

    def run(self, reps=1):
        """`reps` back-to-back launches of the solve kernel; returns the CUDA-event time in ms
        (events recorded on the library's own stream)."""
        ms = C.c_float(0.0)
        self._check(self.f_run(self.ctx, int(reps), C.byref(ms)), "ileqg_run")
        return float(ms.value)

    def fetch(self, want=()):
        n, m, N, B = self._staged
        res = dict(value=np.empty(B), status=np.empty(B, np.int32), iters=np.empty(B, np.int32),
                   trials=np.empty(B, np.int32), restarts=np.empty(B, np.int32), mu=np.empty(B),
                   d_current=np.empty(B))
        if "x" in want:
            res["x"] = np.zeros((n, N + 1, B), order="F")
        if "l" in want:
            res["l"] = np.zeros((m, N, B), order="F")
        if "L" in want:
            res["L"] = np.zeros((m, n, N, B), order="F")
        out = IleqgOut(_dp(res.get("x")), _dp(res.get("l")), _dp(res.get("L")), _dp(res["value"]),
                       _ip(res["status"]), _ip(res["iters"]), _ip(res["trials"]), _ip(res["restarts"]),
                       _dp(res["mu"]), _dp(res["d_current"]), None, 0)
        self._check(self.f_fetch(self.ctx, C.byref(out)), "ileqg_fetch")
        return res

    def fp64_probe(self):
        """measured non-tensor FP64 FMA throughput of this device in TFLOP/s"""
        tf, ms = C.c_double(0.0), C.c_float(0.0)
        self._check(self.f_probe(self.ctx, C.byref(tf), C.byref(ms)), "fp64_peak_probe")
        return float(tf.value)

    def fp64_probe_sustained(self, seconds=1.0):
        """DFMA throughput in TFLOP/s over the second half of a `seconds`-long back-to-back run (power-capped clocks)"""
        tf = C.c_double(0.0)
        self._check(self.f_probe_sus(self.ctx, float(seconds), C.byref(tf)), "fp64_peak_probe_sustained")
        return float(tf.value)

    def launch_count(self):
        return int(self.f_launches(self.ctx))

    def ce_costs(self, spec, x0, u_init, theta, kl_bound, opts=None, P=None):
        """compute_cost (cross_entropy_bilevel_optimization.jl:173-195)."""
        opts = opts or make_opts()
        bi, keep = self._batch(spec, x0, u_init, theta, P)
        B = bi.P * bi.K
        cost, status = np.empty(B), np.empty(B, np.int32)
        d = spec.desc()
        self._check(self.f_ce_costs(self.ctx, C.byref(d), C.byref(opts), C.byref(bi), float(kl_bound),
                                    _dp(cost), _ip(status)), "ce_costs")
        return cost, status

    def rollout_open(self, spec, x0, u):
        n, m, N = spec.n, spec.m, spec.N
        x0 = np.asarray(x0, dtype=np.float64).reshape(n, -1)
        B = x0.shape[1]
        u = np.asarray(u, dtype=np.float64).reshape(m, N, B)
        x = np.zeros((n, N + 1, B), order="F")
        st = np.zeros(B, np.int32)
        d = spec.desc()
        x0f, uf = _f64(x0), _f64(u)
        self._check(self.f_open(self.ctx, C.byref(d), B, _dp(x0f), _dp(uf), _dp(x), _ip(st)), "rollout_open_batch")
        return x, st

    def rollout_closed(self, spec, xbar, l, L):
        n, m, N = spec.n, spec.m, spec.N
        xbar = np.asarray(xbar, dtype=np.float64).reshape(n, N + 1, -1)
        B = xbar.shape[2]
        lf, Lf, xf = _f64(np.asarray(l).reshape(m, N, B)), _f64(np.asarray(L).reshape(m, n, N, B)), _f64(xbar)
        xn = np.zeros((n, N + 1, B), order="F")
        un = np.zeros((m, N, B), order="F")
        st = np.zeros(B, np.int32)
        d = spec.desc()
        self._check(self.f_closed(self.ctx, C.byref(d), B, _dp(xf), _dp(lf), _dp(Lf), _dp(xn), _dp(un), _ip(st)),
                    "rollout_closed_batch")
        return xn, un, st

    def integrate_cost(self, spec, x, u):
        n, m, N = spec.n, spec.m, spec.N
        x = np.asarray(x, dtype=np.float64).reshape(n, N + 1, -1)
        B = x.shape[2]
        xf, uf = _f64(x), _f64(np.asarray(u).reshape(m, N, B))
        cost = np.zeros(B)
        st = np.zeros(B, np.int32)
        d = spec.desc()
        self._check(self.f_cost(self.ctx, C.byref(d), B, _dp(xf), _dp(uf), _dp(cost), _ip(st)), "integrate_cost_batch")
        return cost, st

    def linearize(self, spec, x, u):
        """approximate_model (ileqg.jl:258-322)."""
        n, m, N = spec.n, spec.m, spec.N
        x = np.asarray(x, dtype=np.float64).reshape(n, N + 1, -1)
        B = x.shape[2]
        xf, uf = _f64(x), _f64(np.asarray(u).reshape(m, N, B))
        r = dict(q=np.zeros((N + 1, B), order="F"), qv=np.zeros((n, N + 1, B), order="F"),
                 Q=np.zeros((n, n, N + 1, B), order="F"), r=np.zeros((m, N, B), order="F"),
                 R=np.zeros((m, m, N, B), order="F"), P=np.zeros((m, n, N, B), order="F"),
                 A=np.zeros((n, n, N, B), order="F"), B=np.zeros((n, m, N, B), order="F"))
        st = np.zeros(B, np.int32)
        d = spec.desc()
        self._check(self.f_lin(self.ctx, C.byref(d), B, _dp(xf), _dp(uf), _dp(r["q"]), _dp(r["qv"]), _dp(r["Q"]),
                               _dp(r["r"]), _dp(r["R"]), _dp(r["P"]), _dp(r["A"]), _dp(r["B"]), _ip(st)),
                    "linearize_batch")
        r["status"] = st
        return r

    def riccati(self, lin, W, theta, optimise, L=None, dl=None, mu=0.0, delta=2.0, mu_min=1e-6, delta_0=2.0):
        """solve_approximate_dp! (optimise) / solve_approximate_dp (evaluate) on given approximations."""
        n, N1, B = lin["qv"].shape
        N = N1 - 1
        m = lin["r"].shape[0]
        theta = np.ascontiguousarray(np.broadcast_to(np.asarray(theta, dtype=np.float64), (B,)))
        mu_a = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, dtype=np.float64), (B,))).copy()
        de_a = np.ascontiguousarray(np.broadcast_to(np.asarray(delta, dtype=np.float64), (B,))).copy()
        # W: one n x n matrix, or n x n x N (W(k) per stage, ileqg.jl:364,438)
        Wa = np.asarray(W, dtype=np.float64)
        W_tv = 1 if Wa.ndim == 3 else 0
        if W_tv:
            assert Wa.shape == (n, n, N), "time-varying W must be n x n x N"
        Wf = _f64(Wa)
        arrs = [_f64(lin[k]) for k in ("q", "qv", "Q", "r", "R", "P", "A", "B")]
        if optimise:
            Lf = np.zeros(m * n * N * B)
            dlf = np.zeros(m * N * B)
        else:
            Lf = _f64(np.asarray(L).reshape(m, n, N, B)).copy()
            dlf = None if dl is None else _f64(np.asarray(dl).reshape(m, N, B)).copy()
        s = np.zeros((N + 1, B), order="F")
        sv = np.zeros((n, N + 1, B), order="F")
        S = np.zeros((n, n, N + 1, B), order="F")
        st = np.zeros(B, np.int32)
        rs = np.zeros(B, np.int32)
        self._check(self.f_ric(self.ctx, n, m, N, B, int(optimise), *[_dp(a) for a in arrs], _dp(Wf), W_tv, _dp(theta),
                               float(mu_min), float(delta_0), _dp(mu_a), _dp(de_a), _dp(Lf), _dp(dlf),
                               _dp(s), _dp(sv), _dp(S), _ip(st), _ip(rs)), "riccati_batch")
        return dict(s=s, sv=sv, S=S, status=st, restarts=rs, mu=mu_a, delta=de_a,
                    L=Lf.reshape((m, n, N, B), order="F"),
                    dl=None if dlf is None else dlf.reshape((m, N, B), order="F"))

    def mc_rollout(self, spec, xbar, l, L, n_samples, noise=None, seed=0, theta_risk=0.0, want_x=False, P=1):
        n, m, N = spec.n, spec.m, spec.N
        xf, lf, Lf = _f64(xbar), _f64(l), _f64(L)
        nf = None if noise is None else _f64(noise)
        J = np.zeros(n_samples * P)
        stats = np.zeros(3 * P)
        xo = np.zeros((n, N + 1, n_samples * P), order="F") if want_x else None
        d = spec.desc()
        self._check(self.f_mc(self.ctx, C.byref(d), P, _dp(xf), _dp(lf), _dp(Lf), n_samples, _dp(nf), int(seed),
                              float(theta_risk), _dp(J), _dp(stats), _dp(xo)), "mc_rollout")
        return dict(J=J, stats=stats.reshape(P, 3), x=xo)

    @staticmethod
    def _mixture(mix):
        """mix: dict(weights (k,), means (n, k), covs (n, n, k)) -> (NoiseMixture, keep-alive buffers)"""
        w = _f64(mix["weights"])
        mu = _f64(np.asarray(mix["means"], dtype=np.float64).reshape(-1, w.size))
        cv = _f64(np.asarray(mix["covs"], dtype=np.float64).reshape(mu.size // w.size, -1, w.size))
        return NoiseMixture(int(w.size), _dp(w), _dp(mu), _dp(cv)), (w, mu, cv)

    def mc_rollout_true_model(self, spec, xbar, l, L, n_samples, mixture, seed=0, theta_risk=0.0, want_x=False, P=1):
        """closed-loop MC evaluation under the true (Gaussian-mixture) noise model"""
        n, N = spec.n, spec.N
        xf, lf, Lf = _f64(xbar), _f64(l), _f64(L)
        J = np.zeros(n_samples * P)
        stats = np.zeros(3 * P)
        xo = np.zeros((n, N + 1, n_samples * P), order="F") if want_x else None
        d = spec.desc()
        mx, keep = self._mixture(mixture)
        self._check(self.f_mc_true(self.ctx, C.byref(d), P, _dp(xf), _dp(lf), _dp(Lf), n_samples, C.byref(mx), int(seed),
                                   float(theta_risk), _dp(J), _dp(stats), _dp(xo)), "mc_rollout_true_model")
        return dict(J=J, stats=stats.reshape(P, 3), x=xo)

    def _gen(self, gen):
        gen = gen or {}
        ens = gen.get("ensemble_params")
        ensf = None if ens is None else _f64(ens)
        keep = [ensf]
        g = GenerativeDesc(int(gen.get("noise_kind", 0)), float(gen.get("noise_scale", 1.0)),
                           int(gen.get("n_ensemble", 1)), _dp(ensf), None, 0)
        if gen.get("use_true_model") and gen.get("true_model") is not None:
            mx, bufs = self._mixture(gen["true_model"])
            keep += [mx, bufs]
            g.true_model = C.pointer(mx)
            g.use_true_model = 1
        return g, keep

    def pets_costs(self, spec, x0, controls, particles, noise=None, seed=0, gen=None):
        """compute_cost_serial (pets.jl:128-157). controls (m, N, C)."""
        m, N = spec.m, spec.N
        controls = np.asarray(controls, dtype=np.float64).reshape(m, N, -1)
        Cn = controls.shape[2]
        cf, x0f = _f64(controls), _f64(x0)
        nf = None if noise is None else _f64(noise)
        cost = np.zeros(Cn)
        g, keep = self._gen(gen)
        d = spec.desc()
        self._check(self.f_pets_costs(self.ctx, C.byref(d), C.byref(g), _dp(x0f), _dp(cf), Cn, int(particles),
                                      _dp(nf), int(seed), _dp(cost)), "pets_costs")
        return cost

    def pets_refit(self, controls, cost, num_elite, smoothing, mu, Sigma):
        """get_elite_samples + compute_new_distribution (pets.jl:159-191)."""
        controls = np.asarray(controls, dtype=np.float64)
        m, N, Cn = controls.shape
        cf, costf = _f64(controls), _f64(cost)
        mu_o, Sg_o = _f64(mu).copy(), _f64(Sigma).copy()
        idx = np.zeros(num_elite, np.int32)
        self._check(self.f_pets_refit(self.ctx, m, N, Cn, int(num_elite), float(smoothing), _dp(cf), _dp(costf),
                                      _dp(mu_o), _dp(Sg_o), _ip(idx)), "pets_refit")
        return mu_o.reshape((m, N), order="F"), Sg_o.reshape((m, m, N), order="F"), idx

    def pets_solve(self, spec, x0, mu, Sigma, C_samples, particles, num_elite, iter_max, smoothing,
                   z_inject=None, noise=None, seed=0, gen=None):
        m, N = spec.m, spec.N
        x0f = _f64(x0)
        mu_o, Sg_o = _f64(mu).copy(), _f64(Sigma).copy()
        zf = None if z_inject is None else _f64(z_inject)
        nf = None if noise is None else _f64(noise)
        g, keep = self._gen(gen)
        d = spec.desc()
        self._check(self.f_pets_solve(self.ctx, C.byref(d), C.byref(g), _dp(x0f), int(C_samples), int(particles),
                                      int(num_elite), int(iter_max), float(smoothing), _dp(zf), _dp(nf), int(seed),
                                      _dp(mu_o), _dp(Sg_o)), "pets_solve")
        return mu_o.reshape((m, N), order="F"), Sg_o.reshape((m, m, N), order="F")


class Multi:
    """One process driving several GPUs (ratilqr_create_multi: contexts + NCCL communicators from ncclCommInitAll).
    ce_costs / ce_solve / pets_costs shard ONE population over the devices (all-gather of the cost vector inside the
    library, device buffers); ce_solve_fleet block-partitions independent problems (no collective)."""

    def __init__(self, cdll, device_ids):
        self.dll = cdll
        ids = (C.c_int32 * len(device_ids))(*[int(d) for d in device_ids])
        self.h = C.c_void_p()
        f = cdll.ratilqr_create_multi
        f.restype = C.c_int32
        f.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32]
        rc = f(C.byref(self.h), ids, len(device_ids))
        if rc:
            raise ApiError(f"ratilqr_create_multi failed ({rc}): needs {len(device_ids)} CUDA devices and NCCL; there is no CPU fallback")
        self.n_dev = len(device_ids)
        cdll.ratilqr_multi_last_error.restype = C.c_char_p
        cdll.ratilqr_multi_last_error.argtypes = [C.c_void_p]
        PD, IO, OUT, GD = C.POINTER(ProblemDesc), C.POINTER(IleqgOpts), C.POINTER(IleqgOut), C.POINTER(GenerativeDesc)
        dp, ip, i32, f64, vp = c_double_p, c_int32_p, C.c_int32, C.c_double, C.c_void_p

        def fn(name, at):
            g = getattr(cdll, "ratilqr_multi_" + name)
            g.restype, g.argtypes = C.c_int32, at
            return g
        self.f_costs = fn("ce_costs", [vp, PD, IO, dp, dp, dp, i32, f64, dp, ip])
        self.f_solve = fn("ce_solve", [vp, PD, IO, C.POINTER(CeOpts), dp, dp, f64, dp, C.c_int64, C.c_uint64, dp, dp, dp, dp, dp, dp,
                                        dp, dp, C.POINTER(C.c_int64), C.POINTER(i32), OUT])
        self.f_pets = fn("pets_costs", [vp, PD, GD, dp, dp, i32, i32, dp, C.c_uint64, dp])
        self.f_fleet = fn("ce_solve_fleet", [vp, PD, IO, C.POINTER(CeOpts), i32, dp, dp, i32, f64, C.c_uint64, dp, dp, dp, dp, dp])

    def close(self):
        if self.h:
            self.dll.ratilqr_destroy_multi.argtypes = [C.c_void_p]
            self.dll.ratilqr_destroy_multi(self.h)
            self.h = C.c_void_p()

    def _check(self, rc, what):
        if rc:
            raise ApiError(f"ratilqr_multi_{what} failed ({rc}): {self.dll.ratilqr_multi_last_error(self.h).decode()}")

    def ce_costs(self, spec, x0, u_init, theta, kl_bound, opts=None):
        opts = opts or make_opts()
        theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
        x0f, uf = _f64(x0), _f64(u_init)
        cost, status = np.empty(theta.size), np.empty(theta.size, np.int32)
        d = spec.desc()
        self._check(self.f_costs(self.h, C.byref(d), C.byref(opts), _dp(x0f), _dp(uf), _dp(theta), theta.size, float(kl_bound),
                                 _dp(cost), _ip(status)), "ce_costs")
        return cost, status

    def ce_solve(self, spec, x0, u_init, kl_bound, mu_init, sigma_init, num_samples=10, num_elite=3, iter_max=5, lam=0.5,
                 use_theta_max=False, z_inject=None, seed=0, opts=None):
        opts = opts or make_opts()
        n, m, N = spec.n, spec.m, spec.N
        x0f, uf = _f64(x0), _f64(u_init)
        zf, nz = None, 0
        if z_inject is not None:
            zf = np.ascontiguousarray(np.asarray(z_inject, dtype=np.float64).ravel())
            nz = zf.size
        sc = {k: np.zeros(1) for k in ("theta_opt", "value", "theta_min", "theta_max", "mu", "sigma")}
        mu_i, sg_i = np.array([float(mu_init)]), np.array([float(sigma_init)])
        nzu, st = np.zeros(1, np.int64), np.zeros(1, np.int32)
        x, l, L = np.zeros((n, N + 1), order="F"), np.zeros((m, N), order="F"), np.zeros((m, n, N), order="F")
        out = IleqgOut(_dp(x), _dp(l), _dp(L), None, _ip(st), None, None, None, None, None, None, 0)
        ce = CeOpts(int(num_samples), int(num_elite), int(iter_max), float(lam), int(bool(use_theta_max)))
        rounds = C.c_int32(0)
        d = spec.desc()
        self._check(self.f_solve(self.h, C.byref(d), C.byref(opts), C.byref(ce), _dp(x0f), _dp(uf), float(kl_bound), _dp(zf), nz,
                                 int(seed), _dp(mu_i), _dp(sg_i), _dp(sc["theta_opt"]), _dp(sc["value"]), _dp(sc["theta_min"]),
                                 _dp(sc["theta_max"]), _dp(sc["mu"]), _dp(sc["sigma"]), nzu.ctypes.data_as(C.POINTER(C.c_int64)),
                                 C.byref(rounds), C.byref(out)), "ce_solve")
        res = {k: float(v[0]) for k, v in sc.items()}
        res.update(mu_init=float(mu_i[0]), sigma_init=float(sg_i[0]), nz_used=int(nzu[0]), rounds=int(rounds.value), x=x, l=l, L=L)
        return res

    def pets_costs(self, spec, x0, controls, particles, noise=None, seed=0, gen=None):
        m, N = spec.m, spec.N
        cf = _f64(controls)
        Cn = cf.size // (m * N)
        nf = None if noise is None else _f64(noise)
        cost = np.zeros(Cn)
        gen = gen or {}
        ens = gen.get("ensemble_params")
        ensf = None if ens is None else _f64(ens)
        g = GenerativeDesc(int(gen.get("noise_kind", 0)), float(gen.get("noise_scale", 1.0)), int(gen.get("n_ensemble", 1)),
                           _dp(ensf), None, 0)
        x0f = _f64(x0)
        d = spec.desc()
        self._check(self.f_pets(self.h, C.byref(d), C.byref(g), _dp(x0f), _dp(cf), Cn, int(particles), _dp(nf), int(seed), _dp(cost)),
                    "pets_costs")
        return cost

    def ce_solve_fleet(self, spec, x0, u_init, kl_bound, mu_init, sigma_init, num_samples=10, num_elite=3, iter_max=5, lam=0.5,
                       use_theta_max=False, seed=0, opts=None):
        opts = opts or make_opts()
        n, m, N = spec.n, spec.m, spec.N
        x0 = np.asarray(x0, dtype=np.float64).reshape(n, -1)
        P = x0.shape[1]
        u_init = np.asarray(u_init, dtype=np.float64)
        u_count = 1 if u_init.ndim == 2 else u_init.shape[-1]
        x0f, uf = _f64(x0), _f64(u_init)
        mu_i = np.ascontiguousarray(np.broadcast_to(np.asarray(mu_init, np.float64), (P,))).copy()
        sg_i = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma_init, np.float64), (P,))).copy()
        th, val = np.zeros(P), np.zeros(P)
        l = np.zeros((m, N, P), order="F")
        ce = CeOpts(int(num_samples), int(num_elite), int(iter_max), float(lam), int(bool(use_theta_max)))
        d = spec.desc()
        self._check(self.f_fleet(self.h, C.byref(d), C.byref(opts), C.byref(ce), P, _dp(x0f), _dp(uf), u_count, float(kl_bound),
                                 int(seed), _dp(mu_i), _dp(sg_i), _dp(th), _dp(val), _dp(l)), "ce_solve_fleet")
        return dict(theta_opt=th, value=val, l=l, mu_init=mu_i, sigma_init=sg_i)
