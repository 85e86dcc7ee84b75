"""User-extensible device models: CUDA C++ snippets compiled at run time with NVRTC, differentiated by
forward-mode dual numbers on the device (the analogue of arbitrary Julia closures + ForwardDiff, src/ileqg.jl:265-273).

Three layers of checks:
  * CPU: the snippets of tests/user_models/ are compiled statically by g++ into the host emulation through the SAME
    adapters (rl_user.cuh) and compared with the oracle (a user-written unicycle + goal cost must reproduce the
    registered unicycle + QuadraticCost) and with complex-step / finite-difference derivatives of numpy restatements;
  * CPU: NVRTC compiles the snippets to sm_100a cubins (no GPU needed) and reports errors with the snippet's line;
  * GPU: the registered user models run through the C ABI and are compared with the oracle (where a registered
    equivalent exists) and with the g++ build of the same snippet (where none does) at 1e-9.
"""
import ctypes
import os

import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import _capi, _lib
from ratilqr_b200 import workloads as wl

RTOL = 1e-9  # north_star tolerance, as in test_gpu_parity.py


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    assert np.array_equal(fin, np.isfinite(a)), "non-finite pattern differs"
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin])) / max(np.max(np.abs(b[fin])), 1e-300))


HERE = os.path.dirname(os.path.abspath(__file__))


def snippet(name):
    return open(os.path.join(HERE, "user_models", name + ".inc")).read()


# hostemu-only numbering of the statically compiled snippets (tests/_hostemu/hostemu.cpp)
HM_UNICYCLE, HM_DRAGCAR, HM_GOAL, HM_OBSTACLE = 1000, 1001, 100, 101
HM_UNICYCLE_STRUCT, HM_GOAL_STRUCT = 1002, 102  # the same snippets with declared structure

# declared structure of the unicycle snippet and of the goal cost (include/ratilqr.h: a_kind ... p_kind)
UNI_A = np.array([[1, 0, 2, 2], [0, 1, 2, 2], [0, 0, 1, 0], [0, 0, 0, 1]])
UNI_B = np.array([[0, 0], [0, 0], [0, 2], [2, 0]])
DIAG_Q, DIAG_R, ZERO_P = 2 * np.eye(4, dtype=int), 2 * np.eye(2, dtype=int), np.zeros((2, 4), dtype=int)

QD, RD, XG, QF = np.array([1.0, 1.0, 0.1, 0.1]), np.array([0.1, 0.1]), np.array([5.0, 5.0, 0.0, 0.0]), 10.0
GOAL_CP = np.concatenate([0.01 * QD, 0.01 * RD, XG, [QF]])                       # goal_cost.inc
OBST_CP = np.concatenate([0.01 * QD, 0.01 * RD, XG, [QF], [2.5, 2.0, 0.8, 0.3]])  # obstacle_cost.inc
DRAG_P = np.array([0.1, 0.05, 1.5])                                             # drag_car_dynamics.inc


def c2_like(model_id, cost_id, cp, mp=(0.1,), N=50):
    prob, x0, u = wl.c2_problem(N=N)
    ref = prob.spec()
    return _capi.Spec(model_id, cost_id, 4, 2, N, np.asarray(mp, float), cp, ref.W.reshape(4, 4, order="F")), x0, u, ref


# ------------------------------------------------------------------------------------------------------
# numpy restatements of the snippets (complex-step friendly)
# ------------------------------------------------------------------------------------------------------
def np_dragcar(p, x, u):
    dt = p[0]
    v = x[3]
    absv = np.sqrt(v * v)  # analytic continuation of fabs for the complex step (v != 0 in the tests)
    return np.array([x[0] + dt * v * np.cos(x[2]), x[1] + dt * v * np.sin(x[2]), x[2] + dt * p[2] * np.tanh(u[1] / p[2]),
                     v + dt * (u[0] - p[1] * v * absv)])


def np_obstacle_stage(cp, k, x, u):
    c = 0.5 * cp[4] * u[0] ** 2 + 0.5 * cp[5] * (1.0 + 0.2 * x[3] ** 2) * u[1] ** 2
    for i in range(4):
        c = c + 0.5 * cp[i] * (x[i] - cp[6 + i]) ** 2
    bump = cp[14] * np.exp(-0.5 * ((x[0] - cp[11]) ** 2 + (x[1] - cp[12]) ** 2) / cp[13] ** 2)
    return c + bump + 0.01 * k * np.sqrt(1.0 + u[0] ** 2 + (x[3] * u[1]) ** 2)


def np_obstacle_terminal(cp, x):
    c = cp[14] * np.exp(-0.5 * ((x[0] - cp[11]) ** 2 + (x[1] - cp[12]) ** 2) / cp[13] ** 2)
    for i in range(4):
        c = c + 0.5 * cp[10] * cp[i] * (x[i] - cp[6 + i]) ** 2
    return c + np.log(1.0 + x[3] ** 2)


def cs_grad(fun, z, h=1e-30):
    g = np.zeros(z.size)
    for i in range(z.size):
        zc = z.astype(complex)
        zc[i] += 1j * h
        g[i] = np.imag(fun(zc)) / h
    return g


def cs_hess(fun, z, h=1e-5):
    """central differences of the complex-step gradient: error O(h^2) ~ 1e-10"""
    n = z.size
    H = np.zeros((n, n))
    for j in range(n):
        zp, zm = z.copy(), z.copy()
        zp[j] += h
        zm[j] -= h
        H[:, j] = (cs_grad(fun, zp) - cs_grad(fun, zm)) / (2 * h)
    return 0.5 * (H + H.T)


def check_linearize_against_numpy(be, spec, cp, mp):
    """approximate_model of (drag car, obstacle cost) against derivatives of the numpy restatements"""
    rng = np.random.default_rng(3)
    N = spec.N
    x = rng.standard_normal((4, N + 1)) + np.array([2.0, 2.0, 0.3, 1.5])[:, None]
    u = 0.5 * rng.standard_normal((2, N))
    r = be.linearize(spec, x, u)
    assert np.all(r["status"] == 0)
    for k in range(N):
        z = np.concatenate([x[:, k], u[:, k]])
        f = lambda zz: np_obstacle_stage(cp, k, zz[:4], zz[4:])
        g, H = cs_grad(f, z), cs_hess(f, z)
        assert abs(r["q"][k, 0] - f(z)) < 1e-12 * max(1.0, abs(f(z)))
        assert np.allclose(r["qv"][:, k, 0], g[:4], rtol=1e-10, atol=1e-12)
        assert np.allclose(r["r"][:, k, 0], g[4:], rtol=1e-10, atol=1e-12)
        assert np.allclose(r["Q"][:, :, k, 0], H[:4, :4], rtol=1e-6, atol=1e-8)
        assert np.allclose(r["R"][:, :, k, 0], H[4:, 4:], rtol=1e-6, atol=1e-8)
        assert np.allclose(r["P"][:, :, k, 0], H[4:, :4], rtol=1e-6, atol=1e-8)  # cux: (m, n)
        A = np.stack([cs_grad(lambda zz, i=i: np_dragcar(mp, zz[:4], zz[4:])[i], z) for i in range(4)])
        assert np.allclose(r["A"][:, :, k, 0], A[:, :4], rtol=1e-12, atol=1e-14)
        assert np.allclose(r["B"][:, :, k, 0], A[:, 4:], rtol=1e-12, atol=1e-14)
    ft = lambda zz: np_obstacle_terminal(cp, zz)
    assert np.allclose(r["qv"][:, N, 0], cs_grad(ft, x[:, N]), rtol=1e-10, atol=1e-12)
    assert np.allclose(r["Q"][:, :, N, 0], cs_hess(ft, x[:, N]), rtol=1e-6, atol=1e-8)


# ------------------------------------------------------------------------------------------------------
# CPU: the adapters + dual numbers, compiled by g++ (host emulation)
# ------------------------------------------------------------------------------------------------------
def test_hostemu_user_unicycle_matches_registered(hostemu_be, oracle_be):
    """user-written unicycle (Dual Jacobians) + registered QuadraticCost == registered unicycle, to 1e-9"""
    prob, x0, u = wl.c2_problem()
    ref = prob.spec()
    theta = wl.c2_thetas(48)
    spec_u = _capi.Spec(HM_UNICYCLE, ref.cost_id, 4, 2, ref.N, ref.model_params, ref.cost_params, ref.W.reshape(4, 4, order="F"))
    g = hostemu_be.ileqg_solve_batch(spec_u, x0, u, theta)
    o = oracle_be.ileqg_solve_batch(ref, x0, u, theta)
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["iters"], o["iters"])
    for k in ("value", "x", "l", "L"):
        assert relerr(g[k], o[k]) < RTOL, k


@pytest.mark.parametrize("model_id", [HM_UNICYCLE, R.models.MODEL_UNICYCLE])
def test_hostemu_user_goal_cost_matches_quadratic(hostemu_be, oracle_be, model_id):
    """user-written goal cost (second-order duals: cx, cxx, cu, cuu, cux) == registered QuadraticCost"""
    spec_u, x0, u, ref = c2_like(model_id, HM_GOAL, GOAL_CP)
    theta = wl.c2_thetas(32)
    g = hostemu_be.ileqg_solve_batch(spec_u, x0, u, theta)
    o = oracle_be.ileqg_solve_batch(ref, x0, u, theta)
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["iters"], o["iters"])
    for k in ("value", "x", "l", "L"):
        assert relerr(g[k], o[k]) < RTOL, k


@pytest.mark.parametrize("pair", [(HM_UNICYCLE_STRUCT, HM_GOAL_STRUCT), (HM_UNICYCLE_STRUCT, HM_GOAL), (HM_UNICYCLE, HM_GOAL_STRUCT)])
def test_hostemu_declared_structure_is_bit_identical(hostemu_be, pair):
    """skipping declared zeros / ones changes no finite result (fma(x, 0, acc) == acc): structured == dense, bit for bit"""
    theta = wl.c2_thetas(24)
    spec_d, x0, u, _ = c2_like(HM_UNICYCLE, HM_GOAL, GOAL_CP)
    spec_s, _, _, _ = c2_like(pair[0], pair[1], GOAL_CP)
    d = hostemu_be.ileqg_solve_batch(spec_d, x0, u, theta)
    g = hostemu_be.ileqg_solve_batch(spec_s, x0, u, theta)
    for k in ("status", "iters", "trials", "value", "x", "l", "L"):
        assert np.array_equal(d[k], g[k]), k


def test_hostemu_new_model_derivatives(hostemu_be):
    spec, _, _, _ = c2_like(HM_DRAGCAR, HM_OBSTACLE, OBST_CP, mp=DRAG_P, N=6)
    check_linearize_against_numpy(hostemu_be, spec, OBST_CP, DRAG_P)


def test_hostemu_new_model_solves(hostemu_be):
    """drag car + obstacle cost: the solve converges and improves on the initial rollout for every theta"""
    spec, x0, u, _ = c2_like(HM_DRAGCAR, HM_OBSTACLE, OBST_CP, mp=DRAG_P, N=40)
    theta = np.array([0.0, 0.5, 2.0, 5.0])
    g = hostemu_be.ileqg_solve_batch(spec, x0, u, theta)
    assert np.all(g["status"] == 0) and np.all(g["iters"] >= 2)
    x_init = hostemu_be.rollout_open(spec, x0, np.zeros((2, 40)))[0]
    j_init = hostemu_be.integrate_cost(spec, x_init, np.zeros((2, 40)))[0][0]
    j_opt = hostemu_be.integrate_cost(spec, g["x"][..., 0], g["l"][..., 0])[0][0]
    assert j_opt < 0.9 * j_init


# ------------------------------------------------------------------------------------------------------
# CPU: NVRTC compiles the snippets for sm_100a (no GPU needed)
# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def lib_api():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libratilqr_b200.so not built")
    return _capi.CApi(ctypes.CDLL(_lib.LIB_PATH), "ratilqr_", needs_ctx=True)


def _check(api, **kw):
    rc, log = api.user_model_check(**kw)
    if rc == -20:
        pytest.skip("libnvrtc is not loadable here: " + log)
    return rc, log


def test_nvrtc_compiles_user_pair(lib_api):
    rc, log = _check(lib_api, n=4, m=2, dynamics_src=snippet("drag_car_dynamics"), n_model_params=3,
                     cost_src=snippet("obstacle_cost"), n_cost_params=15)
    assert rc == 0, log


def test_nvrtc_reports_snippet_errors(lib_api):
    bad = snippet("unicycle_dynamics").replace("cos(x[2])", "cosine(x[2])")
    rc, log = _check(lib_api, n=4, m=2, dynamics_src=bad, n_model_params=1, base_cost_id=1)
    assert rc == -21 and "cosine" in log and "dynamics(" in log  # file name + line of the snippet
    rc, log = lib_api.user_model_check(n=4, m=2)
    assert rc == -1 and "neither" in log
    rc, log = lib_api.user_model_check(n=40, m=2, dynamics_src="x", base_cost_id=1)
    assert rc == -1
    rc, log = lib_api.user_model_check(n=4, m=2, cost_src=snippet("goal_cost"), base_model_id=R.models.MODEL_CARTPOLE, n_cost_params=11)
    assert rc == -1 and "registered model" in log  # cart-pole is (4, 1)
    rc, log = lib_api.user_model_check(n=4, m=2, dynamics_src=snippet("unicycle_dynamics"), n_model_params=1, base_cost_id=1,
                                       a_kind=3 * np.ones((4, 4), int))
    assert rc == -1 and "a_kind" in log
    rc, log = lib_api.user_model_check(n=4, m=2, dynamics_src=snippet("unicycle_dynamics"), n_model_params=1,
                                       cost_src=snippet("goal_cost"), n_cost_params=11, q_kind=np.ones((4, 4), int))
    assert rc == -1 and "q_kind" in log


def test_nvrtc_compiles_declared_structure(lib_api):
    rc, log = _check(lib_api, n=4, m=2, dynamics_src=snippet("unicycle_dynamics"), n_model_params=1,
                     cost_src=snippet("goal_cost"), n_cost_params=11, a_kind=UNI_A, b_kind=UNI_B, q_kind=DIAG_Q, r_kind=DIAG_R,
                     p_kind=ZERO_P)
    assert rc == 0, log


# ------------------------------------------------------------------------------------------------------
# GPU: registered user models through the C ABI
# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu_user(gpu_be):
    ids = {}
    ids["uni+quad"] = gpu_be.user_model_register(4, 2, dynamics_src=snippet("unicycle_dynamics"), n_model_params=1, base_cost_id=1)
    ids["uni+goal"] = gpu_be.user_model_register(4, 2, dynamics_src=snippet("unicycle_dynamics"), n_model_params=1,
                                                 cost_src=snippet("goal_cost"), n_cost_params=11)
    ids["reg+goal"] = gpu_be.user_model_register(4, 2, base_model_id=R.models.MODEL_UNICYCLE, cost_src=snippet("goal_cost"), n_cost_params=11)
    ids["drag+obst"] = gpu_be.user_model_register(4, 2, dynamics_src=snippet("drag_car_dynamics"), n_model_params=3,
                                                  cost_src=snippet("obstacle_cost"), n_cost_params=15)
    return ids


@pytest.mark.gpu
def test_gpu_user_unicycle_vs_oracle(gpu_be, oracle_be, gpu_user):
    prob, x0, u = wl.c2_problem()
    ref = prob.spec()
    theta = wl.c2_thetas(256)
    spec_u = _capi.Spec(gpu_user["uni+quad"], 1, 4, 2, ref.N, ref.model_params, ref.cost_params, ref.W.reshape(4, 4, order="F"))
    g = gpu_be.ileqg_solve_batch(spec_u, x0, u, theta)
    o = oracle_be.ileqg_solve_batch(ref, x0, u, theta)
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["iters"], o["iters"])
    assert np.all(o["status"] == 0)
    for k in ("value", "x", "l", "L"):
        assert relerr(g[k], o[k]) < RTOL, k


@pytest.mark.gpu
@pytest.mark.parametrize("pair", ["uni+goal", "reg+goal"])
def test_gpu_user_goal_cost_vs_oracle(gpu_be, oracle_be, gpu_user, pair):
    spec_u, x0, u, ref = c2_like(gpu_user[pair], _capi.COST_USER, GOAL_CP)
    theta = wl.c2_thetas(128)
    g = gpu_be.ileqg_solve_batch(spec_u, x0, u, theta)
    o = oracle_be.ileqg_solve_batch(ref, x0, u, theta)
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["iters"], o["iters"])
    for k in ("value", "x", "l", "L"):
        assert relerr(g[k], o[k]) < RTOL, k
    # the bilevel layer on top: compute_cost for the same population
    cg = gpu_be.ce_costs(spec_u, x0, u, theta, 0.1)[0]
    co = oracle_be.ce_costs(ref, x0, u, theta, 0.1)[0]
    assert relerr(cg, co) < RTOL


@pytest.mark.gpu
def test_gpu_declared_structure(gpu_be, oracle_be, gpu_user):
    """a true declaration: bit-identical to the dense registration and 1e-9 from the oracle; a false one is refused"""
    mid = gpu_be.user_model_register(4, 2, dynamics_src=snippet("unicycle_dynamics"), n_model_params=1, cost_src=snippet("goal_cost"),
                                     n_cost_params=11, a_kind=UNI_A, b_kind=UNI_B, q_kind=DIAG_Q, r_kind=DIAG_R, p_kind=ZERO_P)
    theta = wl.c2_thetas(128)
    spec_s, x0, u, ref = c2_like(mid, _capi.COST_USER, GOAL_CP)
    spec_d, _, _, _ = c2_like(gpu_user["uni+goal"], _capi.COST_USER, GOAL_CP)
    s = gpu_be.ileqg_solve_batch(spec_s, x0, u, theta)
    d = gpu_be.ileqg_solve_batch(spec_d, x0, u, theta)
    o = oracle_be.ileqg_solve_batch(ref, x0, u, theta)
    for k in ("status", "iters", "trials", "value", "x", "l", "L"):
        assert np.array_equal(s[k], d[k]), k
    assert np.array_equal(s["iters"], o["iters"])
    for k in ("value", "x", "l", "L"):
        assert relerr(s[k], o[k]) < RTOL, k
    wrong_a = UNI_A.copy()
    wrong_a[0, 2] = 0  # d(px')/d(psi) = -dt v sin(psi) is not zero
    with pytest.raises(_capi.ApiError, match=r"a_kind\[0,2\] is declared zero"):
        gpu_be.user_model_register(4, 2, dynamics_src=snippet("unicycle_dynamics"), n_model_params=1, base_cost_id=1, a_kind=wrong_a)
    wrong_p = DIAG_Q.copy()
    wrong_p[1, 1] = 0  # cxx[1,1] = Q_11 is not zero
    with pytest.raises(_capi.ApiError, match=r"q_kind"):
        gpu_be.user_model_register(4, 2, base_model_id=R.models.MODEL_UNICYCLE, cost_src=snippet("goal_cost"), n_cost_params=11,
                                   q_kind=wrong_p)


@pytest.mark.gpu
def test_gpu_new_model_vs_host_build_of_same_snippet(gpu_be, hostemu_be, gpu_user):
    """no registered equivalent exists: NVRTC build on the device vs g++ build of the same text, discrete path included"""
    N = 40
    spec_g, x0, u, _ = c2_like(gpu_user["drag+obst"], _capi.COST_USER, OBST_CP, mp=DRAG_P, N=N)
    spec_h, _, _, _ = c2_like(HM_DRAGCAR, HM_OBSTACLE, OBST_CP, mp=DRAG_P, N=N)
    theta = np.concatenate([[0.0], wl.positive_thetas(63, key=3)])
    g = gpu_be.ileqg_solve_batch(spec_g, x0, u, theta, eps_hist_cap=128)
    h = hostemu_be.ileqg_solve_batch(spec_h, x0, u, theta, eps_hist_cap=128)
    for k in ("status", "iters", "trials", "restarts"):
        assert np.array_equal(g[k], h[k]), k
    ok = h["status"] == 0
    assert ok.sum() >= 32
    for k in ("value", "x", "l", "L", "eps_hist"):
        assert relerr(g[k][..., ok], h[k][..., ok]) < RTOL, k
    check_linearize_against_numpy(gpu_be, c2_like(gpu_user["drag+obst"], _capi.COST_USER, OBST_CP, mp=DRAG_P, N=6)[0], OBST_CP, DRAG_P)


@pytest.mark.gpu
def test_gpu_user_components_and_mc(gpu_be, hostemu_be, oracle_be, gpu_user):
    """rollouts, integrate_cost, Monte Carlo (injected noise) and PETS rollouts of a user pair"""
    N = 20
    spec_g, x0, u, ref = c2_like(gpu_user["uni+goal"], _capi.COST_USER, GOAL_CP, N=N)
    rng = np.random.default_rng(9)
    B = 5
    uu = 0.3 * rng.standard_normal((2, N, B))
    xx0 = np.tile(x0[:, None], (1, B))
    xo = gpu_be.rollout_open(spec_g, xx0, uu)[0]
    xr = oracle_be.rollout_open(ref, xx0, uu)[0]
    assert relerr(xo, xr) < 1e-13
    cg = gpu_be.integrate_cost(spec_g, xo, uu)[0]
    cr = oracle_be.integrate_cost(ref, xr, uu)[0]
    assert relerr(cg, cr) < 1e-12
    sol = oracle_be.ileqg_solve_batch(ref, x0, u, [1.0])
    noise = 0.05 * rng.standard_normal((4, N, 64))
    mg = gpu_be.mc_rollout(spec_g, sol["x"][..., 0], sol["l"][..., 0], sol["L"][..., 0], 64, noise=noise, theta_risk=0.5)
    mo = oracle_be.mc_rollout(ref, sol["x"][..., 0], sol["l"][..., 0], sol["L"][..., 0], 64, noise=noise, theta_risk=0.5)
    assert relerr(mg["J"], mo["J"]) < RTOL and relerr(mg["stats"], mo["stats"]) < RTOL
    ctr = 0.3 * rng.standard_normal((2, N, 7))
    pn = 0.02 * rng.standard_normal((4, N, 6, 7))
    pg = gpu_be.pets_costs(spec_g, x0, ctr, 6, noise=pn)
    po = oracle_be.pets_costs(ref, x0, ctr, 6, noise=pn)
    assert relerr(pg, po) < RTOL


@pytest.mark.gpu
def test_gpu_user_domain_error_and_misuse(gpu_be, gpu_user):
    mid = gpu_be.user_model_register(2, 2, dynamics_src=snippet("sqrt_dynamics"), n_model_params=1, base_cost_id=1)
    cost = R.QuadraticCost(2, 2, Q=np.eye(2), R=np.eye(2), Qf=np.eye(2), xg=[1.0, 0.0])  # x = (1, 0) is a fixed point
    spec = _capi.Spec(mid, 1, 2, 2, 5, [0.1], cost.params(), 1e-3 * np.eye(2))
    x0 = np.array([[1.0, 0.0], [-1.0, 0.0]]).T  # second instance: sqrt(-1) -> DomainError
    _, st = gpu_be.rollout_open(spec, x0, np.zeros((2, 5, 2)))
    assert list(st) == [0, 3]  # RATILQR_ST_DOMAIN
    g = gpu_be.ileqg_solve_batch(spec, x0, np.zeros((2, 5)), [0.1, 0.1], P=2)
    assert g["status"][0] == 0 and g["status"][1] == 3 and np.isinf(g["value"][1])
    with pytest.raises(_capi.ApiError, match="unknown user model"):
        gpu_be.rollout_open(_capi.Spec(1999, 1, 2, 2, 5, [0.1], cost.params(), 1e-3 * np.eye(2)), x0, np.zeros((2, 5, 2)))
    with pytest.raises(_capi.ApiError, match="model parameters"):
        gpu_be.rollout_open(_capi.Spec(mid, 1, 2, 2, 5, [0.1, 0.2], cost.params(), 1e-3 * np.eye(2)), x0, np.zeros((2, 5, 2)))
    with pytest.raises(_capi.ApiError, match="cost_id"):
        gpu_be.rollout_open(_capi.Spec(mid, _capi.COST_USER, 2, 2, 5, [0.1], cost.params(), 1e-3 * np.eye(2)), x0, np.zeros((2, 5, 2)))
    with pytest.raises(_capi.ApiError, match="cosine"):
        gpu_be.user_model_register(4, 2, dynamics_src=snippet("unicycle_dynamics").replace("cos(", "cosine("), n_model_params=1, base_cost_id=1)


@pytest.mark.gpu
def test_gpu_user_model_through_reference_api(gpu_be):
    """the host mirror of the reference API with a user model: problem struct -> ILEQGSolver / RAT iLQR solve_"""
    f = R.UserDynamics(4, 2, snippet("drag_car_dynamics"), params=DRAG_P, py=np_dragcar)
    cost = R.UserCost(snippet("obstacle_cost"), params=OBST_CP, stage_py=np_obstacle_stage, terminal_py=np_obstacle_terminal)
    fb = R.register_user_model(gpu_be, f, cost)
    assert fb.model_id >= _capi.MODEL_USER_BASE
    W = np.diag([1e-2, 1e-2, 1e-3, 1e-2]) * 0.1
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(fb, cost.c, cost.h, R.ConstantCovariance(W), 30)
    x0, u0 = np.array([0.0, 0.0, 0.0, 1.0]), [np.zeros(2) for _ in range(30)]
    s = R.ILEQGSolver(prob, backend=gpu_be)
    x, l, L, value, eps = R.ileqg.solve_(s, prob, x0, u0, theta=1.0, verbose=False)
    assert np.isfinite(value) and len(x) == 31 and len(L) == 30
    # the returned trajectory is the host callable's rollout of the returned controls
    xs = [x0]
    for k in range(30):
        xs.append(fb(xs[-1], l[k]))
    assert relerr(np.stack(xs), np.stack(x)) < 1e-10
    ce = R.CrossEntropyBilevelOptimizationSolver(num_samples=8, num_elite=3, iter_max=2, backend=gpu_be)
    out = R.cross_entropy.solve_(ce, prob, x0, u0, np.random.default_rng(1), kl_bound=0.1, verbose=False)
    assert out[0] > 0 and np.isfinite(out[4])


def _golden_user_case():
    g = np.load(os.path.join(HERE, "golden", "user_dragcar_obstacle.npz"))
    return g, (4, 2, 30, g["model_params"], g["cost_params"], g["W"])


def _check_against_golden(r, g):
    for k in ("status", "iters", "trials", "restarts"):
        assert np.array_equal(r[k], g[k]), k
    ok = g["status"] == 0
    assert ok.sum() >= 8
    for k in ("value", "x", "l", "L"):
        assert relerr(r[k][..., ok], g[k][..., ok]) < RTOL, k


def test_hostemu_user_model_golden(hostemu_be):
    """frozen result of the drag-car + obstacle-cost solve (tests/golden/make_goldens.py): pins the dual-number adapters"""
    g, (n, m, N, mp, cp, W) = _golden_user_case()
    r = hostemu_be.ileqg_solve_batch(_capi.Spec(HM_DRAGCAR, HM_OBSTACLE, n, m, N, mp, cp, W), g["x0"], g["u"], g["theta"])
    _check_against_golden(r, g)


@pytest.mark.gpu
def test_gpu_user_model_golden(gpu_be, gpu_user):
    g, (n, m, N, mp, cp, W) = _golden_user_case()
    r = gpu_be.ileqg_solve_batch(_capi.Spec(gpu_user["drag+obst"], _capi.COST_USER, n, m, N, mp, cp, W), g["x0"], g["u"], g["theta"])
    _check_against_golden(r, g)
