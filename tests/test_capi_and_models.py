"""CPU-only checks: the product library loads and exports every symbol include/ratilqr.h declares (no compute
call without a GPU), fails loudly without a device, and the model registry is consistent across its three
statements (numpy in models.py, oracle C++ with duals, device analytic Jacobians via the g++ build)."""
import ctypes
import os
import re

import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import models as M
from ratilqr_b200 import workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ratilqr.h")).read()
    names = sorted(set(re.findall(r"\b(ratilqr_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    dll = R.load_library()
    for nme in names:
        assert hasattr(dll, nme), f"{nme} declared in include/ratilqr.h but not exported"


def test_static_queries_need_no_gpu():
    dll = R.load_library()
    n, m, npar = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    for mid, (en, em, enp, _) in M._DIMS.items():
        assert dll.ratilqr_model_dims(mid, ctypes.byref(n), ctypes.byref(m), ctypes.byref(npar)) == 0
        assert (n.value, m.value, npar.value) == (en, em, enp)
    assert dll.ratilqr_model_dims(99, ctypes.byref(n), ctypes.byref(m), ctypes.byref(npar)) != 0
    assert dll.ratilqr_cost_param_count(1, 4, 2) == wl.unicycle_cost().params().size
    assert dll.ratilqr_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(R.ApiError):
        R.new_backend(0)  # ratilqr_create refuses: the product never computes on the CPU


@pytest.mark.parametrize("mid", sorted(M._DIMS))
def test_jacobians_three_ways(mid, oracle_be, hostemu_be):
    n, m, _, p = M._DIMS[mid]
    rng = np.random.default_rng(mid)
    x = np.abs(rng.standard_normal(n)) * 0.5 + 0.1
    u = np.abs(rng.standard_normal(m)) * 0.5 + 0.1
    A_cs, B_cs = M.jacobians_complex_step(mid, np.asarray(p, float), x, u)
    cost = R.PowerLawCost() if mid == M.MODEL_POWER_LAW else R.QuadraticCost(n, m, Q=np.eye(n), R=np.eye(m), Qf=np.eye(n))
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(M.DeviceDynamics(mid), cost.c, cost.h,
                                                             R.ConstantCovariance(np.eye(n)), 1)
    xs = np.stack([x, x], axis=1)
    for be in (oracle_be, hostemu_be):
        lin = be.linearize(prob.spec(), xs, u[:, None])
        assert np.allclose(lin["A"][..., 0, 0], A_cs, rtol=1e-12, atol=1e-14)
        assert np.allclose(lin["B"][..., 0, 0], B_cs, rtol=1e-12, atol=1e-14)
        xo, st = be.rollout_open(prob.spec(), x, u[:, None])
        assert np.allclose(xo[:, 1, 0], M.dynamics_numpy(mid, np.asarray(p, float), x, u), rtol=1e-14)


def test_cost_derivatives_against_finite_differences(oracle_be):
    rng = np.random.default_rng(3)
    n, m = 4, 2
    a = rng.standard_normal((n, n)); Q = a @ a.T
    b = rng.standard_normal((m, m)); Rm = b @ b.T + np.eye(m)
    cost = R.QuadraticCost(n, m, Q=Q, R=Rm, Qf=2 * Q, xg=rng.standard_normal(n), Pc=rng.standard_normal((n, m)),
                           ws0=0.5, ws1=0.25, c0=0.3, c1=0.1, h0=0.7)
    prob = R.FiniteHorizonRiskSensitiveOptimalControlProblem(R.Unicycle(), cost.c, cost.h, R.ConstantCovariance(np.eye(n)), 3)
    x = rng.standard_normal((n, 4)); u = rng.standard_normal((m, 3))
    lin = oracle_be.linearize(prob.spec(), x, u)
    h = 1e-6
    for k in range(3):
        c0 = cost.stage(k, x[:, k], u[:, k])
        assert np.isclose(lin["q"][k, 0], c0, rtol=1e-13)
        gx = np.array([(cost.stage(k, x[:, k] + h * e, u[:, k]) - cost.stage(k, x[:, k] - h * e, u[:, k])) / (2 * h) for e in np.eye(n)])
        gu = np.array([(cost.stage(k, x[:, k], u[:, k] + h * e) - cost.stage(k, x[:, k], u[:, k] - h * e)) / (2 * h) for e in np.eye(m)])
        assert np.allclose(lin["qv"][:, k, 0], gx, rtol=1e-6, atol=1e-8) and np.allclose(lin["r"][:, k, 0], gu, rtol=1e-6, atol=1e-8)
        w = cost.ws0 + cost.ws1 * k
        assert np.allclose(lin["Q"][..., k, 0], w * Q) and np.allclose(lin["R"][..., k, 0], w * Rm)
        assert np.allclose(lin["P"][..., k, 0], w * cost.Pc.T)  # P = d(grad_u c)/dx is m x n (ileqg.jl:269)
    assert np.isclose(lin["q"][3, 0], cost.terminal(x[:, 3])) and np.allclose(lin["Q"][..., 3, 0], 2 * Q)


def test_scalar_leqg_closed_form(oracle_be):
    """SURVEY.md 8c extra known answer: n = m = 1 LEQG, S~ = S/(1 - theta W S). Uses the (2,1) Riccati with a
    decoupled second state to embed the scalar problem."""
    n, m, N, B = 2, 1, 1, 1
    a, b, q, r, Sf, W, theta = 0.9, 0.5, 1.3, 0.7, 2.0, 0.05, 0.8
    lin = dict(q=np.zeros((N + 1, B)), qv=np.zeros((n, N + 1, B)), Q=np.zeros((n, n, N + 1, B)), r=np.zeros((m, N, B)),
               R=np.full((m, m, N, B), r), P=np.zeros((m, n, N, B)), A=np.zeros((n, n, N, B)), B=np.zeros((n, m, N, B)))
    lin["Q"][0, 0, 0, 0] = q; lin["Q"][0, 0, 1, 0] = Sf; lin["Q"][1, 1, :, 0] = 1.0
    lin["A"][0, 0, 0, 0] = a; lin["A"][1, 1, 0, 0] = 1.0; lin["B"][0, 0, 0, 0] = b
    out = oracle_be.riccati(lin, np.diag([W, 1e-6]), np.array([theta]), True)
    St = Sf / (1 - theta * W * Sf)
    Lc = -(b * St * a) / (r + b * St * b)
    S0 = q + a * St * a - (a * St * b) ** 2 / (r + b * St * b)
    assert np.isclose(out["L"][0, 0, 0, 0], Lc, rtol=1e-10) and np.isclose(out["S"][0, 0, 0, 0], S0, rtol=1e-10)


def test_multi_gpu_handle_fails_loudly_without_devices():
    """ratilqr_create_multi needs CUDA devices (and NCCL for more than one): no silent CPU path"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by scripts/check_multi_gpu.py")
    with pytest.raises(R.ApiError, match="ratilqr_create_multi failed"):
        R.new_multi([0, 1])
