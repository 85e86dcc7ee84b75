"""Freeze the synthetic inputs of BASELINE.json's configs as JSON descriptors under bench/configs/.
The arrays come from ratilqr_b200.workloads (numpy Philox generators, fixed keys), so CPU oracle, host emulation and GPU
all see the same inputs; the descriptors record the parameters and a checksum of every generated array.
    python scripts/export_configs.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ratilqr_b200 import workloads as wl  # noqa: E402

OUT = os.path.join(ROOT, "bench", "configs")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes()).hexdigest()[:16]


def spec_desc(spec):
    return dict(model_id=spec.model_id, cost_id=spec.cost_id, n=spec.n, m=spec.m, N=spec.N, model_params=spec.model_params.tolist(),
                n_cost_params=spec.n_cost_params, cost_params_count=spec.cost_params_count, cost_params_sha256_16=sha(spec.cost_params),
                W=spec.W.reshape(spec.n, spec.n, order="F").tolist() if spec.W.size == spec.n * spec.n else "time-varying",
                cost_params_layout="[ws0, ws1, c0, c1, h0, xg(n), Q(n*n), R(m*m), Pc(n*m), Qf(n*n)] (include/ratilqr.h)" if spec.cost_id == 1
                else "[p, h0]")


def dump(name, d):
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(d, f, indent=1)
    print("wrote", name)


prob, x0, u = wl.c1_problem()
dump("c1_shipped_power_law", dict(
    source="test/ileqg_test.jl:151-155, test/cross_entropy_bilevel_optimization_test.jl:13-21", generator="workloads.c1_problem()",
    problem=spec_desc(prob.spec()), x0=x0.tolist(), u_init="0.1 * ones(2) per stage", thetas=[0.0, 0.1, 0.3, 0.43, 0.5],
    ce=dict(num_samples=3, kl_bound=1.0), nm=dict(iter_max=20, eps=1e-3, theta_high_init=10.0, theta_low_init=1e-8, kl_bound=1.0)))

prob, x0, u = wl.c2_problem()
th = wl.c2_thetas(1024)
dump("c2_unicycle_1024_thetas", dict(
    generator="workloads.c2_problem(), workloads.c2_thetas(1024)", problem=spec_desc(prob.spec()), x0=x0.tolist(), u_init="zeros",
    thetas="first 1024 positive draws of 1 + 2 z, z ~ numpy Philox(key=20201028)", thetas_sha256_16=sha(th), thetas_head=th[:4].tolist(),
    kl_bound=0.1, note="cost = SURVEY's C2 cost scaled by 0.01 (unscaled, 94 % of the theta population is infeasible in initialize!)",
    bench_default="bench.py replicates this shape over 1,776 problems per GPU: workloads.fleet(P, key=7 + 1000 rank) with "
                  "theta keys 20201028 + p + 100000 rank"))

prob, x0, u = wl.c3_problem()
dump("c3_quadrotor_nelder_mead", dict(
    generator="workloads.c3_problem()", problem=spec_desc(prob.spec()), x0=x0.tolist(), u_init="hover thrust m*g, zero torques",
    nm="defaults of nelder_mead_bilevel_optimization.jl:112-119", kl_bound=0.1, mc_rollouts_per_theta=256, mc_noise="Philox, chol(W) z"))

prob, x0 = wl.c4_problem()
fs = prob.f_stochastic
dump("c4_cartpole_pets", dict(
    generator="workloads.c4_problem()", problem=spec_desc(prob.spec()), x0=x0.tolist(), num_control_samples=4096, num_trajectory_samples=150,
    num_elite=409, iter_max=5, smoothing_factor=0.1, ensemble_params=fs.ensemble_params.tolist(),
    ensemble="5 parameter sets (cart / pole mass +- 5 %, numpy Philox key 4) x 30 particles", noise="additive Gaussian, W above"))

P = 8192
prob, cps, x0, u = wl.fleet(P, key=70)
dump("c5_unicycle_fleet", dict(
    generator="workloads.fleet(8192, key=70 + rank) per GPU (65,536 problems over 8 GPUs)", problem=spec_desc(prob.spec(cost_params=cps)),
    x0="px, py ~ U(-1, 1), psi ~ U(-0.5, 0.5), v ~ U(0.5, 1.5)", goals="px, py ~ U(3, 6)", x0_sha256_16=sha(x0), cost_params_sha256_16=sha(cps),
    ce=dict(num_samples=10, num_elite=3, iter_max=5, mu_init=1.0, sigma_init=2.0, kl_bound=0.1, rng="on-device Philox, seed 7 + rank"),
    mc_samples_per_problem=256))
