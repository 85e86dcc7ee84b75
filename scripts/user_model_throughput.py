"""Throughput of an NVRTC-compiled user model against the registered model it restates (unicycle + goal cost, C2 fleet
shape): what the generic path (dense structure, dual-number derivatives, 64x4 launch shape) costs.
    python scripts/user_model_throughput.py [problems]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ratilqr_b200 as R  # noqa: E402
from ratilqr_b200 import _capi  # noqa: E402
from ratilqr_b200 import workloads as wl  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 222
K = 1024
be = R.new_backend(0)
snip = lambda n: open(os.path.join(ROOT, "tests", "user_models", n + ".inc")).read()
prob, cps, x0, u = wl.fleet(P, key=7)
ref = prob.spec(cost_params=cps)
theta = np.concatenate([wl.positive_thetas(K, key=100 + p) for p in range(P)])
t0 = time.perf_counter()
mid = be.user_model_register(4, 2, dynamics_src=snip("unicycle_dynamics"), n_model_params=1, base_cost_id=1)
t_reg = time.perf_counter() - t0
mid2 = be.user_model_register(4, 2, dynamics_src=snip("unicycle_dynamics"), n_model_params=1, cost_src=snip("goal_cost"), n_cost_params=11)
UNI_A = np.array([[1, 0, 2, 2], [0, 1, 2, 2], [0, 0, 1, 0], [0, 0, 0, 1]])
UNI_B = np.array([[0, 0], [0, 0], [0, 2], [2, 0]])
mid3 = be.user_model_register(4, 2, dynamics_src=snip("unicycle_dynamics"), n_model_params=1, cost_src=snip("goal_cost"), n_cost_params=11,
                              a_kind=UNI_A, b_kind=UNI_B, q_kind=2 * np.eye(4, dtype=int), r_kind=2 * np.eye(2, dtype=int),
                              p_kind=np.zeros((2, 4), dtype=int))
mid4 = be.user_model_register(4, 2, dynamics_src=snip("unicycle_dynamics"), n_model_params=1, base_cost_id=1, a_kind=UNI_A, b_kind=UNI_B)
W = ref.W.reshape(4, 4, order="F")
# goal_cost.inc parameters per problem: [Qdiag, Rdiag, xg, qf] from the registered quadratic blocks
cp = ref.cost_params.reshape(P, -1)
n, m = 4, 2
oq, orr, oqf = 5 + n, 5 + n + n * n, 5 + n + n * n + m * m + n * m
Q = cp[:, oq:oq + 16].reshape(P, 4, 4); Rm = cp[:, orr:orr + 4].reshape(P, 2, 2); Qf = cp[:, oqf:oqf + 16].reshape(P, 4, 4)
gcp = np.concatenate([np.diagonal(Q, axis1=1, axis2=2), np.diagonal(Rm, axis1=1, axis2=2), cp[:, 5:9],
                      (Qf[:, 0, 0] / Q[:, 0, 0])[:, None]], axis=1)
specs = {"registered unicycle + QuadraticCost (structure-specialised kernel)": ref,
         "user unicycle snippet + registered QuadraticCost": _capi.Spec(mid, 1, 4, 2, ref.N, ref.model_params, cp, W),
         "user unicycle snippet + user goal-cost snippet (second-order duals)": _capi.Spec(mid2, _capi.COST_USER, 4, 2, ref.N, ref.model_params, gcp, W),
         "user unicycle snippet with declared structure + registered QuadraticCost": _capi.Spec(mid4, 1, 4, 2, ref.N, ref.model_params, cp, W),
         "both snippets with declared structure (a/b/q/r/p kinds)": _capi.Spec(mid3, _capi.COST_USER, 4, 2, ref.N, ref.model_params, gcp, W)}
base = None
for name, spec in specs.items():
    be.stage(spec, x0, u, theta, P=P)
    be.run(1)
    ms = be.run(3) / 3
    r = be.fetch()
    if base is None:
        base = r
    fin = np.isfinite(base["value"])
    print(json.dumps({"pair": name, "problems": P, "solves": P * K, "ms": ms, "solves_per_s": P * K / ms * 1e3,
                      "max_rel_value_diff_vs_registered": float(np.max(np.abs(r["value"][fin] - base["value"][fin]) / np.abs(base["value"][fin]))),
                      "same_status": bool(np.array_equal(r["status"], base["status"])),
                      "same_iters": bool(np.array_equal(r["iters"], base["iters"])), "nvrtc_register_s": round(t_reg, 2)}), flush=True)
