"""Launch shape of k_ileqg_solve against batch size (RATILQR_SOLVE_SHAPE: 0 = 64x4 / 255 registers, 3 = 32x8 / 255, default =
128x3 / 168): single-problem theta batches and CE-round-shaped fleet batches (P problems x 10 theta).
    RATILQR_SOLVE_SHAPE=0 python scripts/shape_latency.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import ratilqr_b200 as R  # noqa: E402
from ratilqr_b200 import workloads as wl  # noqa: E402

be = R.new_backend(0)
prob, x0, u = wl.c2_problem()
spec = prob.spec()
out = {"shape": os.environ.get("RATILQR_SOLVE_SHAPE", "default")}
for K in (10, 1024, 8192):
    th = wl.c2_thetas(K)
    be.stage(spec, x0, u, th)
    be.run(2)
    out[f"single_K{K}_ms"] = round(be.run(5) / 5, 3)
for P in (1638, 3277, 4915, 5683, 8192, 16384):
    fprob, cps, fx0, fu = wl.fleet(P, key=70)
    fspec = fprob.spec(cost_params=cps)
    th = wl.positive_thetas(P * 10, key=3)
    be.ce_costs(fspec, fx0, fu, wl.positive_thetas(P * 10, key=4), 0.1, P=P)  # leaves the work profile
    be.stage(fspec, fx0, fu, th, P=P)
    be.run(1)
    out[f"fleet_P{P}x10_ms"] = round(be.run(3) / 3, 3)
print(json.dumps(out))
