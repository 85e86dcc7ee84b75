"""Latency regime A/B: the speculative kernel (RATILQR_SPEC = lanes per instance; 0 = one thread per instance) on the
small-batch cases of BASELINE.json: configs[1] exactly (1 x 1024 theta), a single problem's CE round (10 theta), the
single-problem RAT iLQR MPC step (host CE loop and the on-device loop with P = 1), against the CPU oracle.
    python scripts/latency_spec_ab.py > gpurun_out/latency_spec_ab.jsonl"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import ratilqr_b200 as R  # noqa: E402
from ratilqr_b200 import cross_entropy as CE  # noqa: E402
from ratilqr_b200 import workloads as wl  # noqa: E402

be = R.new_backend(0)
o = oracle.load()
prob, x0, u = wl.c2_problem()
spec = prob.spec()
ua = [u[:, k].copy() for k in range(u.shape[1])]


def timed(fn, reps=5):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return min(t) * 1e3


for mode in ("0", "2", "4", "8", "default"):
    if mode == "default":
        os.environ.pop("RATILQR_SPEC", None)
    else:
        os.environ["RATILQR_SPEC"] = mode
    out = {"RATILQR_SPEC": mode}
    for K in (10, 128, 1024, 2048, 4096):
        th = wl.c2_thetas(K)
        be.stage(spec, x0, u, th)
        be.run(2)
        out[f"kernel_K{K}_ms"] = round(be.run(5) / 5, 3)
    th = wl.c2_thetas(1024)
    out["e2e_ce_costs_1x1024_ms"] = round(timed(lambda: be.ce_costs(spec, x0, u, th, 0.1)), 3)
    out["e2e_solve_batch_1x1024_xlL_ms"] = round(timed(lambda: be.ileqg_solve_batch(spec, x0, u, th)), 3)

    def mpc_host():
        s = R.CrossEntropyBilevelOptimizationSolver(backend=be)
        return CE.solve_(s, prob, x0, ua, np.random.default_rng(1), kl_bound=0.1)

    out["mpc_step_host_loop_ms"] = round(timed(mpc_host, 3), 3)
    out["mpc_step_device_loop_P1_ms"] = round(timed(lambda: be.ce_solve_fleet(spec, x0[:, None], u, 0.1, 1.0, 2.0, seed=3), 3), 3)
    print(json.dumps(out), flush=True)


def mpc_cpu():
    s = R.CrossEntropyBilevelOptimizationSolver(backend=o)
    return CE.solve_(s, prob, x0, ua, np.random.default_rng(1), kl_bound=0.1)


th = wl.c2_thetas(1024)
print(json.dumps({"cpu_oracle_threads": int(o.raw.oracle_get_threads()), "mpc_step_ms": round(timed(mpc_cpu, 3), 3),
                  "ce_costs_1x1024_ms": round(timed(lambda: o.ce_costs(spec, x0, u, th, 0.1), 2), 3)}))
be.close()
