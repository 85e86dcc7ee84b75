"""Feasibility experiment for a CE round of a RAT iLQR fleet (8192 problems x 10 theta): the ~5 % of the problems that run
to iter_max form the tail of the round.  A: everything in one launch of the thread-per-instance kernel (heaviest first).
B: the heavy problems on the speculative latency kernel (own context / stream / host thread), concurrently with the rest on
the throughput kernel.  Results are identical; prints wall times."""
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ratilqr_b200 as R  # noqa: E402
from ratilqr_b200 import workloads as wl  # noqa: E402

P, S = int(os.environ.get("P", "8192")), 10
prob, cps, x0, u = wl.fleet(P, key=70)
theta = wl.positive_thetas(P * S, key=3)
a, b = R.new_backend(0), R.new_backend(0)
spec = prob.spec(cost_params=cps)
r = a.ileqg_solve_batch(spec, x0, u, theta, want=(), P=P)
key = r["iters"].reshape(P, S).max(axis=1)
med = np.median(key)
out = {"P": P, "median_iters": float(med), "iters_hist": np.bincount(np.minimum(key // 10, 10)).tolist()}


def timed(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return min(t) * 1e3


out["A_one_launch_ms"] = timed(lambda: a.ce_costs(spec, x0, u, theta, 0.1, P=P))
order = np.argsort(-key, kind="stable")
for thr in (200, 300, 423, 473):  # the `thr` heaviest problems (423 of 8192 run to iter_max = 100)
    heavy, light = np.sort(order[:thr]), np.sort(order[thr:])
    if heavy.size * S > 4736 or heavy.size == 0:
        continue
    sh = prob.spec(cost_params=cps[heavy]); sl = prob.spec(cost_params=cps[light])
    th, tl = theta.reshape(P, S)[heavy].ravel(), theta.reshape(P, S)[light].ravel()
    xh, xl = np.ascontiguousarray(x0[:, heavy]), np.ascontiguousarray(x0[:, light])
    res = {}

    def both():
        t = threading.Thread(target=lambda: res.__setitem__("h", b.ce_costs(sh, xh, u, th, 0.1, P=heavy.size)))
        t.start()  # heavy first: its CTAs are resident before the bulk fills the SMs
        res["l"] = a.ce_costs(sl, xl, u, tl, 0.1, P=light.size)
        t.join()

    out[f"B_split_thr{thr}_ms"] = timed(both)
    out[f"B_split_thr{thr}_heavy_problems"] = int(heavy.size)
    out[f"heavy_alone_spec_ms_thr{thr}"] = timed(lambda: b.ce_costs(sh, xh, u, th, 0.1, P=heavy.size))
    out[f"light_alone_ms_thr{thr}"] = timed(lambda: a.ce_costs(sl, xl, u, tl, 0.1, P=light.size))
print(json.dumps(out))
