"""Tuning sweep (GPU box): kernel-only time of the unicycle solve for each launch shape.
usage: python scripts/tune_solve.py <shape> <problems> [reps]   -> one JSON line"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
shape, P = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
os.environ["RATILQR_SOLVE_SHAPE"] = shape
import numpy as np  # noqa: E402

import bench  # noqa: E402
import ratilqr_b200 as R  # noqa: E402

be = R.new_backend(0)
spec, x0, u, theta = bench.build_inputs(P, 0)
if os.environ.get("TUNE_SORT_THETA"):
    theta = np.sort(theta.reshape(P, -1), axis=1).reshape(-1)
be.stage(spec, x0, u, theta, P=P)
be.run(2)
sampler = bench.ClockSampler(0)
sampler.start()
ms = be.run(reps) / reps
clk = sampler.stop()
res = be.fetch()
flops = float(np.sum(bench.algorithmic_flops(res["iters"], res["trials"])))
print(json.dumps({"shape": shape, "problems": P, "ms": ms, "solves_per_s": theta.size / ms * 1e3,
                  "tflops": flops / ms / 1e9, "ok": int((res["status"] == 0).sum()), "B": int(theta.size), "sm_mhz": clk.get("sm_mhz"), "reasons": clk.get("reasons"),
                  "lib": os.path.basename(os.environ.get("RATILQR_B200_LIB", "default")),
                  "sorted": bool(os.environ.get("TUNE_SORT_THETA")), "iters_minmax": [int(res["iters"].min()), int(res["iters"].max())],
                  "trials_minmax": [int(res["trials"].min()), int(res["trials"].max())]}))
