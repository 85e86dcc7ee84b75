"""A/B of the slot ordering in ratilqr_nm_solve_fleet (RATILQR_FLEET_SORT=0/1): RAT iLQR++ on a unicycle fleet.
    python scripts/nm_fleet_ab.py [P]"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 2 and sys.argv[1] == "--child":
    import ratilqr_b200 as R
    from ratilqr_b200 import workloads as wl
    P = int(sys.argv[2])
    be = R.new_backend(0)
    prob, cps, x0, u = wl.fleet(P, key=70)
    spec = prob.spec(cost_params=cps)
    be.nm_solve_fleet(spec, x0, u, 0.1, iter_max=20, want=())
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        r = be.nm_solve_fleet(spec, x0, u, 0.1, iter_max=20, want=("l",))
        ts.append(time.perf_counter() - t0)
    np.savez(sys.argv[3], theta=r["theta_opt"], value=r["value"], l=r["l"])
    print(json.dumps({"sort": os.environ.get("RATILQR_FLEET_SORT", "1"), "problems": P, "ms_per_fleet_solve": min(ts) * 1e3,
                      "nm_iters_mean": float(np.mean(r["nm_iters"])), "evals_mean": float(np.mean(r["n_evals"]))}))
    sys.exit(0)

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ref = None
for k in ("0", "1"):
    out = f"/tmp/nm_fleet_{k}.npz"
    line = subprocess.run([sys.executable, __file__, "--child", str(P), out], env=dict(os.environ, RATILQR_FLEET_SORT=k),
                          capture_output=True, text=True)
    if line.returncode != 0:
        print(json.dumps({"sort": k, "error": line.stderr[-600:]}))
        continue
    d = json.loads(line.stdout.strip().splitlines()[-1])
    z = np.load(out)
    if ref is None:
        ref = z
    d["identical"] = bool(all(np.array_equal(z[f], ref[f]) for f in ("theta", "value", "l")))
    print(json.dumps(d), flush=True)
