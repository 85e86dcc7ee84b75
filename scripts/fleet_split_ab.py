"""A/B of the concurrent sub-fleet blocks of ratilqr_ce_solve_fleet (RATILQR_FLEET_SPLIT=k): C5 fleet MPC step latency and
identity of the results with the one-block run.   python scripts/fleet_split_ab.py [P]"""
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
    be.ce_solve_fleet(spec, x0, u, 0.1, 1.0, 2.0, seed=7, want=())
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        r = be.ce_solve_fleet(spec, x0, u, 0.1, 1.0, 2.0, seed=7, want=("l",))
        ts.append(time.perf_counter() - t0)
    np.savez(sys.argv[3], theta=r["theta_opt"], value=r["value"], l=r["l"], status=r["status"])
    print(json.dumps({"split": os.environ.get("RATILQR_FLEET_SPLIT", "default"), "order": os.environ.get("RATILQR_FLEET_ORDER", "desc"),
                      "problems": P, "ms_per_fleet_step": min(ts) * 1e3,
                      "ms_all": [t * 1e3 for t in ts], "rounds": r["rounds"], "ok": int((r["status"] == 0).sum())}))
    sys.exit(0)

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ref = None
for k, order in (("1", "desc"), ("2", "desc"), ("3", "desc"), ("4", "desc"), ("5", "desc"), ("6", "desc")):
    out = f"/tmp/fleet_split_{k}_{order}.npz"
    env = dict(os.environ, RATILQR_FLEET_SPLIT=k, RATILQR_FLEET_ORDER=order)
    line = subprocess.run([sys.executable, __file__, "--child", str(P), out], env=env, capture_output=True, text=True)
    if line.returncode != 0:
        print(json.dumps({"split": k, "order": order, "error": line.stderr[-400:]}))
        continue
    d = json.loads(line.stdout.strip().splitlines()[-1])
    z = np.load(out)
    if ref is None:
        ref = z
    d["identical_to_one_block"] = bool(all(np.array_equal(z[f], ref[f]) for f in ("theta", "value", "l", "status")))
    print(json.dumps(d), flush=True)
