"""N-GPU check of the multi-GPU C ABI in ONE process (ratilqr_create_multi: ncclCommInitAll + one host thread per device):
   python scripts/check_multi_gpu.py [n_dev]
 * sharded theta population of one problem (configs[1], 1024 theta; and 65,536 theta): cost vector identical to one GPU;
 * ratilqr_multi_ce_solve (sharded CE, redundant on-device elite selection): identical to ratilqr_ce_solve on one GPU;
 * PETS cost vector sharded over action sequences (Philox streams by global index): identical to one GPU;
 * fleet block-partitioned over the devices (no collective): identical per problem.
One JSON line per check (timings are wall clock around the blocking C-ABI call, host buffers in and out)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ratilqr_b200 as R  # noqa: E402
from ratilqr_b200 import workloads as wl  # noqa: E402

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mg = R.new_multi(range(n_dev))
be = R.new_backend(0)


def timed(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); t.append(time.perf_counter() - t0)
    return min(t) * 1e3, r


prob, x0, u = wl.c2_problem()
spec = prob.spec()
for K in (1024, 65536):
    th = wl.c2_thetas(K)
    t1, (c1, s1) = timed(lambda: be.ce_costs(spec, x0, u, th, 0.1))
    tn, (cn, sn) = timed(lambda: mg.ce_costs(spec, x0, u, th, 0.1))
    print(json.dumps({"check": "sharded theta population (ratilqr_multi_ce_costs)", "n_dev": n_dev, "thetas": K,
                      "identical_to_one_gpu": bool(np.array_equal(c1, cn) and np.array_equal(s1, sn)), "ms_one_gpu": round(t1, 3),
                      "ms_sharded": round(tn, 3), "solves_per_s_one_gpu": round(K / t1 * 1e3), "solves_per_s_sharded": round(K / tn * 1e3)}), flush=True)

kw = dict(num_samples=1024, num_elite=100, iter_max=3, seed=11)
t1, r1 = timed(lambda: be.ce_solve(spec, x0, u, 0.1, 1.0, 2.0, **kw), 2)
tn, rn = timed(lambda: mg.ce_solve(spec, x0, u, 0.1, 1.0, 2.0, **kw), 2)
same = all(r1[k] == rn[k] for k in ("theta_opt", "value", "mu", "sigma", "theta_min", "theta_max", "nz_used")) and np.array_equal(r1["L"], rn["L"])
print(json.dumps({"check": "sharded RAT iLQR solve (ratilqr_multi_ce_solve), 1024 theta x 3 CE iterations + final", "n_dev": n_dev,
                  "identical_to_one_gpu": bool(same), "ms_one_gpu": round(t1, 3), "ms_sharded": round(tn, 3), "theta_opt": rn["theta_opt"]}), flush=True)

pprob, px0 = wl.c4_problem()
pspec, gen = pprob.spec(), pprob.f_stochastic.gen()
ctrl = 2.0 * np.random.Generator(np.random.Philox(key=7)).standard_normal((1, 30, 4096))
t1, p1 = timed(lambda: be.pets_costs(pspec, px0, ctrl, 150, seed=5, gen=gen))
tn, pn = timed(lambda: mg.pets_costs(pspec, px0, ctrl, 150, seed=5, gen=gen))
print(json.dumps({"check": "sharded PETS costs 4096 x 150 x T30 (ratilqr_multi_pets_costs)", "n_dev": n_dev,
                  "identical_to_one_gpu": bool(np.array_equal(p1, pn)), "ms_one_gpu": round(t1, 3), "ms_sharded": round(tn, 3)}), flush=True)

P = 2048 * n_dev
fprob, cps, fx0, fu = wl.fleet(P)
fspec = fprob.spec(cost_params=cps)
t1, f1 = timed(lambda: be.ce_solve_fleet(fspec, fx0, fu, 0.1, 1.0, 2.0, seed=3, want=("l",)), 1)
tn, fn_ = timed(lambda: mg.ce_solve_fleet(fspec, fx0, fu, 0.1, 1.0, 2.0, seed=3), 1)
print(json.dumps({"check": f"fleet of {P} RAT iLQR problems block-partitioned (ratilqr_multi_ce_solve_fleet), no collective", "n_dev": n_dev,
                  "identical_to_one_gpu": bool(np.array_equal(f1["theta_opt"], fn_["theta_opt"]) and np.array_equal(f1["l"], fn_["l"])),
                  "ms_one_gpu": round(t1, 3), "ms_multi": round(tn, 3), "problems_per_s_multi": round(P / tn * 1e3)}), flush=True)
mg.close()
be.close()
