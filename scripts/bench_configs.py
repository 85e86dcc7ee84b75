"""Secondary measurements for BASELINE.json configs C2 (single MPC step), C3, C4, C5 and the MC-rollout kernel.
One JSON line per config -> profiles/.  GPU timings are wall-clock around blocking C-ABI calls (host buffers
in/out, i.e. end-to-end); the CPU column is the oracle on all host threads on a bounded sample."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import ratilqr_b200 as R  # noqa: E402
from ratilqr_b200 import cross_entropy as CE  # noqa: E402
from ratilqr_b200 import nelder_mead as NM  # noqa: E402
from ratilqr_b200 import workloads as wl  # noqa: E402

be = R.new_backend(0)
o = oracle.load()
cores = int(o.raw.oracle_get_threads())
which = sys.argv[1:] or ["c2_mpc", "c3", "c4", "c5", "mc"]


def timed(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return min(t)


if "c2_mpc" in which:  # one RAT iLQR MPC step (CE defaults: 10 theta x 5 iterations + final solve), single problem
    prob, x0, u = wl.c2_problem()
    ua = [u[:, k].copy() for k in range(u.shape[1])]

    def gpu_step():
        s = R.CrossEntropyBilevelOptimizationSolver(backend=be)
        return CE.solve_(s, prob, x0, ua, np.random.default_rng(1), kl_bound=0.1)

    def cpu_step():
        s = R.CrossEntropyBilevelOptimizationSolver(backend=o)
        return CE.solve_(s, prob, x0, ua, np.random.default_rng(1), kl_bound=0.1)

    tg, tc = timed(gpu_step), timed(cpu_step, 1)
    a, b = gpu_step(), cpu_step()
    print(json.dumps({"config": "C2 single-problem RAT iLQR MPC step (10 theta x 5 CE its + final)", "gpu_ms": tg * 1e3,
                      "cpu_oracle_ms": tc * 1e3, "cpu_cores": cores, "theta_opt_gpu": a[0], "theta_opt_cpu": b[0],
                      "note": "latency-bound: 6 sequential launches of <= 10 one-thread solves"}))

if "c3" in which:  # RAT iLQR++ (Nelder-Mead) on the 12-state quadrotor, T = 40
    prob, x0, u = wl.c3_problem()
    ua = [u[:, k].copy() for k in range(u.shape[1])]

    def run(backend):
        nm = R.NelderMeadBilevelOptimizationSolver(backend=backend)
        out = NM.solve_(nm, prob, x0, ua, kl_bound=0.1)
        return out[0], out[4], nm.iter_current, nm.n_evals

    tg = timed(lambda: run(be), 1)
    tc = timed(lambda: run(o), 1)
    g, c = run(be), run(o)
    sol = o.ileqg_solve_batch(prob.spec(), x0, u, [g[0]])
    mc = be.mc_rollout(prob.spec(), sol["x"][..., 0], sol["l"][..., 0], sol["L"][..., 0], 256, seed=3, theta_risk=g[0])
    print(json.dumps({"config": "C3 RAT iLQR++ quadrotor n=12 m=4 T=40", "gpu_ms": tg * 1e3, "cpu_oracle_ms": tc * 1e3,
                      "cpu_cores": cores, "theta_opt": [g[0], c[0]], "value": [g[1], c[1]], "nm_iters": g[2], "evals": g[3],
                      "mc256_mean_var_risk": mc["stats"][0].tolist(),
                      "note": "thread-per-instance kernel runs n=12 from local memory; CTA-per-instance kernel is next"}))

if "c3_fleet" in which:  # batched quadrotor iLEQG solves (warp-cooperative kernel) vs the CPU oracle
    prob, x0, u = wl.c3_problem()
    spec = prob.spec()
    for B in (6, 740, 4096):
        th = np.concatenate([[0.0], wl.positive_thetas(B - 1, mu=0.02, sigma=0.02, key=B)])
        tg = timed(lambda: be.ce_costs(spec, x0, u, th, 0.1), 2)
        cost, st = be.ce_costs(spec, x0, u, th, 0.1)
        Bc = min(B, 64)
        tc = timed(lambda: o.ce_costs(spec, x0, u, th[:Bc], 0.1), 1)
        print(json.dumps({"config": f"C3 quadrotor n=12 m=4 T=40, {B} iLEQG solves in one call (coop kernel)", "gpu_ms": tg * 1e3,
                          "gpu_solves_per_s": B / tg, "feasible": int((st == 0).sum()), "cpu_oracle_solves_per_s": Bc / tc,
                          "cpu_cores": cores, "cpu_sample": Bc}))

if "c4" in which:  # PETS CEM on cart-pole: 4096 sequences x (5-model ensemble x 30 particles), T = 30, 5 iterations
    prob, x0 = wl.c4_problem()
    spec, gen = prob.spec(), prob.f_stochastic.gen()
    mu0, Sg0 = np.zeros((1, 30)), np.tile(np.array([[4.0]])[:, :, None], (1, 1, 30))
    tg = timed(lambda: be.pets_solve(spec, x0, mu0, Sg0, 4096, 150, 409, 5, 0.1, seed=1, gen=gen))
    rollouts = 4096 * 150 * 5
    # CPU: one iteration's cost evaluation on a bounded sample (256 sequences) with injected noise
    rng = np.random.default_rng(0)
    ctrl = 2.0 * rng.standard_normal((1, 30, 256))
    noise = 1e-2 * rng.standard_normal((4, 30, 150, 256))
    tc = timed(lambda: o.pets_costs(spec, x0, ctrl, 150, noise=noise, gen=gen), 1)
    print(json.dumps({"config": "C4 PETS cart-pole 4096 x 150 x T30, 5 CEM iterations", "gpu_ms_per_solve": tg * 1e3,
                      "gpu_rollouts_per_s": rollouts / tg, "cpu_oracle_rollouts_per_s": 256 * 150 / tc, "cpu_cores": cores,
                      "cpu_sample": "256 sequences x 150 particles, one cost evaluation"}))

if "c5" in which:  # fleet of independent RAT iLQR unicycle problems, CE defaults, on-device loop (8192 = 65,536 / 8 GPUs)
    P = int(os.environ.get("C5_PROBLEMS", "8192"))
    prob, cps, x0, u = wl.fleet(P)
    spec = prob.spec(cost_params=cps)
    res = {}

    def run():
        res["r"] = be.ce_solve_fleet(spec, x0, u, 0.1, 1.0, 2.0, seed=7, want=("x", "l", "L"))

    tg = timed(run, 2)
    r = res["r"]
    solves = P * 10 * r["rounds"] + P
    tm = timed(lambda: be.mc_rollout(spec, r["x"], r["l"], r["L"], 256, seed=9, theta_risk=1.0, P=P), 2)
    Pc = 16
    probc, cpsc, x0c, uc = wl.fleet(Pc)
    z = np.random.default_rng(0).standard_normal((Pc, 4000))
    import ctypes as C
    from tests.test_reference_bilevel import OracleCEOpts, dp
    f = o.raw.oracle_ce_solve
    f.restype = C.c_int32

    def cpu():
        for p in range(Pc):
            sp = probc.spec(cost_params=cpsc[p]); d = sp.desc(); opts = R.make_opts(); ce = OracleCEOpts(1.0, 2.0, 10, 3, 5, 0.5, 0)
            outs = [C.c_double() for _ in range(6)]; nz, st = C.c_int64(), C.c_int32()
            x0p = np.ascontiguousarray(x0c[:, p]); uf = np.ascontiguousarray(uc.ravel(order="F"))
            f(C.byref(d), C.byref(opts), C.byref(ce), x0p.ctypes.data_as(dp), uf.ctypes.data_as(dp), C.c_double(0.1),
              z[p].ctypes.data_as(dp), C.c_int64(4000), *[C.byref(q) for q in outs], C.byref(nz), None, None, None, C.byref(st))

    tc = timed(cpu, 1)
    print(json.dumps({"config": f"C5 fleet of {P} RAT iLQR unicycle problems (10 theta x 5 CE its + final) on one GPU",
                      "gpu_ms_per_fleet_step": tg * 1e3, "gpu_problems_per_s": P / tg, "rounds": r["rounds"],
                      "gpu_ileqg_solves_per_s": solves / tg, "ok_final": int((r["status"] == 0).sum()),
                      "mc_256_per_problem_ms": tm * 1e3, "mc_rollouts_per_s": P * 256 / tm,
                      "cpu_oracle_problems_per_s": Pc / tc, "cpu_cores": cores, "cpu_sample": f"{Pc} problems (the oracle's CE loop is serial per problem; its theta fan-out uses all threads)"}))

if "mc" in which:  # injected-noise MC rollouts: HBM-bound (8 n N bytes of noise per sample)
    prob, x0, u = wl.c2_problem()
    spec = prob.spec()
    sol = o.ileqg_solve_batch(spec, x0, u, [1.0])
    S = 1 << 20
    w = 1e-2 * np.random.default_rng(0).standard_normal((4, 50, S))
    t0 = time.perf_counter()
    be.mc_rollout(spec, sol["x"][..., 0], sol["l"][..., 0], sol["L"][..., 0], S, noise=w)
    te2e = time.perf_counter() - t0
    tp = timed(lambda: be.mc_rollout(spec, sol["x"][..., 0], sol["l"][..., 0], sol["L"][..., 0], S, seed=1), 2)
    print(json.dumps({"config": "MC closed-loop rollouts, unicycle T=50, 2^20 samples", "injected_noise_e2e_ms": te2e * 1e3,
                      "injected_noise_bytes": int(w.nbytes), "philox_ms": tp * 1e3, "philox_rollouts_per_s": S / tp}))
be.close()
