"""Generates tests/golden/literal_*.npz from tests/ref_literal/ileqg_literal.py (the operation-by-operation
restatement of src/ileqg.jl in numpy-float64 and in 60-digit mpmath; it shares no code with the oracle or the kernels).

    python tests/golden/make_literal_goldens.py [riccati] [c1] [c2] [c3]      (default: all; c2/c3 take minutes in mpmath)

Fixtures (inputs as float64; outputs = the 60-digit results rounded to float64, plus the float64-literal results):
  literal_riccati.npz   dense random SPD stage data at all five (n, m); theta in {0, 0.3 theta*, 0.97 theta*} (theta* =
                        the neurotic-breakdown threshold found by bisection in mpmath); both passes
                        (solve_approximate_dp!, solve_approximate_dp with given L, dl and mu > 0), plus a case whose
                        R is indefinite so that the mu-restart loop (:372-378) runs
  literal_c1.npz        whole solve! of the reference's shipped test problem (C1) at theta in {0, .1, .3, .43, .5, 30.7}
                        and the infeasible theta = 31 (M not PD in initialize!)
  literal_c2.npz        whole solve! of configs[1]'s unicycle (N = 50) for the first 10 theta of the 1024 population
  literal_c3.npz        whole solve! of configs[2]'s quadrotor (N = 40) at theta in {0, 0.01, 0.02}
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from tests.ref_literal import ileqg_literal as lit  # noqa: E402


# ---------------------------------------------------------------------------------------------------------------
# stage-level data
# ---------------------------------------------------------------------------------------------------------------
def random_stage_data(n, m, N, rng, indefinite_R=False):
    def spd(k, scale):
        G = rng.standard_normal((k, k))
        return scale * (G @ G.T / k + 0.5 * np.eye(k))

    d = {}
    d["q"] = rng.standard_normal(N + 1)
    d["qv"] = rng.standard_normal((n, N + 1))
    d["Q"] = np.stack([spd(n, 1.0) for _ in range(N + 1)], axis=-1)
    d["r"] = rng.standard_normal((m, N))
    d["R"] = np.stack([spd(m, 1.0) - (3.0 * n * np.eye(m) if indefinite_R else 0.0) for _ in range(N)], axis=-1)
    d["P"] = 0.3 * rng.standard_normal((m, n, N))
    d["A"] = np.stack([np.eye(n) + 0.3 * rng.standard_normal((n, n)) for _ in range(N)], axis=-1)
    d["B"] = rng.standard_normal((n, m, N))
    d["W"] = spd(n, 0.05)
    return d


def approx_from(LA, d):
    N = d["r"].shape[1]
    ap = lit.Approx()
    ap.q = [LA.num(v) for v in d["q"]]
    ap.q_vec = [LA.arr(d["qv"][:, k]) for k in range(N + 1)]
    ap.Q = [LA.arr(d["Q"][:, :, k]) for k in range(N + 1)]
    ap.r = [LA.arr(d["r"][:, k]) for k in range(N)]
    ap.R = [LA.arr(d["R"][:, :, k]) for k in range(N)]
    ap.P = [LA.arr(d["P"][:, :, k]) for k in range(N)]
    ap.A = [LA.arr(d["A"][:, :, k]) for k in range(N)]
    ap.B = [LA.arr(d["B"][:, :, k]) for k in range(N)]
    W = LA.arr(d["W"])
    ap.W = [W for _ in range(N)]
    return ap


def run_pass(LA, d, theta, optimise, L=None, dl=None, mu=0.0):
    """returns dict(s, sv, S, L, dl, mu, delta, restarts) as float64 arrays, or None when M is not PD"""
    ap = approx_from(LA, d)
    n, m, N = d["qv"].shape[0], d["r"].shape[0], d["r"].shape[1]
    sol = lit.Solver(LA)
    sol.mu, sol.delta = LA.num(mu), sol.delta_0
    sol.L_array = [None] * N
    try:
        if optimise:
            res, dln = lit.solve_approximate_dp_opt(LA, sol, ap, LA.num(theta))
            Ls = sol.L_array
        else:
            Ls = [LA.arr(L[:, :, k]) for k in range(N)]
            dln = None if dl is None else [LA.arr(dl[:, k]) for k in range(N)]
            res = lit.solve_approximate_dp(LA, ap, Ls, dln, LA.num(theta), LA.num(mu))
    except lit.NotPosDef:
        return None
    out = dict(s=LA.to_f64(res.s), sv=np.stack([LA.to_f64(v) for v in res.s_vec], axis=-1),
               S=np.stack([LA.to_f64(v) for v in res.S], axis=-1),
               L=np.stack([LA.to_f64(v) for v in Ls], axis=-1), mu=float(sol.mu), delta=float(sol.delta),
               restarts=sol.restarts)
    if dln is not None:
        out["dl"] = np.stack([LA.to_f64(v) for v in dln], axis=-1)
    return out


def breakdown_theta(d, Lg, dlg):
    """largest theta for which all three passes stored in the fixture stay feasible (bisection in mpmath, 40 steps)"""
    def ok(theta):
        return (run_pass(lit.MP, d, theta, True) is not None and run_pass(lit.MP, d, theta, False, Lg, dlg, mu=0.5) is not None
                and run_pass(lit.MP, d, theta, False, Lg, None, mu=0.0) is not None)
    lo, hi = 0.0, 1.0
    while ok(hi):
        lo, hi = hi, hi * 2
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        if ok(mid):
            lo = mid
        else:
            hi = mid
    return lo


def make_riccati():
    rng = np.random.Generator(np.random.Philox(key=20241017))
    out = {}
    cases = []
    for (n, m) in [(2, 2), (2, 1), (4, 2), (4, 1), (12, 4)]:
        N = 3
        d = random_stage_data(n, m, N, rng)
        Lg, dlg = 0.2 * rng.standard_normal((m, n, N)), 0.2 * rng.standard_normal((m, N))
        tstar = breakdown_theta(d, Lg, dlg)
        for lvl, theta in (("zero", 0.0), ("mid", 0.3 * tstar), ("near", 0.97 * tstar)):
            cases.append((f"n{n}m{m}_{lvl}", d, theta, Lg, dlg))
        dR = random_stage_data(n, m, N, rng, indefinite_R=True)
        Lg, dlg = 0.2 * rng.standard_normal((m, n, N)), 0.2 * rng.standard_normal((m, N))
        cases.append((f"n{n}m{m}_restart", dR, 0.2 * breakdown_theta(dR, Lg, dlg), Lg, dlg))
    for name, d, theta, Lg, dlg in cases:
        n, m, N = d["qv"].shape[0], d["r"].shape[0], d["r"].shape[1]
        for k, v in d.items():
            out[f"{name}/in/{k}"] = v
        out[f"{name}/in/theta"] = np.float64(theta)
        out[f"{name}/in/L_eval"] = Lg
        out[f"{name}/in/dl_eval"] = dlg
        for LA in (lit.MP, lit.F64):
            ro = run_pass(LA, d, theta, True)
            re = run_pass(LA, d, theta, False, Lg, dlg, mu=0.5)
            rn = run_pass(LA, d, theta, False, Lg, None, mu=0.0)  # dl = nothing (the line-search form)
            for tag, r in (("opt", ro), ("eval", re), ("eval_nodl", rn)):
                assert r is not None, (name, tag)
                for k, v in r.items():
                    out[f"{name}/{LA.name}/{tag}/{k}"] = np.asarray(v)
        print(name, "theta", theta, "restarts", out[f"{name}/mp/opt/restarts"], "s0", out[f"{name}/mp/opt/s"][0], flush=True)
    np.savez_compressed(os.path.join(HERE, "literal_riccati.npz"), **out)


# ---------------------------------------------------------------------------------------------------------------
# whole solves
# ---------------------------------------------------------------------------------------------------------------
def _solve_case(args):
    la_name, model, mparams, costkind, cp, W, N, n, m, x0, u, theta = args
    LA = lit.MP if la_name == "mp" else lit.F64
    f = lit.make_dynamics(LA, model, mparams)
    c, h = lit.make_power_law_cost(LA, cp) if costkind == "pow" else lit.make_quadratic_cost(LA, n, m, cp)
    Wm = LA.arr(W)
    prob = lit.Problem(f, c, h, lambda k: Wm, N, n, m)
    sol = lit.Solver(LA)
    t0 = time.time()
    try:
        x, l, L, v, eh = lit.solve(LA, sol, prob, x0, [u[:, k] for k in range(N)], theta)
    except lit.NotPosDef:
        status = 1 if sol.iter_current == 0 else 2
        return dict(status=status, theta=theta, iters=sol.iter_current, seconds=time.time() - t0)
    except lit.DomainError:
        return dict(status=3, theta=theta, iters=sol.iter_current, seconds=time.time() - t0)
    return dict(status=0, theta=theta, x=np.stack([LA.to_f64(a) for a in x], axis=-1),
                l=np.stack([LA.to_f64(a) for a in l], axis=-1), L=np.stack([LA.to_f64(a) for a in L], axis=-1),
                value=float(v), iters=sol.iter_current, trials=len(eh), restarts=sol.restarts, mu=float(sol.mu),
                d_current=float(sol.d_current), eps_hist=np.array([[float(a), float(b)] for a, b in eh]).reshape(-1, 2),
                seconds=time.time() - t0)


def make_solves(tag, model, mparams, costkind, cp, W, N, n, m, x0, u, thetas, pool):
    jobs = [(la, model, list(map(float, mparams)), costkind, list(map(float, cp)), np.asarray(W, float), N, n, m,
             np.asarray(x0, float), np.asarray(u, float), float(t)) for la in ("mp", "f64") for t in thetas]
    res = pool.map(_solve_case, jobs, chunksize=1)
    out = dict(thetas=np.asarray(thetas, float), x0=np.asarray(x0, float), u_init=np.asarray(u, float))
    for (la, *_rest), r in zip(jobs, res):
        i = list(thetas).index(r["theta"])
        for k, v in r.items():
            out[f"{la}/{i}/{k}"] = np.asarray(v)
        print(tag, la, "theta", r["theta"], "status", r["status"], "iters", r.get("iters"), "trials", r.get("trials"),
              "value", r.get("value"), f"{r['seconds']:.1f}s", flush=True)
    np.savez_compressed(os.path.join(HERE, f"literal_{tag}.npz"), **out)


def main():
    import ratilqr_b200  # noqa: F401  (only for the frozen workload definitions: inputs, not results)
    from ratilqr_b200 import workloads as wl
    which = set(sys.argv[1:]) or {"riccati", "c1", "c2", "c3"}
    if "riccati" in which:
        make_riccati()
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        if "c1" in which:
            prob, x0, u = wl.c1_problem()
            sp = prob.spec()
            make_solves("c1", "power_law", sp.model_params, "pow", sp.cost_params, np.asarray(prob.W(0)), 10, 2, 2, x0, u,
                        [0.0, 0.1, 0.3, 0.43, 0.5, 30.7, 31.0], pool)
        if "c2" in which:
            prob, x0, u = wl.c2_problem()
            sp = prob.spec()
            make_solves("c2", "unicycle", sp.model_params, "quad", sp.cost_params, np.asarray(prob.W(0)), 50, 4, 2, x0, u,
                        list(wl.c2_thetas(10)), pool)
        if "c3" in which:
            prob, x0, u = wl.c3_problem()
            sp = prob.spec()
            make_solves("c3", "quadrotor", sp.model_params, "quad", sp.cost_params, np.asarray(prob.W(0)), 40, 12, 4, x0, u,
                        [0.0, 0.01, 0.02], pool)


if __name__ == "__main__":
    main()
