#!/usr/bin/env python
"""bench.py -- batched iLEQG solves/sec (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                   (CPU oracle on the host cores, same config)

Workload (config.workload = "c2_fleet"): BASELINE.json configs[1] -- 4-state unicycle, T=50, 1024 theta
samples per problem -- replicated over P independent problems per GPU (x0 and goal drawn per problem, as
in configs[4]) so that one step fills the device: 1776 x 1024 = 1,818,624 iLEQG solves = 32 full waves of
148 SMs x 384 resident instances (12 warps/SM at 168 registers, 128-thread CTAs; 16 GB of SoA workspace).  A "step" =
one batched solve of all instances (one launch of the persistent solve kernel).  (--problems 444 = 8 waves: -8 %.)
Weak scaling: every rank owns P problems; there is no data-path collective (instances are independent).

value  : solves/s, kernel only, inputs resident in HBM, CUDA events on the library's stream.
e2e    : solves/s through the reference-facing call ratilqr_ce_costs (= compute_cost,
         cross_entropy_bilevel_optimization.jl:173-195) with HOST buffers: H2D of x0/u/theta/cost parameters and
         D2H of the cost + status vectors inside the timed region, every step.
extra  : "latency" = the small-batch cases of BASELINE.json with the CPU oracle timed beside each: configs[1] exactly
         (1 problem x 1024 theta), a single problem's RAT iLQR MPC step (10 theta x 5 CE iterations + final solve, whole
         loop on the device) and configs[2]'s RAT iLQR++ on the quadrotor; "roofline_c3" / "roofline_c4" = the
         warp-cooperative (n = 12) solve kernel and the PETS rollout kernel against the FP64 roof.
The CPU baseline / reference arm is the in-repo C++ ORACLE (a restatement of the Julia reference: there is no Julia in
the image), all host threads, on a sample of problems drawn evenly from the SAME fleet; the sample is named in the line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STATE, N_CTRL, HORIZON, THETAS = 4, 2, 50, 1024


def f_opt(n, m):  # algorithmic flops of one optimising Riccati stage
    return 25.0 / 3.0 * n ** 3 + 8 * n * n * m + 6 * n * m * m + m ** 3 / 3.0 + 14 * n * n + 6 * n * m + 6 * m * m


def f_eval(n, m):
    return f_opt(n, m) - (m ** 3 / 3.0 + 2 * m * m * n + 2 * m * m)


F_LIN_UNICYCLE = 150.0  # quadratic cost derivatives + unicycle Jacobian per stage (DESIGN.md "Algorithmic work")
F_F_UNICYCLE = 12.0     # one unicycle step (sin, cos counted as 1 flop each)


def algorithmic_flops(iters, trials, n=N_STATE, m=N_CTRL, N=HORIZON):
    """per-instance algorithmic flops from the device counters (SURVEY.md 8d formula)"""
    iters, trials = iters.astype(np.float64), trials.astype(np.float64)
    return (N * (iters * f_opt(n, m) + (1 + trials) * f_eval(n, m)) + (1 + iters + trials) * N * F_LIN_UNICYCLE
            + (1 + trials) * N * F_F_UNICYCLE)


def algorithmic_bytes(iters, trials, n=N_STATE, m=N_CTRL, N=HORIZON):
    """unavoidable HBM bytes per instance: trajectory + policy traffic of every pass"""
    passes = (1 + iters + 2 * trials).astype(np.float64)  # backward passes + candidate rollouts
    return 8.0 * (n + 2 * m + m * n) * N * passes + 8.0 * (2 + 4)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


DISTINCT_FLEETS = False  # --distinct-fleets: every rank draws its own problems (round-1 behaviour: +-5 % work per rank)


def build_inputs(P, rank):
    """Weak scaling with EQUAL per-GPU work: every rank owns the same P problem definitions (x0, goal) and its own theta
    populations.  The iterations a solve needs are a property of the problem far more than of theta, so ranks that draw
    their own problems differ by +-5 % in work and the job is timed by the heaviest draw (round 1: 410 vs 384 ms)."""
    from ratilqr_b200 import workloads as wl
    prob, cps, x0, u = wl.fleet(P, key=7 + (1000 * rank if DISTINCT_FLEETS else 0))
    theta = np.concatenate([wl.positive_thetas(THETAS, key=20201028 + p + 100000 * rank) for p in range(P)])
    return prob.spec(cost_params=cps), x0, u, theta


_FLEET_CACHE = {}


def fleet_sample(P, count, shift=0):
    """`count` problems taken evenly from rank 0's fleet of P problems (same generator, same theta populations): a sample
    of the SAME workload, heavy (iter_max) and light problems in the fleet's own proportions.  `shift` rotates the
    selection so that successive steps of the CPU arm measure DIFFERENT problems."""
    from ratilqr_b200 import workloads as wl
    if P not in _FLEET_CACHE:
        _FLEET_CACHE[P] = wl.fleet(P, key=7)
    prob, cps, x0, u = _FLEET_CACHE[P]
    stride = max(P // max(count, 1), 1)
    idx = np.unique((np.arange(count) * stride + shift * max(stride // 7, 1) + shift) % P)
    theta = np.concatenate([wl.positive_thetas(THETAS, key=20201028 + int(p)) for p in idx])
    return prob.spec(cost_params=cps[idx]), np.ascontiguousarray(x0[:, idx]), u, theta, idx


def cpu_arm(P, steps, warmup, per_step):
    """The oracle (CPU restatement of the reference) on all host threads.  Every step solves `per_step` problems x 1024
    theta taken evenly from the same fleet, a different selection each step, so K steps sample K * per_step distinct
    problems of the workload (about 5 % of the problems run to iter_max and cost 5x the median: a small fixed sample
    would be a noisy estimate of the fleet's rate).  value = all solves / all time."""
    import oracle
    o = oracle.load()
    cores = int(o.raw.oracle_get_threads())
    for i in range(max(warmup, 0)):
        spec, x0, u, theta, idx = fleet_sample(P, min(per_step, 4), shift=1000 + i)
        o.ce_costs(spec, x0, u, theta, 0.1, P=len(idx))
    times, solves, seen = [], 0, set()
    for i in range(steps):
        spec, x0, u, theta, idx = fleet_sample(P, per_step, shift=i)
        t0 = time.perf_counter()
        o.ce_costs(spec, x0, u, theta, 0.1, P=len(idx))
        times.append(time.perf_counter() - t0)
        solves += theta.size
        seen.update(int(j) for j in idx)
    total = float(np.sum(times))
    sample = (f"{steps} step(s) x {per_step} problems x {THETAS} theta taken evenly from the {P}-problem fleet "
              f"({len(seen)} distinct problems, {solves} solves, {total:.1f} s on {cores} threads; per-step rate "
              f"{min(per_step * THETAS / t for t in times):.0f}-{max(per_step * THETAS / t for t in times):.0f} solves/s)")
    return solves / total, cores, total / steps, sample, len(seen)


def latency_cases(be, peak_tf, no_cpu):
    """The small-batch cases of BASELINE.json (wall clock around blocking C-ABI calls, host buffers in and out), the CPU
    oracle on all host threads beside each, and roofline-shaped figures for the n = 12 solve kernel and the PETS kernel."""
    import oracle
    import ratilqr_b200 as R
    from ratilqr_b200 import cross_entropy as CE
    from ratilqr_b200 import nelder_mead as NM
    from ratilqr_b200 import workloads as wl
    o = None if no_cpu else oracle.load()

    def timed(fn, reps=5):
        fn()
        t = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
        return min(t) * 1e3

    lat = {}
    prob, x0, u = wl.c2_problem()
    spec = prob.spec()
    ua = [u[:, k].copy() for k in range(u.shape[1])]
    th = wl.c2_thetas(THETAS)
    g = timed(lambda: be.ce_costs(spec, x0, u, th, 0.1))
    lat["c2_single_1x1024"] = {"what": "configs[1] exactly through ratilqr_ce_costs (compute_cost), host buffers", "gpu_ms": g,
                               "gpu_solves_per_sec": THETAS / g * 1e3,
                               "cpu_oracle_ms": None if no_cpu else timed(lambda: o.ce_costs(spec, x0, u, th, 0.1), 2)}
    g = timed(lambda: be.ce_solve(spec, x0, u, 0.1, 1.0, 2.0, seed=3))

    def cpu_mpc():
        s = R.CrossEntropyBilevelOptimizationSolver(backend=o)
        return CE.solve_(s, prob, x0, ua, np.random.default_rng(1), kl_bound=0.1)

    lat["mpc_step_single_problem"] = {"what": "one RAT iLQR solve! (10 theta x 5 CE iterations + final solve), ratilqr_ce_solve: whole loop on the device",
                                      "gpu_ms": g, "cpu_oracle_ms": None if no_cpu else timed(cpu_mpc, 3)}
    p3, x03, u3 = wl.c3_problem()
    ua3 = [u3[:, k].copy() for k in range(u3.shape[1])]

    def nm(backend):
        s = R.NelderMeadBilevelOptimizationSolver(backend=backend)
        return NM.solve_(s, p3, x03, ua3, kl_bound=0.1)

    lat["c3_rat_ilqr_pp_quadrotor"] = {"what": "configs[2]: RAT iLQR++ (Nelder-Mead defaults) on the 12-state quadrotor, T = 40, single problem",
                                       "gpu_ms": timed(lambda: nm(be), 2), "cpu_oracle_ms": None if no_cpu else timed(lambda: nm(o), 2)}
    if not no_cpu:
        lat["cpu_threads"] = int(o.raw.oracle_get_threads())
    # n = 12 solve kernel: 888 quadrotor iLEQG solves (6 per SM), algorithmic flops from the device counters
    s3 = p3.spec()
    B3 = 888
    th3 = np.concatenate([[0.0], wl.positive_thetas(B3 - 1, mu=0.012, sigma=0.006, key=B3)])
    be.stage(s3, x03, u3, th3)
    be.run(1)
    ms3 = be.run(3) / 3
    r3 = be.fetch()
    N3, F_LIN_Q, F_F_Q = 40, 400.0, 120.0  # quadrotor: Jacobians + quadratic cost derivatives, one dynamics step
    it, tr = r3["iters"].astype(np.float64), r3["trials"].astype(np.float64)
    fl3 = float(np.sum(N3 * (it * f_opt(12, 4) + (1 + tr) * f_eval(12, 4)) + (1 + it + tr) * N3 * F_LIN_Q + (1 + tr) * N3 * F_F_Q))
    roof_c3 = {"bound": "fp64", "kernel": "k_ileqg_solve_coop<quadrotor, quadratic> (one warp per instance)", "achieved": fl3 / (ms3 * 1e-3) / 1e12,
               "peak": peak_tf, "unit": "TFLOP/s", "frac": fl3 / (ms3 * 1e-3) / 1e12 / peak_tf, "instances": B3, "feasible": int((r3["status"] == 0).sum()),
               "ms_per_launch": ms3, "solves_per_sec": B3 / ms3 * 1e3, "flops_per_launch": fl3,
               "flops_per_stage": {"F_opt(12,4)": f_opt(12, 4), "F_eval(12,4)": f_eval(12, 4), "F_lin": F_LIN_Q, "F_f": F_F_Q}, "traffic": None}
    # PETS: configs[3] in full, 5 CEM iterations; per particle N (F_f + F_c) + F_h flops (cart-pole step 45, quadratic cost 20)
    p4, x04 = wl.c4_problem()
    s4, gen = p4.spec(), p4.f_stochastic.gen()
    mu0, Sg0 = np.zeros((1, 30)), np.tile(np.array([[4.0]])[:, :, None], (1, 1, 30))
    ms4 = timed(lambda: be.pets_solve(s4, x04, mu0, Sg0, 4096, 150, 409, 5, 0.1, seed=1, gen=gen), 3)
    fl4 = 4096.0 * 150 * 5 * (30 * (45.0 + 20.0) + 10.0)
    roof_c4 = {"bound": "fp64", "kernel": "k_pets_costs<cartpole, quadratic> (+ Philox / Box-Muller noise, not counted)", "achieved": fl4 / (ms4 * 1e-3) / 1e12,
               "peak": peak_tf, "unit": "TFLOP/s", "frac": fl4 / (ms4 * 1e-3) / 1e12 / peak_tf, "ms_per_solve": ms4,
               "rollouts_per_sec": 4096 * 150 * 5 / ms4 * 1e3, "flops_per_solve": fl4,
               "note": "end to end through ratilqr_pets_solve (sample, rollout, rank, refit x 5): the rollout kernel is SFU/latency bound "
                       "(sincos of the dynamics, log/sqrt/sincos of Box-Muller), not FMA bound", "traffic": None}
    return lat, roof_c3, roof_c4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems", type=int, default=1776, help="independent unicycle problems per GPU (x 1024 theta each)")
    ap.add_argument("--no-profile-warm", action="store_true", help="stage the timed launch without a previous call's work profile")
    ap.add_argument("--fleet-problems", type=int, default=8192, help="RAT iLQR problems per GPU for the MPC-step figure")
    ap.add_argument("--cpu-sample-problems", type=int, default=0, help="problems per CPU step; 0 = 8 per step for the reference arm (a different selection every step), 48 for the in-run baseline")
    ap.add_argument("--no-latency-cases", action="store_true")
    ap.add_argument("--distinct-fleets", action="store_true", help="every rank draws its own problems instead of the same problem set with its own theta populations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global DISTINCT_FLEETS
    DISTINCT_FLEETS = args.distinct_fleets
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "c2_fleet", "model": None, "dynamics": "unicycle n=4 m=2", "horizon": HORIZON,
              "thetas_per_problem": THETAS, "problems_per_gpu": args.problems,
              "solves_per_step_per_gpu": args.problems * THETAS, "kl_bound": 0.1,
              "l2_policy": f"inputs_larger_than_l2 ({args.problems * THETAS * 8864 / 1e6:.0f} MB SoA workspace per step vs 126 MB L2)",
              "parallelism": f"dp{world}",
              "rank_inputs": ("every rank draws its own problems and theta populations" if args.distinct_fleets else
                              "every rank owns the same problem definitions with its own theta populations (equal per-GPU work)"),
              "slot_order": ("natural" if args.no_profile_warm else
                             "problems heaviest-first by the iterations a previous call (different theta population) needed")}
    config.pop("model")

    if args.impl == "reference":
        if rank != 0:
            return
        value, cores, dt, sample, ns = cpu_arm(args.problems, max(args.steps, 1), min(args.warmup, 1), args.cpu_sample_problems or 8)
        config["reference_arm_distinct_sample_problems"] = ns
        config["reference_arm"] = "in-repo C++ oracle (restatement of the Julia reference; no Julia toolchain in the image)"
        print(json.dumps({"impl": "reference", "metric": "batched_ileqg_solves_per_sec", "value": value, "unit": "solves/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample,
                                           "note": "C++ oracle (no Julia toolchain in the image); faster than Julia+ForwardDiff"},
                          "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist

    import ratilqr_b200 as R

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    be = R.new_backend(local_rank)  # raises if the CUDA library is missing: no CPU fallback
    P = args.problems
    spec, x0, u, theta = build_inputs(P, rank)
    B = theta.size

    # ---- kernel-only throughput, inputs resident in HBM ------------------------------------------------
    # steady state of a CE loop: one earlier evaluation of the same fleet with a DIFFERENT theta population leaves the
    # per-problem work profile that orders this launch heaviest-first (rl_capi.cu "work profile"; --no-profile-warm skips it)
    from ratilqr_b200 import workloads as wl_
    if not args.no_profile_warm:
        be.ce_costs(spec, x0, u, wl_.positive_thetas(B, key=8999 + 100000 * rank), 0.1, P=P)
    be.stage(spec, x0, u, theta, P=P)
    be.run(warmup)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = be.launch_count()
    t0 = time.perf_counter()
    ms_dev = be.run(args.steps)  # K launches back to back, CUDA events on the library's stream
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = be.launch_count() - launches0
    clocks = sampler.stop()
    ms_dev_rank = ms_dev
    ms_dev = max_over_ranks(ms_dev)
    ms_per_step = ms_dev / args.steps
    value = sum_over_ranks(float(B)) / (ms_per_step * 1e-3)
    res = be.fetch()
    ok = int(np.sum(res["status"] == 0))
    flops_rank = float(np.sum(algorithmic_flops(res["iters"], res["trials"])))
    byts_rank = float(np.sum(algorithmic_bytes(res["iters"], res["trials"])))
    # whole-job roofline: work summed over the ranks / the slowest rank's time (each rank draws its own fleet, so the
    # ranks' work differs by a few percent: reported per rank below)
    flops, byts = sum_over_ranks(flops_rank), sum_over_ranks(byts_rank)
    per_rank = None
    if world > 1:
        mine = torch.tensor([ms_dev_rank / args.steps, float(res["iters"].mean()), flops_rank, float(ok)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rank": r, "ms_per_step": float(t[0]), "mean_iters": float(t[1]), "algorithmic_flops": float(t[2]),
                     "converged": int(t[3])} for r, t in enumerate(allr)]

    # ---- end to end through the reference-facing C-ABI call with host buffers --------------------------
    # Every call evaluates a FRESH theta population for the same fleet -- what consecutive CE iterations do
    # (cross_entropy_bilevel_optimization.jl:291-334); the library orders the problems by the iterations they needed in
    # the previous call (rl_capi.cu "work profile"), which is only a prediction here, as in real use.
    pops = [wl_.positive_thetas(B, key=9000 + 97 * i + 100000 * rank) for i in range(2 + args.steps)]
    for i in range(2):
        be.ce_costs(spec, x0, u, pops[i], 0.1, P=P)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        cost, st = be.ce_costs(spec, x0, u, pops[2 + i], 0.1, P=P)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e_value = sum_over_ranks(float(B)) / e2e_s
    h2d = int(x0.nbytes + u.nbytes + theta.nbytes + spec.cost_params.nbytes + spec.W.nbytes * 3)
    d2h = int(cost.nbytes + st.nbytes)

    # ---- second half of the metric: RAT iLQR MPC step latency (configs[4]: fleet of independent unicycle problems,
    #      CE defaults 10 theta x 5 iterations + final solve, whole loop on the device, host buffers in/out) ----------
    Pf = args.fleet_problems
    from ratilqr_b200 import workloads as wl
    fprob, fcps, fx0, fu = wl.fleet(Pf, key=70 + (rank if args.distinct_fleets else 0))
    fspec = fprob.spec(cost_params=fcps)
    be.ce_solve_fleet(fspec, fx0, fu, 0.1, 1.0, 2.0, seed=7 + rank, want=())  # warm-up (allocations)
    fleet_all = []
    for _ in range(3):  # latency-bound (six sequential solves per problem): report the best of three and all three
        barrier()
        t0 = time.perf_counter()
        fr = be.ce_solve_fleet(fspec, fx0, fu, 0.1, 1.0, 2.0, seed=7 + rank, want=("l",))
        barrier()
        fleet_all.append(max_over_ranks(time.perf_counter() - t0))
    fleet_s = min(fleet_all)
    fleet_ok = sum_over_ranks(float((fr["status"] == 0).sum()))
    # configs[4] "256 MC samples each": noisy closed-loop rollouts of every problem's optimised policy (Philox noise
    # coloured with chol(W)), host policy buffers in, J + per-problem statistics out
    MC = 256
    fr2 = be.ce_solve_fleet(fspec, fx0, fu, 0.1, 1.0, 2.0, seed=7 + rank, want=("x", "l", "L"))
    be.mc_rollout(fspec, fr2["x"], fr2["l"], fr2["L"], MC, seed=11 + rank, P=Pf)  # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    mc = be.mc_rollout(fspec, fr2["x"], fr2["l"], fr2["L"], MC, seed=11 + rank, P=Pf)
    barrier()
    mc_s = max_over_ranks(time.perf_counter() - t0)
    mc_finite = sum_over_ranks(float(np.isfinite(mc["stats"][:, 0]).sum()))
    # the receding-horizon driver: 4 consecutive MPC steps of the same fleet without leaving the device (plan, true-system
    # step with Philox disturbances, plan shift: ratilqr_mpc_fleet_run); steps after the first are warm-started
    barrier()
    mr = be.mpc_fleet_run(fspec, fx0, fu, 4, 0.1, 1.0, 2.0, seed=7 + rank, noise_seed=13 + rank)
    barrier()
    mpc_dev_ms = [max_over_ranks(float(t)) for t in mr["ms"]]

    # ---- the path's one real exchange step at scale: ONE problem's 65,536-theta population sharded over the ranks, the cost
    #      vector all-gathered on device buffers inside the library (ratilqr_ce_costs_sharded); N = 1: plain ratilqr_ce_costs
    from ratilqr_b200 import distributed as D_
    pp, px0, pu = wl.c2_problem()
    big_theta = wl.positive_thetas(65536, key=65536)
    sharded = None
    try:
        D_.sharded_ce_costs(be, pp.spec(), px0, pu, big_theta, 0.1)  # warm-up: attaches the library's communicator, allocations
        ts = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            big_cost = D_.sharded_ce_costs(be, pp.spec(), px0, pu, big_theta, 0.1)
            barrier()
            ts.append(max_over_ranks(time.perf_counter() - t0))
        sharded = {"workload": "configs[1]'s problem with a 65,536-theta population, block-sharded over the ranks; one in-library "
                               "ncclAllGather of (value, status) on device buffers; every rank ends up with the whole cost vector",
                   "thetas": 65536, "ms": min(ts) * 1e3, "solves_per_sec": 65536 / min(ts), "finite_costs": int(np.isfinite(big_cost).sum())}
    except Exception as e:  # noqa: BLE001  (secondary figure: never take the headline down with it)
        sharded = {"error": str(e)[:200]}

    out = None
    if rank == 0:
        peak_tf = be.fp64_probe()
        peak_sus = be.fp64_probe_sustained(1.0)
        ach_tf = flops / (ms_per_step * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None  # dram__bytes_read+write of the solve kernel from the committed ncu --set full capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "solve_traffic.json")))
            for ent in tj.get("entries", [tj]):
                if ent.get("problems_per_gpu") == P:
                    traffic = ent["dram_bytes_per_launch"]
        except OSError:
            pass
        ach_gbs = byts / (ms_per_step * 1e-3) / 1e9
        # the exact configs[1] shape: ONE problem x 1024 theta (latency-bound: the speculative kernel, rl_spec.cuh)
        from ratilqr_b200 import workloads as wl1
        p1, x01, u1 = wl1.c2_problem()
        be.stage(p1.spec(), x01, u1, wl1.c2_thetas(THETAS))
        be.run(3)
        ms1 = be.run(args.steps) / args.steps
        nominal_tf = 148 * 64 * 2 * 1.965e9 / 1e12              # SMs x FP64 lanes x 2 flops x boost clock (B200 datasheet-level)
        clock_tf = 148 * 64 * 2 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12   # the same at the SM clock measured under load
        latency, roof_c3, roof_c4 = None, None, None
        if not args.no_latency_cases:
            latency, roof_c3, roof_c4 = latency_cases(be, peak_tf, args.no_cpu_baseline)
        out = {"metric": "batched_ileqg_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world,
               "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
               "converged_instances": ok, "instances": int(B), "wall_ms_per_step": wall_ms / args.steps,
               "mean_iters": float(res["iters"].mean()), "mean_trials": float(res["trials"].mean()),
               "clocks": clocks, "gpu_launches": int(launches),
               "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "api": "ratilqr_ce_costs (compute_cost), host buffers in, cost+status vectors out; a fresh theta population per step"},
               "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf * world, "unit": "TFLOP/s", "frac": ach_tf / (peak_tf * world),
                            "peak_per_gpu": peak_tf,
                            "traffic": traffic, "algorithmic_bytes_per_launch": byts, "kernel": "k_ileqg_solve<unicycle, quadratic>",
                            "peak_source": "burst DFMA probe kernel run in this process (MEASURED_PEAKS.json has no FP64 figure)",
                            "peak_sustained": peak_sus * world, "frac_of_sustained": ach_tf / (peak_sus * world),
                            "peak_nominal": nominal_tf * world, "frac_of_nominal": ach_tf / (nominal_tf * world),
                            "peak_at_measured_clock": clock_tf * world, "frac_of_peak_at_measured_clock": ach_tf / (clock_tf * world),
                            "flops_per_launch": flops,
                            "flops_basis": "work summed over all ranks / slowest rank's time" if world > 1 else "this GPU"},
               "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak * world, "unit": "GB/s", "frac": ach_gbs / (hbm_peak * world),
                                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"},
               "c2_single": {"workload": "configs[1] exactly: 1 problem x 1024 theta (kernel only, CUDA events; speculative latency kernel)", "ms_per_batch": ms1,
                             "solves_per_sec": THETAS / (ms1 * 1e-3)},
               "mpc_step": {"workload": f"configs[4]: fleet of {Pf} independent RAT iLQR unicycle problems per GPU (CE: 10 theta x 5 "
                                        "iterations + final solve), ratilqr_ce_solve_fleet, host buffers in, theta_opt/value/l out",
                            "ms_per_fleet_step": fleet_s * 1e3, "ms_all": [t * 1e3 for t in fleet_all], "problems_per_sec": Pf * world / fleet_s,
                            "us_per_problem_step": fleet_s * 1e6 / (Pf * world), "ce_rounds": fr["rounds"],
                            "final_solves_ok": int(fleet_ok), "problems": Pf * world,
                            "on_device_driver_ms_per_step": mpc_dev_ms,
                            "on_device_driver": "ratilqr_mpc_fleet_run, 4 receding-horizon steps (cold start, then warm-started by the shifted plans)",
                            "mc_eval": {"samples_per_problem": MC, "ms": mc_s * 1e3, "rollouts_per_sec": MC * Pf * world / mc_s,
                                        "problems_with_finite_mean": int(mc_finite),
                                        "call": "ratilqr_mc_rollout, host policy buffers in, J + stats out, Philox noise"}}}
        out["sharded_population"] = sharded
        if per_rank is not None:
            out["per_rank"] = per_rank
        if latency is not None:
            out["latency"], out["roofline_c3"], out["roofline_c4"] = latency, roof_c3, roof_c4
        if not args.no_cpu_baseline:
            v, cores, dt, sample, _ = cpu_arm(P, 1, 0, args.cpu_sample_problems or 48)
            out["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample,
                                   "what": "in-repo C++ oracle (restatement of the Julia reference), not Julia itself"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    be.close()
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
