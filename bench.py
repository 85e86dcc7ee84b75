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
extra  : the exact configs[1] shape (1 problem x 1024 theta, latency-bound) is reported under "c2_single".
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


def build_inputs(P, rank):
    from ratilqr_b200 import workloads as wl
    prob, cps, x0, u = wl.fleet(P, key=7 + 1000 * rank)
    theta = np.concatenate([wl.positive_thetas(THETAS, key=20201028 + p + 100000 * rank) for p in range(P)])
    return prob.spec(cost_params=cps), x0, u, theta


def cpu_arm(spec_fn, sample_problems, steps, warmup):
    """the oracle (CPU restatement of the reference) on all host threads, bounded sample of the same workload"""
    import oracle
    o = oracle.load()
    cores = int(o.raw.oracle_get_threads())
    spec, x0, u, theta = spec_fn(sample_problems)
    n_solves = theta.size
    for _ in range(max(warmup, 0)):
        o.ce_costs(spec, x0, u, theta, 0.1, P=sample_problems)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.ce_costs(spec, x0, u, theta, 0.1, P=sample_problems)
    dt = (time.perf_counter() - t0) / steps
    return n_solves / dt, cores, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems", type=int, default=1776, help="independent unicycle problems per GPU (x 1024 theta each)")
    ap.add_argument("--no-profile-warm", action="store_true", help="stage the timed launch without a previous call's work profile")
    ap.add_argument("--fleet-problems", type=int, default=8192, help="RAT iLQR problems per GPU for the MPC-step figure")
    ap.add_argument("--cpu-sample-problems", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "c2_fleet", "model": None, "dynamics": "unicycle n=4 m=2", "horizon": HORIZON,
              "thetas_per_problem": THETAS, "problems_per_gpu": args.problems,
              "solves_per_step_per_gpu": args.problems * THETAS, "kl_bound": 0.1,
              "l2_policy": f"inputs_larger_than_l2 ({args.problems * THETAS * 8864 / 1e6:.0f} MB SoA workspace per step vs 126 MB L2)",
              "parallelism": f"dp{world}",
              "slot_order": ("natural" if args.no_profile_warm else
                             "problems heaviest-first by the iterations a previous call (different theta population) needed")}
    config.pop("model")

    if args.impl == "reference":
        if rank != 0:
            return
        value, cores, dt = cpu_arm(lambda P: build_inputs(P, 0), args.cpu_sample_problems, max(args.steps, 1), min(args.warmup, 1))
        sample = f"{args.cpu_sample_problems} problems x {THETAS} theta = {args.cpu_sample_problems * THETAS} solves per step of the same workload"
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
    ms_dev = max_over_ranks(ms_dev)
    ms_per_step = ms_dev / args.steps
    value = sum_over_ranks(float(B)) / (ms_per_step * 1e-3)
    res = be.fetch()
    ok = int(np.sum(res["status"] == 0))
    flops = float(np.sum(algorithmic_flops(res["iters"], res["trials"])))
    byts = float(np.sum(algorithmic_bytes(res["iters"], res["trials"])))

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
    fprob, fcps, fx0, fu = wl.fleet(Pf, key=70 + rank)
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
        # the exact configs[1] shape: ONE problem x 1024 theta (latency-bound: 32 warps on a 148-SM device)
        s1, x01, u1, th1 = build_inputs(1, 0)
        be.stage(s1, x01, u1, th1, P=1)
        be.run(3)
        ms1 = be.run(args.steps) / args.steps
        out = {"metric": "batched_ileqg_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world,
               "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
               "converged_instances": ok, "instances": int(B), "wall_ms_per_step": wall_ms / args.steps,
               "mean_iters": float(res["iters"].mean()), "mean_trials": float(res["trials"].mean()),
               "clocks": clocks, "gpu_launches": int(launches),
               "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "api": "ratilqr_ce_costs (compute_cost), host buffers in, cost+status vectors out; a fresh theta population per step"},
               "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                            "traffic": traffic, "algorithmic_bytes_per_launch": byts, "kernel": "k_ileqg_solve<unicycle, quadratic>",
                            "peak_source": "burst DFMA probe kernel run in this process (MEASURED_PEAKS.json has no FP64 figure)",
                            "peak_sustained": peak_sus, "frac_of_sustained": ach_tf / peak_sus,
                            "flops_per_launch": flops},
               "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"},
               "c2_single": {"workload": "configs[1] exactly: 1 problem x 1024 theta", "ms_per_batch": ms1,
                             "solves_per_sec": THETAS / (ms1 * 1e-3)},
               "mpc_step": {"workload": f"configs[4]: fleet of {Pf} independent RAT iLQR unicycle problems per GPU (CE: 10 theta x 5 "
                                        "iterations + final solve), ratilqr_ce_solve_fleet, host buffers in, theta_opt/value/l out",
                            "ms_per_fleet_step": fleet_s * 1e3, "ms_all": [t * 1e3 for t in fleet_all], "problems_per_sec": Pf * world / fleet_s,
                            "us_per_problem_step": fleet_s * 1e6 / (Pf * world), "ce_rounds": fr["rounds"],
                            "final_solves_ok": int(fleet_ok), "problems": Pf * world,
                            "mc_eval": {"samples_per_problem": MC, "ms": mc_s * 1e3, "rollouts_per_sec": MC * Pf * world / mc_s,
                                        "problems_with_finite_mean": int(mc_finite),
                                        "call": "ratilqr_mc_rollout, host policy buffers in, J + stats out, Philox noise"}}}
        if not args.no_cpu_baseline:
            v, cores, dt = cpu_arm(lambda PP: build_inputs(PP, 0), args.cpu_sample_problems, 1, 0)
            out["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                                   "sample": f"{args.cpu_sample_problems} problems x {THETAS} theta of the same workload, {dt:.2f} s wall"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    be.close()
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
