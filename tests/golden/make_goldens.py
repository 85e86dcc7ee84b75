"""Generates tests/golden/*.npz.  The reference (Julia) cannot run in this image, so these vectors come from
the CPU oracle (oracle/oracle.cpp), which is itself pinned against the reference's known-answer tests
(tests/test_reference_*.py) and against the survey's independent numpy restatement.
They freeze today's behaviour so that later kernel work cannot drift silently.
    python tests/golden/make_goldens.py
A maintainer with Julia >= 1.5 can overwrite them with true-reference output using julia/export_goldens.jl."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from ratilqr_b200 import workloads as wl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
o = oracle.load()


def save(name, spec_args, res, **extra):
    keep = {k: res[k] for k in ("value", "status", "iters", "trials", "restarts", "x", "l", "L") if k in res}
    np.savez_compressed(os.path.join(HERE, name), **keep, **spec_args, **extra)


# C1: the shipped test problem at the thetas the reference's tests use, plus the breakdown boundary
prob, x0, u = wl.c1_problem()
th = np.array([0.0, 0.1, 0.3, 0.43, 0.5, 30.7, 30.9])
save("c1_power_law.npz", dict(x0=x0, u=u, theta=th), o.ileqg_solve_batch(prob.spec(), x0, u, th))

# C2: first 48 thetas of the 1024-sample population
prob, x0, u = wl.c2_problem()
th = wl.c2_thetas(1024)[:48]
save("c2_unicycle_48.npz", dict(x0=x0, u=u, theta=th), o.ileqg_solve_batch(prob.spec(), x0, u, th))

# fleet: 4 problems x 6 thetas with per-problem x0 / goal
prob, cps, x0, u = wl.fleet(4, N=25)
th = wl.positive_thetas(24, key=11)
save("fleet_4x6.npz", dict(x0=x0, u=u, theta=th, cost_params=cps), o.ileqg_solve_batch(prob.spec(cost_params=cps), x0, u, th, P=4))

# C3: quadrotor, short horizon
prob, x0, u = wl.c3_problem(N=12)
th = np.array([0.0, 0.05, 0.5])
save("c3_quadrotor_N12.npz", dict(x0=x0, u=u, theta=th), o.ileqg_solve_batch(prob.spec(), x0, u, th))
print("goldens written to", HERE)
