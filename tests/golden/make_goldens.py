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

# user-extensible models: the drag-car dynamics + obstacle cost snippets of tests/user_models/ (no registered
# equivalent, so no oracle): frozen from the g++ build of the snippets (tests/_hostemu), which is itself checked against
# complex-step / finite-difference derivatives of a numpy restatement (tests/test_user_models.py)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _hostemu  # noqa: E402
from ratilqr_b200 import _capi  # noqa: E402

h = _hostemu.load()
prob, x0, u = wl.c2_problem(N=30)
ref = prob.spec()
QD, RD, XG = np.array([1.0, 1.0, 0.1, 0.1]), np.array([0.1, 0.1]), np.array([5.0, 5.0, 0.0, 0.0])
obst_cp = np.concatenate([0.01 * QD, 0.01 * RD, XG, [10.0], [2.5, 2.0, 0.8, 0.3]])
drag_p = np.array([0.1, 0.05, 1.5])
spec = _capi.Spec(1001, 101, 4, 2, 30, drag_p, obst_cp, ref.W.reshape(4, 4, order="F"))  # hostemu numbering of the snippets
th = np.concatenate([[0.0], wl.positive_thetas(15, key=21)])
save("user_dragcar_obstacle.npz", dict(x0=x0, u=u, theta=th, cost_params=obst_cp, model_params=drag_p, W=ref.W.reshape(4, 4, order="F")),
     h.ileqg_solve_batch(spec, x0, u, th))
