"""One staged launch of configs[1] exactly (1 problem x K theta, unicycle T = 50) for ncu captures of the latency regime.
    ncu --set full -k regex:k_ileqg_solve -c 1 ... python scripts/prof_single.py [K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ratilqr_b200 as R  # noqa: E402
from ratilqr_b200 import workloads as wl  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
be = R.new_backend(0)
prob, x0, u = wl.c2_problem()
be.stage(prob.spec(), x0, u, wl.c2_thetas(K))
print("ms", be.run(1))
be.close()
