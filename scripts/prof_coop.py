import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ratilqr_b200 as R
from ratilqr_b200 import workloads as wl
be = R.new_backend(0)
prob, x0, u = wl.c3_problem()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
th = np.concatenate([[0.0], wl.positive_thetas(B - 1, mu=0.02, sigma=0.02, key=B)])
for _ in range(2):
    be.ce_costs(prob.spec(), x0, u, th, 0.1)
