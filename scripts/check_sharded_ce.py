"""2-GPU (or N-GPU) check of the single-population sharding: every rank solves a block of the theta samples on
its own GPU, ONE all_gather of the cost vector over NCCL, and every rank must hold exactly the single-GPU result.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded_ce.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ratilqr_b200 as R  # noqa: E402
from ratilqr_b200 import distributed as D  # noqa: E402
from ratilqr_b200 import workloads as wl  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
be = R.new_backend(local)
prob, x0, u = wl.c2_problem()
spec = prob.spec()
theta = wl.c2_thetas(1024)
full = D.sharded_ce_costs(be, spec, x0, u, theta, 0.1)  # warm-up + result
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
full = D.sharded_ce_costs(be, spec, x0, u, theta, 0.1)
dist.barrier(); torch.cuda.synchronize()
dt = time.perf_counter() - t0
single = be.ce_costs(spec, x0, u, theta, 0.1)[0]
ok = bool(np.array_equal(full, single))
flags = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"check": "sharded_ce_costs: all-gather inside the C library (ratilqr_attach_comm + ratilqr_ce_costs_sharded), one process per GPU", "world": world, "thetas": 1024, "identical_on_all_ranks": bool(flags.item() == 1.0),
                      "ms_sharded": dt * 1e3, "elite_theta": float(theta[np.argsort(full, kind="stable")[0]])}))

# ---- PETS (configs[3]): action sequences sharded, injected noise => bit-identical costs + redundant refit
pprob, px0 = wl.c4_problem()
pspec, gen = pprob.spec(), pprob.f_stochastic.gen()
rng = np.random.default_rng(3)
C_, Kp, N_ = 1024, 30, pspec.N
controls = rng.standard_normal((1, N_, C_))
noise = 1e-2 * rng.standard_normal((4, N_, Kp, C_))
mu, Sg = np.zeros((1, N_)), np.ones((1, 1, N_))
D.sharded_pets_step(be, pspec, gen, px0, mu, Sg, controls, Kp, 102, 0.1, noise=noise)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
mu_n, Sg_n, idx, cost = D.sharded_pets_step(be, pspec, gen, px0, mu, Sg, controls, Kp, 102, 0.1, noise=noise)
dist.barrier(); torch.cuda.synchronize()
dt = time.perf_counter() - t0
cost1 = be.pets_costs(pspec, px0, controls, Kp, noise=noise, gen=gen)
mu1, Sg1, idx1 = be.pets_refit(controls, cost1, 102, 0.1, mu, Sg)
ok = bool(np.array_equal(cost, cost1) and np.array_equal(idx, idx1) and np.array_equal(mu_n, mu1) and np.array_equal(Sg_n, Sg1))
flags = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"check": "sharded_pets_step over NCCL", "world": world, "sequences": C_, "particles": Kp,
                      "identical_on_all_ranks": bool(flags.item() == 1.0), "ms_sharded": dt * 1e3}))
dist.destroy_process_group()
be.close()
