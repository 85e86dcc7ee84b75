"""N>1 host logic on CPU: world_size-2 gloo processes shard a theta population, all_gather the cost vector
and must reproduce the single-process result bit for bit (the oracle stands in for the device library here)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle
    from ratilqr_b200 import distributed as D
    from ratilqr_b200 import workloads as wl
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    be = oracle.load()
    be.raw.oracle_set_threads(1)
    prob, x0, u = wl.c1_problem()
    theta = wl.positive_thetas(13, mu=1.0, sigma=2.0, key=5)  # 13: uneven split across 2 ranks
    full = D.sharded_ce_costs(be, prob.spec(), x0, u, theta, 1.0)
    lo, hi = D.fleet_block(7, rank, world)
    q.put((rank, full, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ce_costs_world2():
    sys.path.insert(0, ROOT)
    import oracle
    from ratilqr_b200 import workloads as wl
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prob, x0, u = wl.c1_problem()
    theta = wl.positive_thetas(13, mu=1.0, sigma=2.0, key=5)
    ref = oracle.load().ce_costs(prob.spec(), x0, u, theta, 1.0)[0]
    blocks = {}
    for rank, full, blk in got:
        assert np.array_equal(full, ref)  # every rank holds the full, identical cost vector
        blocks[rank] = blk
    assert blocks[0] == (0, 3) and blocks[1] == (3, 7)  # fleet partition covers all problems exactly once


def _pets_inputs():
    from ratilqr_b200 import workloads as wl
    prob, x0 = wl.c4_problem(N=8)
    rng = np.random.default_rng(11)
    C_, Kp = 11, 10  # 11: uneven split across 2 ranks; 10 particles = 2 per ensemble member
    controls = rng.standard_normal((1, 8, C_))
    noise = 1e-2 * rng.standard_normal((4, 8, Kp, C_))
    mu, Sg = np.zeros((1, 8)), np.ones((1, 1, 8))
    return prob, x0, controls, noise, mu, Sg, Kp


def _pets_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle
    from ratilqr_b200 import distributed as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    be = oracle.load()
    be.raw.oracle_set_threads(1)
    prob, x0, controls, noise, mu, Sg, Kp = _pets_inputs()
    out = D.sharded_pets_step(be, prob.spec(), prob.f_stochastic.gen(), x0, mu, Sg, controls, Kp, 3, 0.1, noise=noise)
    q.put((rank,) + tuple(out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_pets_step_world2():
    """action sequences sharded over 2 ranks + all_gather of the cost vector + redundant refit == one process"""
    sys.path.insert(0, ROOT)
    import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pets_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    be = oracle.load()
    prob, x0, controls, noise, mu, Sg, Kp = _pets_inputs()
    cost = be.pets_costs(prob.spec(), x0, controls, Kp, noise=noise, gen=prob.f_stochastic.gen())
    mu_r, Sg_r, idx_r = be.pets_refit(controls, cost, 3, 0.1, mu, Sg)
    assert np.all(np.isfinite(cost))
    for rank, mu_n, Sg_n, idx, c in got:
        assert np.array_equal(c, cost) and np.array_equal(idx, idx_r)
        assert np.array_equal(mu_n, mu_r) and np.array_equal(Sg_n, Sg_r)


def test_block_range_partition():
    sys.path.insert(0, ROOT)
    from ratilqr_b200.distributed import block_range
    for count in (0, 1, 7, 1024, 65536):
        for world in (1, 2, 3, 8):
            edges = [block_range(count, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == count
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in edges]
            assert max(sizes) - min(sizes) <= 1
