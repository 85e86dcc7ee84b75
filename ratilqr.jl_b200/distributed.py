"""Multi-GPU plumbing: one process per GPU (torch.distributed), block partition of independent units.

The path shards naturally: iLEQG instances, MC samples, action sequences and whole
problems are independent.
  * fleet (many problems): each rank owns a block of problems, NO collective;
  * one bilevel problem with a big theta population: each rank solves a block of the theta samples,
    then ONE all_gather of the cost vector (num_samples doubles) per CE iteration -- the analogue of
    the reference's `remotecall_fetch` gather (cross_entropy_bilevel_optimization.jl:186-193).  Elite
    selection runs redundantly (and deterministically) on every rank, so no broadcast is needed.
"""
import numpy as np


def block_range(count, rank, world):
    """contiguous block [lo, hi) of `count` units owned by `rank` (sizes differ by at most one)"""
    lo = (count * rank) // world
    hi = (count * (rank + 1)) // world
    return lo, hi


def sharded_ce_costs(be, spec, x0, u_init, theta, kl_bound, opts=None, group=None):
    """compute_cost over a theta population split across the ranks of `group`; every rank returns the full
    cost vector.  Works with the nccl backend (GPU tensors) and gloo (CPU tensors, used by the tests)."""
    import torch
    import torch.distributed as dist
    theta = np.asarray(theta, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return be.ce_costs(spec, x0, u_init, theta, kl_bound, opts=opts)[0]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = block_range(theta.size, rank, world)
    local = be.ce_costs(spec, x0, u_init, theta[lo:hi], kl_bound, opts=opts)[0] if hi > lo else np.zeros(0)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    sizes = [block_range(theta.size, r, world) for r in range(world)]
    width = max(h - l for l, h in sizes)
    buf = torch.full((width,), float("nan"), dtype=torch.float64, device=dev)
    buf[: hi - lo] = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)  # num_samples doubles: latency-bound on NVLink 5 / NVSwitch
    return np.concatenate([o[: h - l].cpu().numpy() for o, (l, h) in zip(out, sizes)])


def fleet_block(P, rank, world):
    """problems owned by `rank` in a fleet of P independent problems (no data-path collective)"""
    return block_range(P, rank, world)
