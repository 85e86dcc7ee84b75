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


def attach_library_comm(be, group=None):
    """Gives the C library its own NCCL communicator over the ranks of `group` (one process per GPU): rank 0 creates the
    NCCL unique id, torch.distributed only carries its 128 bytes to the other ranks, every rank calls ratilqr_attach_comm.
    Afterwards the *_sharded calls all-gather on device buffers inside the library -- no host bounce, no torch tensors."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1 or getattr(be, "comm_world", 1) == world:
        return be
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.tensor(list(be.nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    be.attach_comm(bytes(uid.cpu().tolist()), rank, world)
    return be


def sharded_ce_costs(be, spec, x0, u_init, theta, kl_bound, opts=None, group=None):
    """compute_cost over a theta population split across the ranks of `group`; every rank returns the full
    cost vector.  Works with the nccl backend (GPU tensors) and gloo (CPU tensors, used by the tests)."""
    import torch
    import torch.distributed as dist
    theta = np.asarray(theta, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return be.ce_costs(spec, x0, u_init, theta, kl_bound, opts=opts)[0]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if dist.get_backend(group) == "nccl" and theta.size >= world and hasattr(be, "ce_costs_sharded"):
        attach_library_comm(be, group)  # GPU ranks: the all-gather runs inside the library on device buffers
        return be.ce_costs_sharded(spec, x0, u_init, theta, kl_bound, opts=opts)[0]
    lo, hi = block_range(theta.size, rank, world)
    local = be.ce_costs(spec, x0, u_init, theta[lo:hi], kl_bound, opts=opts)[0] if hi > lo else np.zeros(0)
    return _all_gather_blocks(local, theta.size, group)  # num_samples doubles: latency-bound on NVLink 5 / NVSwitch


def _all_gather_blocks(local, count, group):
    """all_gather of a block-partitioned float64 vector; every rank returns the full vector of `count` entries"""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [block_range(count, r, world) for r in range(world)]
    lo, hi = sizes[rank]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    width = max(h - l for l, h in sizes)
    buf = torch.full((width,), float("nan"), dtype=torch.float64, device=dev)
    if hi > lo:
        buf[: hi - lo] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([o[: h - l].cpu().numpy() for o, (l, h) in zip(out, sizes)])


def sharded_pets_costs(be, spec, gen, x0, controls, particles, noise=None, seed=0, group=None):
    """PETS compute_cost (pets.jl:100-126) with the action sequences block-partitioned over the ranks of `group`
    and ONE all_gather of the cost vector (the analogue of the reference's per-worker `remotecall_fetch`,
    pets.jl:108-125).  `controls` (m, N, C) and, when given, `noise` (n, N, particles, C) are replicated inputs;
    each rank rolls out its own block of sequences.  With injected noise the result equals the one-process result
    bit for bit; with on-device Philox each rank draws from the stream `seed + rank` (statistical equivalence,
    like the reference's `randjump` streams)."""
    import torch.distributed as dist
    controls = np.asarray(controls, dtype=np.float64)
    Cn = controls.shape[2]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return be.pets_costs(spec, x0, controls, particles, noise=noise, seed=seed, gen=gen)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = block_range(Cn, rank, world)
    local = np.zeros(0)
    if hi > lo:
        nz = None if noise is None else np.asarray(noise, dtype=np.float64)[..., lo:hi]
        local = be.pets_costs(spec, x0, controls[:, :, lo:hi], particles, noise=nz, seed=seed + rank, gen=gen)
    return _all_gather_blocks(local, Cn, group)


def sharded_pets_step(be, spec, gen, x0, mu, Sigma, controls, particles, num_elite, smoothing, noise=None, seed=0,
                      group=None):
    """one PETS CEM iteration (pets.jl:193-245) over a replicated population `controls` (m, N, C): sharded rollouts,
    all_gather of the costs, then the stable top-k elite selection + smoothed refit run redundantly and
    deterministically on every rank (ties broken by global index, pets.jl:167), so (mu, Sigma) stay replicated
    without a broadcast.  Returns (mu, Sigma, elite_idx, cost)."""
    cost = sharded_pets_costs(be, spec, gen, x0, controls, particles, noise=noise, seed=seed, group=group)
    mu_n, Sg_n, idx = be.pets_refit(controls, cost, num_elite, smoothing, mu, Sigma)
    return mu_n, Sg_n, idx, cost


def fleet_block(P, rank, world):
    """problems owned by `rank` in a fleet of P independent problems (no data-path collective)"""
    return block_range(P, rank, world)
