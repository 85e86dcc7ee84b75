"""Host mirror of src/pets.jl (PETS: CEM over action sequences).

`solve_` runs the whole CEM loop on the device (`ratilqr_pets_solve`: sample -> rollout ->
particle mean -> stable top-k elites -> smoothed refit, pets.jl:193-245).  The finer functions
the reference exports are mirrored on the component entry points for the reference's tests.
Randomness: injected tensors (exact parity) or a Philox seed (statistical parity); Julia's
MersenneTwister/randjump streams cannot be reproduced.
"""
import numpy as np

from . import _lib
from .ileqg import _cols, _mats, _stack


class CrossEntropyDirectOptimizationSolver:  # a.k.a. "PETS"   pets.jl:35-68
    def __init__(self, mu_init_array, Sigma_init_array, num_control_samples=10, num_trajectory_samples=10,
                 num_elite=3, iter_max=5, smoothing_factor=0.1, backend=None):
        assert len(mu_init_array) == len(Sigma_init_array)
        self.num_control_samples, self.num_trajectory_samples = int(num_control_samples), int(num_trajectory_samples)
        self.num_elite, self.iter_max, self.smoothing_factor = int(num_elite), int(iter_max), float(smoothing_factor)
        self.mu_init_array = [np.array(v, dtype=np.float64) for v in mu_init_array]
        self.Sigma_init_array = [np.array(v, dtype=np.float64) for v in Sigma_init_array]
        self.mu_array = [v.copy() for v in self.mu_init_array]
        self.Sigma_array = [v.copy() for v in self.Sigma_init_array]
        self.N = len(mu_init_array)
        self.iter_current = 0
        self.backend = backend

    def _be(self):
        return self.backend or _lib.default_backend()


PETSSolver = CrossEntropyDirectOptimizationSolver  # north-star spelling


def initialize_(s):  # pets.jl:70-74
    s.iter_current = 0
    s.mu_array = [v.copy() for v in s.mu_init_array]
    s.Sigma_array = [v.copy() for v in s.Sigma_init_array]


def _controls(control_sequence_array):  # Vector{Vector{Vector}} -> (m, N, C)
    return np.stack([_stack(seq) for seq in control_sequence_array], axis=-1)


def _noise_for(problem, s, rng, C, use_true_model=False):
    """Noise tensor (n, N, particles, C) in the reference's consumption order ii -> kk -> tt (pets.jl:137-152)."""
    fs = problem.f_stochastic
    n, N, Kp = fs.dynamics.n, s.N, s.num_trajectory_samples
    w = np.zeros((n, N, Kp, C))
    chol = np.linalg.cholesky(fs.W) if fs.noise_kind == 0 else None
    zero = np.zeros(n)
    for ii in range(C):
        for kk in range(Kp):
            for tt in range(N):
                if use_true_model and fs.true_mixture is not None:  # additive: f_stochastic(0-dynamics) draws the noise
                    w[:, tt, kk, ii] = fs(zero, np.zeros(fs.dynamics.m), rng, True) - fs.dynamics(zero, np.zeros(fs.dynamics.m))
                elif fs.noise_kind == 1:
                    w[:, tt, kk, ii] = fs.noise_scale * rng.random(n)
                else:
                    w[:, tt, kk, ii] = chol @ rng.standard_normal(n)
    return w


def compute_cost_serial(s, problem, x, control_sequence_array, rng, use_true_model=False, noise=None):
    """compute_cost_serial (pets.jl:128-157). `noise` injects the tensor; else drawn from rng on the
    host in the reference's order; rng=None uses on-device Philox."""
    assert len(control_sequence_array) == s.num_control_samples
    assert all(len(seq) == s.N for seq in control_sequence_array)
    C = s.num_control_samples
    if noise is None and rng is not None and not isinstance(rng, int):
        noise = _noise_for(problem, s, rng, C, use_true_model)
    seed = rng if isinstance(rng, int) else 0
    return s._be().pets_costs(problem.spec(), np.asarray(x, float), _controls(control_sequence_array),
                              s.num_trajectory_samples, noise=noise, seed=seed,
                              gen=problem.f_stochastic.gen(use_true_model))


def compute_cost(s, problem, x, control_sequence_array, rng, use_true_model=False, noise=None):
    """compute_cost (pets.jl:100-126). One device, one stream: identical to the serial variant
    (the reference's per-process randjump streams do not exist here)."""
    return compute_cost_serial(s, problem, x, control_sequence_array, rng, use_true_model, noise)


def get_elite_samples(s, control_sequence_array, cost_array):  # pets.jl:159-171
    assert len(cost_array) == s.num_control_samples == len(control_sequence_array)
    _, _, idx = s._be().pets_refit(_controls(control_sequence_array), cost_array, s.num_elite, 0.0,
                                   _stack(s.mu_array), _stack(s.Sigma_array))
    return [control_sequence_array[i] for i in idx]


def compute_new_distribution(s, control_sequence_elite_array):  # pets.jl:173-191
    assert len(control_sequence_elite_array) == s.num_elite
    ctrl = _controls(control_sequence_elite_array)
    mu, Sg, _ = s._be().pets_refit(ctrl, np.arange(s.num_elite, dtype=float), s.num_elite, s.smoothing_factor,
                                   _stack(s.mu_array), _stack(s.Sigma_array))
    return _cols(mu), _mats(Sg)


def step_(s, problem, x, rng, use_true_model=False, verbose=False, serial=False, z=None, noise=None):  # :193-245
    s.iter_current += 1
    m, N, C = s.mu_array[0].size, s.N, s.num_control_samples
    seqs = []
    for ii in range(C):  # sampling order ii outer, tt inner (:208-216)
        seq = []
        for tt in range(N):
            zz = z[:, tt, ii] if z is not None else rng.standard_normal(m)
            seq.append(s.mu_array[tt] + np.linalg.cholesky(s.Sigma_array[tt]) @ zz)
        seqs.append(seq)
    costs = compute_cost_serial(s, problem, x, seqs, rng, use_true_model, noise=noise)
    elite = get_elite_samples(s, seqs, costs)
    s.mu_array, s.Sigma_array = compute_new_distribution(s, elite)


def solve_(s, problem, x_0, rng, use_true_model=False, verbose=False, serial=True, z_inject=None, noise=None):
    """solve! (pets.jl:270-281) -> (μ_array, Σ_array); the whole CEM loop runs on the device.
    rng: int seed -> on-device Philox;  numpy Generator -> z and noise are drawn on the host in the reference's
    consumption order (pets.jl:208-216 then :137-152, per iteration) and injected;
    z_inject (m,N,C,iter_max) + noise (n,N,particles,C,iter_max) -> exact replay."""
    initialize_(s)
    seed = 0
    if isinstance(rng, (int, np.integer)):
        seed = int(rng)
    elif rng is not None and z_inject is None:
        m, N, C, Kp = s.mu_array[0].size, s.N, s.num_control_samples, s.num_trajectory_samples
        n = problem.f_stochastic.dynamics.n
        z_inject = np.zeros((m, N, C, s.iter_max))
        noise = np.zeros((n, N, Kp, C, s.iter_max))
        for it in range(s.iter_max):
            for ii in range(C):
                for tt in range(N):
                    z_inject[:, tt, ii, it] = rng.standard_normal(m)
            noise[..., it] = _noise_for(problem, s, rng, C, use_true_model)
    mu, Sg = s._be().pets_solve(problem.spec(), np.asarray(x_0, float), _stack(s.mu_array), _stack(s.Sigma_array),
                                s.num_control_samples, s.num_trajectory_samples, s.num_elite, s.iter_max,
                                s.smoothing_factor, z_inject=z_inject, noise=noise, seed=seed,
                                gen=problem.f_stochastic.gen(use_true_model))
    s.mu_array, s.Sigma_array = _cols(mu), _mats(Sg)
    s.iter_current = s.iter_max
    return [v.copy() for v in s.mu_array], [v.copy() for v in s.Sigma_array]
