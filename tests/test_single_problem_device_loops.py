"""ratilqr_ce_solve / ratilqr_nm_solve: the reference's single-problem solve! calls (cross_entropy...jl:364-415,
nelder_mead...jl:276-352) with the WHOLE bilevel loop on the device, against the host loops of the Python mirror (which are
transliterations of the Julia control flow and fan only compute_cost out to the GPU) on the same random stream."""
import numpy as np
import pytest

import ratilqr_b200 as R
from ratilqr_b200 import cross_entropy as CE
from ratilqr_b200 import nelder_mead as NM
from ratilqr_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _ua(u):
    return [u[:, k].copy() for k in range(u.shape[1])]


@pytest.mark.parametrize("num_samples,num_elite,iter_max", [(10, 3, 5), (1024, 100, 3), (300, 30, 2)])
def test_ce_solve_on_device_equals_host_loop(gpu_be, num_samples, num_elite, iter_max):
    """configs[1]'s population (1024 theta, CTA-per-problem draw / rank / refit kernels) and the reference defaults
    (10 theta, thread-per-problem kernels): same theta draws from the same generator => same CE trajectory"""
    prob, x0, u = wl.c2_problem()
    kw = dict(num_samples=num_samples, num_elite=num_elite, iter_max=iter_max, backend=gpu_be)
    a, b = R.CrossEntropyBilevelOptimizationSolver(**kw), R.CrossEntropyBilevelOptimizationSolver(**kw)
    ra, rb = np.random.default_rng(2024), np.random.default_rng(2024)
    host = CE.solve_(a, prob, x0, _ua(u), ra, kl_bound=0.1)
    dev = CE.solve_on_device_(b, prob, x0, _ua(u), rb, kl_bound=0.1)
    assert np.isclose(dev[0], host[0], rtol=1e-12) and np.isclose(dev[4], host[4], rtol=1e-12)      # theta_opt, cost
    assert np.isclose(dev[5], host[5], rtol=0) and np.isclose(dev[6], host[6], rtol=0)              # theta_min, theta_max
    assert np.isclose(b.mu, a.mu, rtol=1e-12) and np.isclose(b.sigma, a.sigma, rtol=1e-10)
    assert b.mu_init == a.mu_init and b.sigma_init == a.sigma_init
    assert np.allclose(np.stack(dev[1], -1), np.stack(host[1], -1), rtol=1e-9, atol=1e-12)
    assert np.allclose(np.stack(dev[3], -1), np.stack(host[3], -1), rtol=1e-9, atol=1e-12)
    assert ra.standard_normal() == rb.standard_normal()   # the generator was advanced by exactly the draws consumed


def test_ce_solve_on_device_infeasible_start_shrinks_like_host_loop(gpu_be):
    """mu_init far beyond the breakdown threshold: first-iteration redraws with mu_init, sigma_init *= lambda (:293-298)"""
    prob, x0, u = wl.c2_problem(N=20)
    kw = dict(num_samples=64, num_elite=8, iter_max=3, mu_init=400.0, sigma_init=100.0, backend=gpu_be)
    a, b = R.CrossEntropyBilevelOptimizationSolver(**kw), R.CrossEntropyBilevelOptimizationSolver(**kw)
    ra, rb = np.random.default_rng(7), np.random.default_rng(7)
    host = CE.solve_(a, prob, x0, _ua(u), ra, kl_bound=0.5)
    dev = CE.solve_on_device_(b, prob, x0, _ua(u), rb, kl_bound=0.5)
    assert a.mu_init < 400.0 and b.mu_init == a.mu_init and b.sigma_init == a.sigma_init
    assert np.isclose(dev[0], host[0], rtol=1e-12) and np.isclose(dev[4], host[4], rtol=1e-12)
    assert ra.standard_normal() == rb.standard_normal()


def test_nm_solve_on_device_equals_host_loop(gpu_be):
    prob, x0, u = wl.c1_problem()
    nm = R.NelderMeadBilevelOptimizationSolver(iter_max=20, eps=1e-3, theta_high_init=10.0, theta_low_init=1e-8, backend=gpu_be)
    host = NM.solve_(nm, prob, x0, _ua(u), kl_bound=1.0)
    dev = gpu_be.nm_solve(prob.spec(), x0, u, 1.0, eps=1e-3, iter_max=20, theta_high_init=10.0, theta_low_init=1e-8)
    assert dev["status"] == 0 and dev["nm_iters"] == nm.iter_current and dev["n_evals"] == nm.n_evals == 12
    assert np.isclose(dev["theta_opt"], host[0], rtol=1e-12) and np.isclose(dev["value"], host[4], rtol=1e-12)
    assert np.isclose(dev["state"]["c_low"][0], nm.c_low, rtol=1e-12) and dev["state"]["theta_high_init"][0] == nm.theta_high_init
    assert np.allclose(dev["x"], np.stack(host[1], -1), rtol=1e-9, atol=1e-12)
