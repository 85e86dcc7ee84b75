"""Problem structs of the reference API (optimal_control_problems.jl), device-dispatchable.

Same constructor signatures as the reference:
    FiniteHorizonRiskSensitiveOptimalControlProblem(f, c, h, W, N)     (:67-73)
    FiniteHorizonGenerativeOptimalControlProblem(f_stochastic, c, h, N) (:126-131)
The fields must be registered device callables (models.py) for the GPU path; there is no CPU
fallback inside the solvers -- an unregistered closure raises.
"""
import numpy as np

from ._capi import Spec
from .models import ConstantCovariance, DeviceCost, DeviceDynamics, DeviceStochasticDynamics, _StageCost, _TerminalCost


class OptimalControlProblem:  # optimal_control_problems.jl:12
    pass


def _cost_of(c, h):
    if not isinstance(c, _StageCost) or not isinstance(h, _TerminalCost) or c.cost is not h.cost:
        raise TypeError("c and h must be the `.c` and `.h` of one registered DeviceCost "
                        "(arbitrary closures are not accelerated; see INTEGRATION.md)")
    return c.cost


class FiniteHorizonRiskSensitiveOptimalControlProblem(OptimalControlProblem):
    def __init__(self, f, c, h, W, N):
        self.f, self.c, self.h, self.W, self.N = f, c, h, W, int(N)

    def spec(self, cost_params=None):
        """ratilqr_problem_desc contents. `cost_params` (P, ncp) overrides with per-problem blocks."""
        if not isinstance(self.f, DeviceDynamics):
            raise TypeError("problem.f is not a registered DeviceDynamics")
        cost = _cost_of(self.c, self.h)
        if isinstance(cost, DeviceCost) and hasattr(cost, "n"):
            assert cost.n == self.f.n and cost.m == self.f.m
        if isinstance(self.W, ConstantCovariance):
            W = self.W.W
        elif callable(self.W):
            Ws = np.stack([np.asarray(self.W(k), float) for k in range(self.N)])
            W = Ws[0] if all(np.array_equal(Ws[0], w) for w in Ws) else Ws
        else:
            W = np.asarray(self.W, float)
        cp = cost.params() if cost_params is None else cost_params
        return Spec(self.f.model_id, cost.cost_id, self.f.n, self.f.m, self.N, self.f.params, cp, W)


class FiniteHorizonGenerativeOptimalControlProblem(OptimalControlProblem):
    def __init__(self, f_stochastic, c, h, N):
        self.f_stochastic, self.c, self.h, self.N = f_stochastic, c, h, int(N)

    def spec(self):
        fs = self.f_stochastic
        if not isinstance(fs, DeviceStochasticDynamics):
            raise TypeError("problem.f_stochastic is not a registered DeviceStochasticDynamics")
        cost = _cost_of(self.c, self.h)
        return Spec(fs.dynamics.model_id, cost.cost_id, fs.dynamics.n, fs.dynamics.m, self.N,
                    fs.dynamics.params, cost.params(), fs.W)
