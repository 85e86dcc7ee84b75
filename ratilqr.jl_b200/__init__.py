"""ratilqr_b200 -- B200-native hot path of StanfordMSL/RATiLQR.jl behind the reference's API.

Export list mirrors src/RATiLQR.jl:20-74 (Julia `f!` -> `f_`; per-solver functions live in their
modules because Python has no multiple dispatch: `ileqg.solve_`, `cross_entropy.solve_`, ...).
"""
from . import cross_entropy, ileqg, models, mpc, nelder_mead, pets  # noqa: F401
from ._capi import ApiError, Spec, make_opts  # noqa: F401
from ._lib import default_backend, load_library, new_backend, new_multi, set_default_backend  # noqa: F401
from .cross_entropy import CrossEntropyBilevelOptimizationSolver, InjectedNormals  # noqa: F401
from .ileqg import (ILEQGSolver, NotPositiveDefinite, approximate_model, decrease_mu_and_delta_,  # noqa: F401
                    increase_mu_and_delta_, integrate_cost, line_search_, simulate_dynamics,
                    solve_approximate_dp, solve_approximate_dp_)
from .models import (CartPole, ConstantCovariance, DeviceStochasticDynamics, DomainError, DoubleIntegrator,  # noqa: F401
                     L1ControlCost, Pendulum, PowerLawCost, PowerLawDynamics, QuadraticCost, Quadrotor,
                     SingleIntegrator, Unicycle, UserCost, UserDynamics, register_user_model)
from .nelder_mead import NelderMeadBilevelOptimizationSolver  # noqa: F401
from .pets import CrossEntropyDirectOptimizationSolver, PETSSolver  # noqa: F401
from .problems import (FiniteHorizonGenerativeOptimalControlProblem,  # noqa: F401
                       FiniteHorizonRiskSensitiveOptimalControlProblem, OptimalControlProblem)

# north-star spellings
CrossEntropyBilevelOptSolver = CrossEntropyBilevelOptimizationSolver
NelderMeadBilevelOptSolver = NelderMeadBilevelOptimizationSolver
