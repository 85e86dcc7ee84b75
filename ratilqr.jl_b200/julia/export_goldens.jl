# export_goldens.jl -- run the UNMODIFIED reference (RATiLQR.jl, Julia >= 1.5) on the shipped test problem and dump
# true-reference golden vectors as JSON.  The build image has no Julia, so this script has never been executed here;
# it uses only the reference's public API.  Usage (on a machine with Julia and the reference checked out):
#     julia --project=/path/to/RATiLQR.jl export_goldens.jl  /path/to/ratilqr-b200/tests/golden/julia
# tests/test_golden.py::test_true_reference_goldens picks the files up when the directory exists.
using RATiLQR, LinearAlgebra, Random
import JSON   # ] add JSON

outdir = length(ARGS) >= 1 ? ARGS[1] : "goldens_julia"
mkpath(outdir)

# C1: test/ileqg_test.jl:151-155, test/cross_entropy_bilevel_optimization_test.jl:13-21
f(x, u) = x.^1.3 + u.^1.5
c(k, x, u) = sum(x.^2.5 + u.^2.5)
h(x) = 1.0
W(k) = Matrix(0.01I, 2, 2)
N = 10
problem = FiniteHorizonRiskSensitiveOptimalControlProblem(f, c, h, W, N)
x_0 = zeros(2)
u_array = [0.1 * ones(2) for _ in 1:N]

cases = []
for θ in [0.0, 0.1, 0.3, 0.43, 0.5, 30.7]
    solver = ILEQGSolver(problem)
    x_array, l_array, L_array, value, ϵ_history = solve!(solver, problem, x_0, u_array, θ=θ, verbose=false)
    push!(cases, Dict("theta" => θ, "value" => value, "iters" => solver.iter_current,
                      "x" => reduce(hcat, x_array), "l" => reduce(hcat, l_array),
                      "L" => cat(L_array..., dims=3), "eps_history" => [[e[1], e[2]] for e in ϵ_history]))
end
open(joinpath(outdir, "c1_power_law.json"), "w") do io
    JSON.print(io, Dict("problem" => "f=x.^1.3+u.^1.5, c=sum(x.^2.5+u.^2.5), h=1, W=0.01I, N=10, x0=0, u=0.1",
                        "julia" => string(VERSION), "cases" => cases))
end

# RAT iLQR++ on the same problem (test/nelder_mead_bilevel_optimization_test.jl:21-26)
nm = NelderMeadBilevelOptimizationSolver(iter_max=20, ϵ=1e-3, θ_high_init=10.0, θ_low_init=1e-8)
θ_opt, x_array, l_array, L_array, c_opt = solve!(nm, problem, x_0, u_array, kl_bound=1.0, verbose=false)
open(joinpath(outdir, "c1_nelder_mead.json"), "w") do io
    JSON.print(io, Dict("theta_opt" => θ_opt, "c_opt" => c_opt, "nm_iters" => nm.iter_current))
end
println("wrote goldens to ", outdir)
