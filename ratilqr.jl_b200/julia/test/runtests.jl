# Smoke test of the ccall binding on a machine that has Julia, RATiLQR.jl and a B200 (written blind, see
# src/RATiLQRB200.jl).  The device path must agree with the reference's own CPU path on the reference's shipped test
# problem (test/ileqg_test.jl:151-174, test/cross_entropy_bilevel_optimization_test.jl:24-41,
# test/nelder_mead_bilevel_optimization_test.jl:21-32 of RATiLQR.jl) to 1e-9.
using Test, LinearAlgebra, Random, RATiLQR, RATiLQRB200

a, b, p, N = 1.3, 1.5, 2.5, 10
f_cpu(x, u) = x .^ a + u .^ b
c_cpu(k, x, u) = sum(x .^ p + u .^ p)
h_cpu(x) = 1.0
W(k) = Matrix(0.01I, 2, 2)
x_0, u_array = zeros(2), [0.1 * ones(2) for _ in 1:N]
ref = FiniteHorizonRiskSensitiveOptimalControlProblem(f_cpu, c_cpu, h_cpu, W, N)
cost = PowerLawCost(p, 1.0)
plain = FiniteHorizonRiskSensitiveOptimalControlProblem(DeviceDynamics(:power_law, [a, b]), cost.c, cost.h,
                                                        ConstantCovariance(Matrix(0.01I, 2, 2)), N)
dev = device(plain)   # DeviceProblem: the reference's generic functions have GPU methods for it

@testset "iLEQG on the device vs the reference" begin
    x_r, l_r, L_r, v_r, e_r = solve!(ILEQGSolver(ref), ref, x_0, u_array, θ=0.3, verbose=false)
    x_d, l_d, L_d, v_d, e_d = solve!(ILEQGSolver(dev), dev, x_0, u_array, θ=0.3, verbose=false)
    @test isapprox(v_d, v_r; rtol=1e-9)
    @test all(isapprox.(x_d, x_r; rtol=1e-9, atol=1e-12))
    @test all(isapprox.(L_d, L_r; rtol=1e-9, atol=1e-12))
    @test length(e_d) == length(e_r) && all(first.(e_d) .== first.(e_r))
    # the device callables are still valid CPU closures: the untouched reference path on the same object
    x_p, _, _, v_p, _ = solve!(ILEQGSolver(plain), plain, x_0, u_array, θ=0.3, verbose=false)
    @test isapprox(v_p, v_r; rtol=1e-12)
    @test_throws AssertionError solve!(ILEQGSolver(dev), dev, x_0, u_array, θ=100.0, verbose=false)   # neurotic breakdown
end

@testset "RAT iLQR (CE) and RAT iLQR++ (NM) on the device vs the reference" begin
    ce_r = CrossEntropyBilevelOptimizationSolver(num_samples=3); ce_d = CrossEntropyBilevelOptimizationSolver(num_samples=3)
    θs = [0.1, 0.3, 0.43]
    @test isapprox(compute_cost(ce_d, dev, x_0, u_array, θs, 1.0), compute_cost_serial(ce_r, ref, x_0, u_array, θs, 1.0); rtol=1e-9)
    r_r = solve!(ce_r, ref, x_0, u_array, MersenneTwister(12344), kl_bound=1.0, verbose=false, serial=true)
    rng = MersenneTwister(12344)
    r_d = solve!(ce_d, dev, x_0, u_array, rng, kl_bound=1.0, verbose=false)
    @test isapprox(r_d[1], r_r[1]; rtol=1e-9) && isapprox(r_d[5], r_r[5]; rtol=1e-9)   # same θ draws => same θ_opt, cost
    @test isapprox(ce_d.μ_init, ce_r.μ_init) && isapprox(ce_d.σ_init, ce_r.σ_init)
    nm_r = NelderMeadBilevelOptimizationSolver(iter_max=20, ϵ=1e-3, θ_high_init=10.0, θ_low_init=1e-8)
    nm_d = NelderMeadBilevelOptimizationSolver(iter_max=20, ϵ=1e-3, θ_high_init=10.0, θ_low_init=1e-8)
    n_r = solve!(nm_r, ref, x_0, u_array, kl_bound=1.0, verbose=false)
    n_d = solve!(nm_d, dev, x_0, u_array, kl_bound=1.0, verbose=false)
    @test isapprox(n_d[1], n_r[1]; rtol=1e-12) && isapprox(n_d[5], n_r[5]; rtol=1e-9)
    @test nm_d.iter_current == nm_r.iter_current
end

@testset "run-time hooks: plain reference problems with device callables" begin
    if isdefined(Base, :invoke_in_world)
        install_hooks!()
        _, _, _, v_h, _ = solve!(ILEQGSolver(plain), plain, x_0, u_array, θ=0.3, verbose=false)    # device path
        _, _, _, v_c, _ = solve!(ILEQGSolver(ref), ref, x_0, u_array, θ=0.3, verbose=false)        # forwarded to the original
        @test isapprox(v_h, v_c; rtol=1e-9)
    end
end
