# Smoke test of the ccall shim on a machine that has Julia, RATiLQR.jl and a B200 (written blind, see src/RATiLQRB200.jl).
# The device path must agree with the reference's own CPU path on the reference's shipped test problem
# (test/ileqg_test.jl:151-174 of RATiLQR.jl) to 1e-9.
using Test, LinearAlgebra, RATiLQR, RATiLQRB200

@testset "iLEQG on the device vs the reference" begin
    a, b, p, N = 1.3, 1.5, 2.5, 10
    f_cpu(x, u) = x .^ a + u .^ b
    c_cpu(k, x, u) = sum(x .^ p + u .^ p)
    h_cpu(x) = 1.0
    W(k) = Matrix(0.01I, 2, 2)
    x_0, u_array = zeros(2), [0.1 * ones(2) for _ in 1:N]
    ref = FiniteHorizonRiskSensitiveOptimalControlProblem(f_cpu, c_cpu, h_cpu, W, N)
    x_r, l_r, L_r, v_r, _ = solve!(ILEQGSolver(ref), ref, x_0, u_array, θ=0.3, verbose=false)

    f = DeviceDynamics(:power_law, [a, b])
    cost = PowerLawCost(p, 1.0)
    dev = FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, ConstantCovariance(Matrix(0.01I, 2, 2)), N)
    x_d, l_d, L_d, v_d, _ = solve!(ILEQGSolver(dev), dev, x_0, u_array, θ=0.3, verbose=false)
    @test isapprox(v_d, v_r; rtol=1e-9)
    @test all(isapprox.(x_d, x_r; rtol=1e-9, atol=1e-12))
    @test all(isapprox.(L_d, L_r; rtol=1e-9, atol=1e-12))
end
