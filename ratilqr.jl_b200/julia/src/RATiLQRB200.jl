# RATiLQRB200.jl -- Julia-side shim: dispatches the hot path of RATiLQR.jl to libratilqr_b200.so via ccall.
#
# STATUS: written blind -- there is no Julia toolchain in the build image, so this file has never been
# executed.  It mirrors, call for call, the ctypes binding that IS tested (ratilqr.jl_b200/_capi.py):
# same structs (include/ratilqr.h), same array layouts (column-major, instance slowest -- i.e. plain Julia
# Arrays), same status -> exception mapping.  Usage:
#     using RATiLQR; include("RATiLQRB200.jl"); using .RATiLQRB200
#     f = DeviceDynamics(:unicycle, [0.1]); cost = QuadraticCost(4, 2; Q=..., R=..., Qf=..., xg=...)
#     problem = FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, ConstantCovariance(W), N)
#     solve!(ILEQGSolver(problem), problem, x0, u_array, θ=0.5)          # runs on the B200
# Problems whose fields are ordinary closures keep using the original Julia methods.
module RATiLQRB200

using RATiLQR
import RATiLQR: solve!, compute_cost, compute_cost_serial, compute_cost_worker
using LinearAlgebra

const LIB = get(ENV, "RATILQR_B200_LIB", joinpath(@__DIR__, "..", "..", "csrc", "libratilqr_b200.so"))

export DeviceDynamics, QuadraticCost, PowerLawCost, ConstantCovariance, b200_context,
       UserDeviceDynamics, user_cost, register_user_model!

# ---- registered callables (subtypes of Function so the reference structs accept them,
#      optimal_control_problems.jl:68-71) -----------------------------------------------------------
const MODEL_IDS = Dict(:single_integrator => 1, :power_law => 2, :double_integrator => 3, :pendulum => 4,
                       :cartpole => 5, :unicycle => 6, :quadrotor => 7)
const MODEL_DIMS = Dict(1 => (2, 2), 2 => (2, 2), 3 => (4, 2), 4 => (2, 1), 5 => (4, 1), 6 => (4, 2), 7 => (12, 4))

struct DeviceDynamics <: Function
    model_id::Int32
    params::Vector{Float64}
end
DeviceDynamics(name::Symbol, params) = DeviceDynamics(Int32(MODEL_IDS[name]), Float64.(params))

# CPU evaluation (so the object is still a valid `f` for the original Julia code path); only the models the
# reference itself ships plus the unicycle are spelled out here.
function (f::DeviceDynamics)(x, u, f_returns_jacobian=false)
    p = f.params
    if f.model_id == 1
        return x + p[1] .* u
    elseif f.model_id == 2
        return x .^ p[1] + u .^ p[2]
    elseif f.model_id == 6
        dt = p[1]
        return [x[1] + dt * (x[4] * cos(x[3])), x[2] + dt * (x[4] * sin(x[3])), x[3] + dt * u[2], x[4] + dt * u[1]]
    end
    error("CPU evaluation of model $(f.model_id) is not spelled out in the shim")
end

struct QuadraticCost
    n::Int; m::Int
    params::Vector{Float64}   # [ws0, ws1, c0, c1, h0, xg, Q, R, Pc, Qf] (include/ratilqr.h)
    c::Function; h::Function
end
struct StageCost <: Function; id::Int32; params::Vector{Float64}; n::Int; m::Int; end
struct TerminalCost <: Function; id::Int32; params::Vector{Float64}; n::Int; m::Int; end
function QuadraticCost(n, m; Q=zeros(n, n), R=zeros(m, m), Qf=zeros(n, n), xg=zeros(n), Pc=zeros(n, m),
                       ws0=1.0, ws1=0.0, c0=0.0, c1=0.0, h0=0.0)
    p = vcat([ws0, ws1, c0, c1, h0], xg, vec(Q), vec(R), vec(Pc), vec(Qf))
    QuadraticCost(n, m, p, StageCost(1, p, n, m), TerminalCost(1, p, n, m))
end
# c = sum(x.^p + u.^p), h = h0 (the reference's shipped test problem, test/ileqg_test.jl:152-153); params [p, h0]
struct PowerLawCost
    params::Vector{Float64}
    c::Function; h::Function
end
PowerLawCost(p, h0; n=2, m=2) = (q = Float64[p, h0]; PowerLawCost(q, StageCost(2, q, n, m), TerminalCost(2, q, n, m)))
function (c::StageCost)(k, x, u)
    c.id == 2 && return sum(x .^ c.params[1]) + sum(u .^ c.params[1])
    n, m, p = c.n, c.m, c.params
    xg = p[6:5+n]; Q = reshape(p[6+n:5+n+n^2], n, n); o = 5 + n + n^2
    R = reshape(p[o+1:o+m^2], m, m); Pc = reshape(p[o+m^2+1:o+m^2+n*m], n, m)
    dx = x - xg
    (p[1] + p[2] * k) * (0.5 * dx' * Q * dx + 0.5 * u' * R * u + dx' * Pc * u) + p[3] + p[4] * k
end
function (h::TerminalCost)(x)
    h.id == 2 && return h.params[2]
    n, m, p = h.n, h.m, h.params
    o = 5 + n + n^2 + m^2 + n * m
    Qf = reshape(p[o+1:o+n^2], n, n); dx = x - p[6:5+n]
    0.5 * dx' * Qf * dx + p[5]
end
struct ConstantCovariance <: Function; W::Matrix{Float64}; end
(w::ConstantCovariance)(k) = w.W

# ---- C structs (include/ratilqr.h) ------------------------------------------------------------------
struct ProblemDesc
    model_id::Int32; cost_id::Int32; n::Int32; m::Int32; N::Int32
    model_params::Ptr{Float64}; n_model_params::Int32
    cost_params::Ptr{Float64}; n_cost_params::Int32; cost_params_count::Int32
    W::Ptr{Float64}; W_time_varying::Int32
end
struct IleqgOpts
    mu_min::Float64; delta_0::Float64; lambda::Float64; d::Float64; iter_max::Int32; adaptive_eps_init::Int32
    eps_init::Float64; eps_min::Float64; f_returns_jacobian::Int32
end
struct BatchIn
    P::Int32; K::Int32; x0::Ptr{Float64}; x0_count::Int32; u_init::Ptr{Float64}; u_count::Int32; theta::Ptr{Float64}
end
struct IleqgOut
    x::Ptr{Float64}; l::Ptr{Float64}; L::Ptr{Float64}; value::Ptr{Float64}
    status::Ptr{Int32}; iters::Ptr{Int32}; trials::Ptr{Int32}; restarts::Ptr{Int32}
    mu::Ptr{Float64}; d_current::Ptr{Float64}; eps_hist::Ptr{Float64}; eps_hist_cap::Int32
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function b200_context(device::Integer=0)
    if CTX[] == C_NULL
        rc = ccall((:ratilqr_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), CTX, device)
        rc == 0 || error("ratilqr_create failed ($rc): no CUDA device / library (there is no CPU fallback)")
    end
    CTX[]
end
last_error() = unsafe_string(ccall((:ratilqr_last_error, LIB), Cstring, (Ptr{Cvoid},), CTX[]))

# ---- user-extensible device models (include/ratilqr.h: ratilqr_user_model_register) ------------------------------
# The user keeps the Julia closure for the CPU path and adds a CUDA C++ snippet for the device:
#   f = UserDeviceDynamics(4, 2, [0.1], src_dynamics, (x, u) -> ...)     # `template <class T> void dynamics(p, x, u, xn)`
#   c, h = user_cost(4, 2, cp, src_cost, (k, x, u) -> ..., x -> ...)       # stage_cost<T> / terminal_cost<T>
#   register_user_model!(f, c)      # NVRTC compile + load; afterwards solve!/compute_cost dispatch to the GPU
mutable struct UserDeviceDynamics <: Function
    n::Int; m::Int; params::Vector{Float64}; src::String; cpu::Function; model_id::Int32
end
UserDeviceDynamics(n, m, params, src, cpu) = UserDeviceDynamics(n, m, Float64.(params), src, cpu, Int32(0))
(f::UserDeviceDynamics)(x, u, f_returns_jacobian=false) = f.cpu(x, u)
struct UserStageCost <: Function; id::Int32; params::Vector{Float64}; n::Int; m::Int; src::String; cpu::Function; end
struct UserTerminalCost <: Function; id::Int32; params::Vector{Float64}; n::Int; m::Int; cpu::Function; end
(c::UserStageCost)(k, x, u) = c.cpu(k, x, u)
(h::UserTerminalCost)(x) = h.cpu(x)
user_cost(n, m, params, src, c_cpu, h_cpu) = (UserStageCost(Int32(100), Float64.(params), n, m, src, c_cpu),
                                              UserTerminalCost(Int32(100), Float64.(params), n, m, h_cpu))
struct UserModelDesc
    n::Int32; m::Int32; dynamics_src::Cstring; base_model_id::Int32; n_model_params::Int32
    cost_src::Cstring; base_cost_id::Int32; n_cost_params::Int32
    a_kind::Ptr{Int8}; b_kind::Ptr{Int8}; q_kind::Ptr{Int8}; r_kind::Ptr{Int8}; p_kind::Ptr{Int8}   # optional structure, C_NULL = dense
end
dims_of(f::DeviceDynamics) = MODEL_DIMS[f.model_id]
dims_of(f::UserDeviceDynamics) = (f.n, f.m)

# f: UserDeviceDynamics or DeviceDynamics; c: UserStageCost or StageCost.  Returns the dynamics object to put into the
# problem struct (its model_id names the compiled pair inside this process's context).
function register_user_model!(f, c)
    ud, uc = f isa UserDeviceDynamics, c isa UserStageCost
    (ud || uc) || error("neither the dynamics nor the cost is user-supplied")
    n, m = dims_of(f)
    dsrc = ud ? f.src : ""; csrc = uc ? c.src : ""
    log = zeros(UInt8, 1 << 16); id = Ref{Int32}(0)
    GC.@preserve dsrc csrc log begin
        um = UserModelDesc(n, m, ud ? Base.unsafe_convert(Cstring, dsrc) : Cstring(C_NULL), ud ? 0 : f.model_id,
                           length(f.params), uc ? Base.unsafe_convert(Cstring, csrc) : Cstring(C_NULL), uc ? 0 : c.id,
                           uc ? length(c.params) : 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)
        rc = ccall((:ratilqr_user_model_register, LIB), Int32, (Ptr{Cvoid}, Ref{UserModelDesc}, Ref{Int32}, Ptr{UInt8}, Int64),
                   b200_context(), um, id, log, length(log))
        rc == 0 || error("ratilqr_user_model_register failed ($rc):\n" * unsafe_string(pointer(log)))
    end
    if ud
        f.model_id = id[]
        return f
    end
    return DeviceDynamics(id[], f.params)   # registered dynamics bound to the user cost (CPU evaluation: use `f` itself)
end

is_device(problem) = (problem.f isa DeviceDynamics || (problem.f isa UserDeviceDynamics && problem.f.model_id != 0)) &&
                     (problem.c isa StageCost || problem.c isa UserStageCost) &&
                     (problem.h isa TerminalCost || problem.h isa UserTerminalCost)

opts_of(s::ILEQGSolver) = IleqgOpts(s.μ_min, s.Δ_0, s.λ, s.d, s.iter_max, s.ϵ_init_auto, s.ϵ_init_init, s.ϵ_min, s.f_returns_jacobian)

function status_error(st)
    st == 1 || st == 2 ? AssertionError("M = inv(W) - θ*S is not PSD") :
    st == 3 ? DomainError(-1.0, "negative base of a real power in the model") :
    ErrorException("iLEQG status $st")
end

# batched core: θ vector in, (value, status, x, l, L) out
function solve_batch(problem, opts::IleqgOpts, x_0::Vector{Float64}, u_array::Vector{Vector{Float64}},
                     θs::Vector{Float64}; want_traj::Bool=true, eps_cap::Int=0)
    n, m = dims_of(problem.f); N = problem.N; B = length(θs)
    U = reduce(hcat, u_array)                        # m x N, column-major
    Ws = [Matrix{Float64}(problem.W(k)) for k in 0:N-1]
    tv = any(w != Ws[1] for w in Ws)
    Wbuf = tv ? reduce(vcat, vec.(Ws)) : vec(Ws[1])
    x = want_traj ? zeros(n, N + 1, B) : zeros(0); l = want_traj ? zeros(m, N, B) : zeros(0)
    L = want_traj ? zeros(m, n, N, B) : zeros(0)
    value = zeros(B); status = zeros(Int32, B); iters = zeros(Int32, B); trials = zeros(Int32, B)
    restarts = zeros(Int32, B); mu = zeros(B); dcur = zeros(B); eh = zeros(2, max(eps_cap, 1), B)
    mp = problem.f.params; cp = problem.c.params
    GC.@preserve mp cp Wbuf U x_0 θs x l L value status iters trials restarts mu dcur eh begin
        desc = ProblemDesc(problem.f.model_id, problem.c.id, n, m, N, pointer(mp), length(mp), pointer(cp), length(cp), 1,
                           pointer(Wbuf), tv)
        bin = BatchIn(1, B, pointer(x_0), 1, pointer(U), 1, pointer(θs))
        out = IleqgOut(want_traj ? pointer(x) : C_NULL, want_traj ? pointer(l) : C_NULL, want_traj ? pointer(L) : C_NULL,
                       pointer(value), pointer(status), pointer(iters), pointer(trials), pointer(restarts), pointer(mu),
                       pointer(dcur), eps_cap > 0 ? pointer(eh) : C_NULL, eps_cap)
        rc = ccall((:ratilqr_ileqg_solve_batch, LIB), Int32,
                   (Ptr{Cvoid}, Ref{ProblemDesc}, Ref{IleqgOpts}, Ref{BatchIn}, Ref{IleqgOut}),
                   b200_context(), desc, opts, bin, out)
        rc == 0 || error("ratilqr_ileqg_solve_batch failed ($rc): $(last_error())")
    end
    (value=value, status=status, iters=iters, trials=trials, mu=mu, d_current=dcur, x=x, l=l, L=L, eps_hist=eh)
end

# ---- solve!(::ILEQGSolver, ...) ileqg.jl:635-659 -------------------------------------------------------
function solve!(ileqg::ILEQGSolver, problem::FiniteHorizonRiskSensitiveOptimalControlProblem,
                x_0::Vector{Float64}, u_array::Vector{Vector{Float64}}; θ::Float64, verbose=true)
    is_device(problem) || return invoke(solve!, Tuple{ILEQGSolver, FiniteHorizonRiskSensitiveOptimalControlProblem,
                                                      Vector{Float64}, Vector{Vector{Float64}}},
                                        ileqg, problem, x_0, u_array; θ=θ, verbose=verbose)
    cap = 4 * ileqg.iter_max
    r = solve_batch(problem, opts_of(ileqg), x_0, u_array, [θ]; eps_cap=cap)
    r.status[1] == 0 || throw(status_error(r.status[1]))
    N = problem.N
    ileqg.x_array = [r.x[:, k, 1] for k in 1:N+1]; ileqg.l_array = [r.l[:, k, 1] for k in 1:N]
    ileqg.L_array = [r.L[:, :, k, 1] for k in 1:N]
    ileqg.value_current = r.value[1]; ileqg.iter_current = r.iters[1]; ileqg.d_current = r.d_current[1]; ileqg.μ = r.mu[1]
    ileqg.ϵ_history = [(r.eps_hist[1, i, 1], r.eps_hist[2, i, 1]) for i in 1:min(r.trials[1], cap)]
    return copy(ileqg.x_array), copy(ileqg.l_array), copy(ileqg.L_array), ileqg.value_current, copy(ileqg.ϵ_history)
end

# ---- compute_cost (cross_entropy_bilevel_optimization.jl:173-195): the θ fan-out in ONE launch ------------
function ce_opts(s)
    IleqgOpts(s.μ_min_ileqg, s.Δ_0_ileqg, s.λ_ileqg, s.d_ileqg, s.iter_max_ileqg, s.ϵ_init_auto_ileqg,
              s.ϵ_init_ileqg, s.ϵ_min_ileqg, s.f_returns_jacobian)
end
function device_costs(solver, problem, x, u_array, θ_array, kl_bound)
    r = solve_batch(problem, ce_opts(solver), x, u_array, Float64.(θ_array); want_traj=false)
    [r.status[i] == 0 ? r.value[i] + kl_bound / θ_array[i] : Inf for i in eachindex(θ_array)]
end
function compute_cost(ce_solver::CrossEntropyBilevelOptimizationSolver, problem::FiniteHorizonRiskSensitiveOptimalControlProblem,
                      x::Vector{Float64}, u_array::Vector{Vector{Float64}}, θ_array::Vector{Float64}, kl_bound::Float64)
    is_device(problem) ? device_costs(ce_solver, problem, x, u_array, θ_array, kl_bound) :
        invoke(compute_cost, Tuple{CrossEntropyBilevelOptimizationSolver, FiniteHorizonRiskSensitiveOptimalControlProblem,
                                   Vector{Float64}, Vector{Vector{Float64}}, Vector{Float64}, Float64},
               ce_solver, problem, x, u_array, θ_array, kl_bound)
end
function compute_cost_serial(ce_solver::CrossEntropyBilevelOptimizationSolver, problem::FiniteHorizonRiskSensitiveOptimalControlProblem,
                             x::Vector{Float64}, u_array::Vector{Vector{Float64}}, θ_array::Vector{Float64}, kl_bound::Float64)
    @assert length(θ_array) == ce_solver.num_samples
    compute_cost(ce_solver, problem, x, u_array, θ_array, kl_bound)
end
# RAT iLQR++: compute_cost_worker (nelder_mead_bilevel_optimization.jl:134-158)
function compute_cost_worker(nm_solver::NelderMeadBilevelOptimizationSolver, problem::FiniteHorizonRiskSensitiveOptimalControlProblem,
                             x::Vector{Float64}, u_array::Vector{Vector{Float64}}, θ::Float64, kl_bound::Float64)
    is_device(problem) ? device_costs(nm_solver, problem, x, u_array, [θ], kl_bound)[1] :
        invoke(compute_cost_worker, Tuple{NelderMeadBilevelOptimizationSolver, FiniteHorizonRiskSensitiveOptimalControlProblem,
                                          Vector{Float64}, Vector{Vector{Float64}}, Float64, Float64},
               nm_solver, problem, x, u_array, θ, kl_bound)
end

end # module
