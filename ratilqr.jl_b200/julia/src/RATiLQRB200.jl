# RATiLQRB200.jl -- Julia-side binding of libratilqr_b200.so (include/ratilqr.h) for RATiLQR.jl.
#
# STATUS: written blind -- there is no Julia toolchain in the build image, so this file has never been executed.
# What IS checked here: tests/test_julia_shim_layout.py parses the C-struct declarations and the ccall signatures
# below and compares them field by field (types, sizes, offsets, argument counts) with the ctypes binding that the
# whole test-suite drives (ratilqr.jl_b200/_capi.py) and with include/ratilqr.h.
#
# Design (no method of RATiLQR.jl is re-defined, so there is nothing to overwrite and nothing that can recurse):
#
#   * the reference's problem structs are not parametric (`f::Function`, optimal_control_problems.jl:68-71), so a
#     package cannot dispatch on "this problem holds device callables".  The binding therefore adds its own problem
#     types, `DeviceProblem` / `DeviceGenerativeProblem` (<: RATiLQR.OptimalControlProblem), thin wrappers around a
#     reference problem whose fields are registered device callables, and NEW methods of the reference's generic
#     functions for them -- same argument lists, same return tuples (src/RATiLQR.jl:20-74):
#         ILEQGSolver(problem; kw...)                                                  ileqg.jl:191-194
#         solve!(::ILEQGSolver, problem, x_0, u_array; θ, verbose)                     ileqg.jl:635-659
#         compute_cost / compute_cost_serial(::CrossEntropyBilevelOptimizationSolver, ...)   cross_entropy...jl:173-227
#         solve!(::CrossEntropyBilevelOptimizationSolver, problem, x_0, u_array, rng; kl_bound, verbose, serial)  :364-415
#         compute_cost_worker(::NelderMeadBilevelOptimizationSolver, ...)              nelder_mead...jl:134-158
#         solve!(::NelderMeadBilevelOptimizationSolver, problem, x_0, u_array; kl_bound, verbose)   :276-352
#         solve!(::CrossEntropyDirectOptimizationSolver, problem, x_0, rng; use_true_model, verbose, serial)  pets.jl:270-281
#     Usage:  problem = device(FiniteHorizonRiskSensitiveOptimalControlProblem(f, cost.c, cost.h, ConstantCovariance(W), N))
#             solve!(ILEQGSolver(problem), problem, x_0, u_array, θ=0.5)           # runs on the B200
#   * `install_hooks!()` (optional, Julia >= 1.6) makes plain reference problems with device callables take the device
#     path too: it records the current world age, then re-defines the five reference entry points at run time (never
#     during precompilation); non-device problems are forwarded with `Base.invoke_in_world(old_world, ...)`, i.e. to
#     the ORIGINAL method as it existed before the hook -- not to the hook itself.
#   * INTEGRATION.md shows the alternative a maintainer of RATiLQR.jl would choose: five two-line guards inside the
#     reference's own methods.
module RATiLQRB200

using RATiLQR
using LinearAlgebra
using Random

const LIB = get(ENV, "RATILQR_B200_LIB", joinpath(@__DIR__, "..", "..", "csrc", "libratilqr_b200.so"))

export DeviceDynamics, DeviceStochasticDynamics, QuadraticCost, PowerLawCost, ConstantCovariance, b200_context,
       DeviceProblem, DeviceGenerativeProblem, device, is_device, install_hooks!,
       UserDeviceDynamics, user_cost, register_user_model!,
       solve_batch, ce_solve_fleet, nm_solve_fleet, mc_rollout

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

# f_stochastic(x, u, rng, use_true_model) of FiniteHorizonGenerativeOptimalControlProblem (optimal_control_problems.jl:82-87)
# for a registered model: x+ = f(x, u) + w, w ~ N(0, W) (noise_kind 0) or uniform[0,1)*scale (noise_kind 1, test/pets_test.jl:15)
struct DeviceStochasticDynamics <: Function
    f::DeviceDynamics
    W::Matrix{Float64}
    noise_kind::Int32
    noise_scale::Float64
    ensemble_params::Matrix{Float64}   # n_model_params x n_ensemble (0 columns = none)
end
DeviceStochasticDynamics(f, W; noise_kind=0, noise_scale=1.0, ensemble_params=zeros(length(f.params), 0)) =
    DeviceStochasticDynamics(f, Matrix{Float64}(W), Int32(noise_kind), noise_scale, ensemble_params)
function (fs::DeviceStochasticDynamics)(x, u, rng, use_true_model=false)
    n = length(x)
    w = fs.noise_kind == 1 ? fs.noise_scale .* rand(rng, n) : cholesky(Symmetric(fs.W)).L * randn(rng, n)
    fs.f(x, u) + w
end

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
struct CeOpts
    num_samples::Int32; num_elite::Int32; iter_max::Int32; lambda::Float64; use_theta_max::Int32
end
struct NmOpts
    alpha::Float64; beta::Float64; gamma::Float64; eps::Float64; lambda::Float64; iter_max::Int32
end
struct NoiseMixture
    n_components::Int32; weights::Ptr{Float64}; means::Ptr{Float64}; covs::Ptr{Float64}
end
struct GenerativeDesc
    noise_kind::Int32; noise_scale::Float64; n_ensemble::Int32; ensemble_params::Ptr{Float64}
    true_model::Ptr{NoiseMixture}; use_true_model::Int32
end
struct UserModelDesc
    n::Int32; m::Int32; dynamics_src::Cstring; base_model_id::Int32; n_model_params::Int32
    cost_src::Cstring; base_cost_id::Int32; n_cost_params::Int32
    a_kind::Ptr{Int8}; b_kind::Ptr{Int8}; q_kind::Ptr{Int8}; r_kind::Ptr{Int8}; p_kind::Ptr{Int8}   # optional structure, C_NULL = dense
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
#   register_user_model!(f, c)      # NVRTC compile + load; afterwards the device methods accept the pair
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

# ---- device problem types ------------------------------------------------------------------------------------------
const RSProblem = FiniteHorizonRiskSensitiveOptimalControlProblem
const GenProblem = FiniteHorizonGenerativeOptimalControlProblem

is_device(problem::RSProblem) = (problem.f isa DeviceDynamics || (problem.f isa UserDeviceDynamics && problem.f.model_id != 0)) &&
                                (problem.c isa StageCost || problem.c isa UserStageCost) &&
                                (problem.h isa TerminalCost || problem.h isa UserTerminalCost)
is_device(problem::GenProblem) = problem.f_stochastic isa DeviceStochasticDynamics &&
                                 (problem.c isa StageCost || problem.c isa UserStageCost) &&
                                 (problem.h isa TerminalCost || problem.h isa UserTerminalCost)

"""A reference problem whose `f`, `c`, `h` are registered device callables; every solver entry point of RATiLQR.jl has
a method for it that runs on the GPU.  Field access is forwarded, so `problem.N`, `problem.f(x, u)` ... keep working."""
struct DeviceProblem <: RATiLQR.OptimalControlProblem
    problem::RSProblem
    function DeviceProblem(p::RSProblem)
        is_device(p) || throw(ArgumentError("f, c, h must be registered device callables (DeviceDynamics, QuadraticCost(...).c/.h, ...): arbitrary closures cannot run on the GPU and there is no CPU fallback in the library"))
        new(p)
    end
end
struct DeviceGenerativeProblem <: RATiLQR.OptimalControlProblem
    problem::GenProblem
    function DeviceGenerativeProblem(p::GenProblem)
        is_device(p) || throw(ArgumentError("f_stochastic must be a DeviceStochasticDynamics and c, h registered costs"))
        new(p)
    end
end
device(p::RSProblem) = DeviceProblem(p)
device(p::GenProblem) = DeviceGenerativeProblem(p)
Base.getproperty(p::DeviceProblem, s::Symbol) = s === :problem ? getfield(p, :problem) : getproperty(getfield(p, :problem), s)
Base.getproperty(p::DeviceGenerativeProblem, s::Symbol) = s === :problem ? getfield(p, :problem) : getproperty(getfield(p, :problem), s)

RATiLQR.ILEQGSolver(p::DeviceProblem; kw...) = ILEQGSolver(p.problem; kw...)   # ileqg.jl:191-194 (sizes its arrays from problem.N)

opts_of(s::ILEQGSolver) = IleqgOpts(s.μ_min, s.Δ_0, s.λ, s.d, s.iter_max, s.ϵ_init_auto, s.ϵ_init_init, s.ϵ_min, s.f_returns_jacobian)
# the iLEQG options carried by the two bilevel solvers (cross_entropy...jl:71-80, nelder_mead...jl:72-81)
bilevel_opts(s) = IleqgOpts(s.μ_min_ileqg, s.Δ_0_ileqg, s.λ_ileqg, s.d_ileqg, s.iter_max_ileqg, s.ϵ_init_auto_ileqg,
                            s.ϵ_init_ileqg, s.ϵ_min_ileqg, s.f_returns_jacobian)

function status_error(st)
    st == 1 || st == 2 ? AssertionError("M = inv(W) - θ*S is not PSD") :
    st == 3 ? DomainError(-1.0, "negative base of a real power in the model") :
    ErrorException("iLEQG status $st")
end

# host buffers of a problem description; keep the returned arrays alive (GC.@preserve) while `desc` is in use
function pack_problem(problem::RSProblem)
    n, m = dims_of(problem.f); N = problem.N
    Ws = [Matrix{Float64}(problem.W(k)) for k in 0:N-1]
    tv = any(w != Ws[1] for w in Ws)
    Wbuf = tv ? reduce(vcat, vec.(Ws)) : vec(Ws[1])
    (n=n, m=m, N=N, mp=problem.f.params, cp=problem.c.params, Wbuf=Wbuf, tv=Int32(tv),
     model_id=Int32(problem.f.model_id), cost_id=Int32(problem.c.id))
end
desc_of(b, cost_params=b.cp, count=1) = ProblemDesc(b.model_id, b.cost_id, b.n, b.m, b.N, pointer(b.mp), length(b.mp),
                                                    pointer(cost_params), length(cost_params) ÷ count, count, pointer(b.Wbuf), b.tv)
unpack_traj(x, l, L, N, b=1) = ([x[:, k, b] for k in 1:N+1], [l[:, k, b] for k in 1:N], [L[:, :, k, b] for k in 1:N])

# ---- batched core: θ vector in, (value, status, x, l, L) out -- ratilqr_ileqg_solve_batch ---------------------------
function solve_batch(problem::RSProblem, opts::IleqgOpts, x_0::Vector{Float64}, u_array::Vector{Vector{Float64}},
                     θs::Vector{Float64}; want_traj::Bool=true, eps_cap::Int=0)
    b = pack_problem(problem); n, m, N = b.n, b.m, b.N; B = length(θs)
    U = reduce(hcat, u_array)                        # m x N, column-major
    x = want_traj ? zeros(n, N + 1, B) : zeros(0); l = want_traj ? zeros(m, N, B) : zeros(0)
    L = want_traj ? zeros(m, n, N, B) : zeros(0)
    value = zeros(B); status = zeros(Int32, B); iters = zeros(Int32, B); trials = zeros(Int32, B)
    restarts = zeros(Int32, B); mu = zeros(B); dcur = zeros(B); eh = zeros(2, max(eps_cap, 1), B)
    GC.@preserve b U x_0 θs x l L value status iters trials restarts mu dcur eh begin
        desc = desc_of(b)
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

# ---- solve!(::ILEQGSolver, ...) ileqg.jl:635-659 -------------------------------------------------------------------
function device_solve!(ileqg::ILEQGSolver, problem::RSProblem, x_0::Vector{Float64}, u_array::Vector{Vector{Float64}};
                       θ::Float64, verbose=true)
    per_iter = ceil(Int, log(ileqg.ϵ_min / ileqg.ϵ_init_init) / log(ileqg.λ)) + 2   # line-search trials of one iteration, at most
    cap = max(16, ileqg.iter_max * per_iter)
    r = solve_batch(problem, opts_of(ileqg), x_0, u_array, [θ]; eps_cap=cap)
    r.status[1] == 0 || throw(status_error(r.status[1]))
    N = problem.N
    ileqg.x_array, ileqg.l_array, ileqg.L_array = unpack_traj(r.x, r.l, r.L, N)
    ileqg.value_current = r.value[1]; ileqg.iter_current = r.iters[1]; ileqg.d_current = r.d_current[1]; ileqg.μ = r.mu[1]
    ileqg.ϵ_history = [(r.eps_hist[1, i, 1], r.eps_hist[2, i, 1]) for i in 1:r.trials[1]]
    return copy(ileqg.x_array), copy(ileqg.l_array), copy(ileqg.L_array), ileqg.value_current, copy(ileqg.ϵ_history)
end
RATiLQR.solve!(ileqg::ILEQGSolver, p::DeviceProblem, x_0::Vector{Float64}, u_array::Vector{Vector{Float64}};
               θ::Float64, verbose=true) = device_solve!(ileqg, p.problem, x_0, u_array; θ=θ, verbose=verbose)

# ---- compute_cost (cross_entropy...jl:173-227) / compute_cost_worker (nelder_mead...jl:134-158): ONE launch ---------
function device_costs(solver, problem::RSProblem, x, u_array, θ_array, kl_bound)
    r = solve_batch(problem, bilevel_opts(solver), x, u_array, Float64.(θ_array); want_traj=false)
    [r.status[i] == 0 ? r.value[i] + kl_bound / θ_array[i] : Inf for i in eachindex(θ_array)]
end
RATiLQR.compute_cost(ce::CrossEntropyBilevelOptimizationSolver, p::DeviceProblem, x::Vector{Float64},
                     u_array::Vector{Vector{Float64}}, θ_array::Vector{Float64}, kl_bound::Float64) =
    device_costs(ce, p.problem, x, u_array, θ_array, kl_bound)
function RATiLQR.compute_cost_serial(ce::CrossEntropyBilevelOptimizationSolver, p::DeviceProblem, x::Vector{Float64},
                                     u_array::Vector{Vector{Float64}}, θ_array::Vector{Float64}, kl_bound::Float64)
    @assert length(θ_array) == ce.num_samples   # cross_entropy...jl:204
    device_costs(ce, p.problem, x, u_array, θ_array, kl_bound)
end
RATiLQR.compute_cost_worker(nm::NelderMeadBilevelOptimizationSolver, p::DeviceProblem, x::Vector{Float64},
                            u_array::Vector{Vector{Float64}}, θ::Float64, kl_bound::Float64) =
    device_costs(nm, p.problem, x, u_array, [θ], kl_bound)[1]

# ---- RAT iLQR for P problems: ratilqr_ce_solve_fleet (whole CE loop on the device) ----------------------------------
# x0: n x P; cost_params: ncp x P (one block per problem) or a vector (shared); z: nz x P standard normals (problem p
# consumes column p in order, exactly like successive rand(rng, Normal(μ, σ)) = μ + σ randn(rng)) or nothing -> Philox(seed)
function ce_solve_fleet(ce::CrossEntropyBilevelOptimizationSolver, problem::RSProblem, x0::Matrix{Float64},
                        u_array::Vector{Vector{Float64}}, kl_bound::Float64; cost_params=problem.c.params,
                        μ_init=fill(ce.μ_init, size(x0, 2)), σ_init=fill(ce.σ_init, size(x0, 2)),
                        z::Union{Nothing,Matrix{Float64}}=nothing, seed::UInt64=UInt64(0), want_traj::Bool=true)
    b = pack_problem(problem); n, m, N = b.n, b.m, b.N; P = size(x0, 2)
    cpv = vec(Float64.(cost_params)); count = cost_params isa AbstractMatrix ? size(cost_params, 2) : 1
    U = reduce(hcat, u_array)
    μi = Float64.(μ_init); σi = Float64.(σ_init)
    θ_opt = zeros(P); value = zeros(P); θ_min = zeros(P); θ_max = zeros(P); μ = zeros(P); σ = zeros(P)
    nz_used = zeros(Int64, P); rounds = Ref{Int32}(0)
    x = want_traj ? zeros(n, N + 1, P) : zeros(0); l = want_traj ? zeros(m, N, P) : zeros(0)
    L = want_traj ? zeros(m, n, N, P) : zeros(0); status = zeros(Int32, P)
    nz = z === nothing ? 0 : size(z, 1)
    GC.@preserve b cpv U x0 μi σi θ_opt value θ_min θ_max μ σ nz_used x l L status z begin
        desc = desc_of(b, cpv, count)
        opts = bilevel_opts(ce)
        co = CeOpts(ce.num_samples, ce.num_elite, ce.iter_max, ce.λ, ce.use_θ_max)
        out = IleqgOut(want_traj ? pointer(x) : C_NULL, want_traj ? pointer(l) : C_NULL, want_traj ? pointer(L) : C_NULL,
                       C_NULL, pointer(status), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0)
        rc = ccall((:ratilqr_ce_solve_fleet, LIB), Int32,
                   (Ptr{Cvoid}, Ref{ProblemDesc}, Ref{IleqgOpts}, Ref{CeOpts}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Int32,
                    Float64, Ptr{Float64}, Int64, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ref{Int32}, Ref{IleqgOut}),
                   b200_context(), desc, opts, co, P, pointer(x0), P, pointer(U), 1, kl_bound,
                   z === nothing ? Ptr{Float64}(C_NULL) : pointer(z), nz, seed, pointer(μi), pointer(σi), pointer(θ_opt),
                   pointer(value), pointer(θ_min), pointer(θ_max), pointer(μ), pointer(σ), pointer(nz_used), rounds, out)
        rc == 0 || return (rc=rc, error=last_error())
    end
    (rc=Int32(0), θ_opt=θ_opt, value=value, θ_min=θ_min, θ_max=θ_max, μ=μ, σ=σ, μ_init=μi, σ_init=σi, nz_used=nz_used,
     rounds=rounds[], status=status, x=x, l=l, L=L)
end

# solve!(::CrossEntropyBilevelOptimizationSolver, ...) cross_entropy...jl:364-415 for ONE problem: the fleet call with P = 1.
# The θ draws come from the caller's rng exactly as the reference draws them: standard normals are generated one by one
# from a COPY of rng and injected; afterwards rng itself is advanced by the number of normals the solve consumed.
function device_solve!(ce::CrossEntropyBilevelOptimizationSolver, problem::RSProblem, x_0::Vector{Float64},
                       u_array::Vector{Vector{Float64}}, rng::AbstractRNG; kl_bound::Float64, verbose=true, serial=false)
    @assert kl_bound >= 0 "KL Divergence Bound must be non-negative"   # :368
    nz = max(4096, 64 * ce.num_samples * ce.iter_max)
    while true
        r2 = copy(rng)
        z = reshape([randn(r2) for _ in 1:nz], nz, 1)
        r = ce_solve_fleet(ce, problem, reshape(x_0, :, 1), u_array, kl_bound; z=z)
        if r.rc != 0
            (r.rc == -5 && nz < 1 << 24) || error("ratilqr_ce_solve_fleet failed ($(r.rc)): $(r.error)")
            nz *= 4        # the injected stream was exhausted (long redraw phase): retry with a longer one
            continue
        end
        for _ in 1:r.nz_used[1]; randn(rng); end
        ce.μ_init, ce.σ_init, ce.μ, ce.σ = r.μ_init[1], r.σ_init[1], r.μ[1], r.σ[1]   # persist like the reference (:66-68,297-301)
        ce.θ_min, ce.θ_max = kl_bound > 0 ? (r.θ_min[1], r.θ_max[1]) : (Inf, 0.0)
        ce.iter_current = kl_bound > 0 ? ce.iter_max : 0
        xa, la, La = unpack_traj(r.x, r.l, r.L, problem.N)
        return r.θ_opt[1], xa, la, La, r.value[1], r.θ_min[1], r.θ_max[1]
    end
end
RATiLQR.solve!(ce::CrossEntropyBilevelOptimizationSolver, p::DeviceProblem, x_0::Vector{Float64}, u_array::Vector{Vector{Float64}},
               rng::AbstractRNG; kl_bound::Float64, verbose=true, serial=false) =
    device_solve!(ce, p.problem, x_0, u_array, rng; kl_bound=kl_bound, verbose=verbose, serial=serial)

# ---- RAT iLQR++ for P problems: ratilqr_nm_solve_fleet ---------------------------------------------------------------
function nm_solve_fleet(nm::NelderMeadBilevelOptimizationSolver, problem::RSProblem, x0::Matrix{Float64},
                        u_array::Vector{Vector{Float64}}, kl_bound::Float64; cost_params=problem.c.params,
                        θ_high_init=fill(nm.θ_high_init, size(x0, 2)), θ_low_init=fill(nm.θ_low_init, size(x0, 2)),
                        c_high=zeros(size(x0, 2)), c_low=zeros(size(x0, 2)), has_c=zeros(Int32, 2, size(x0, 2)),
                        want_traj::Bool=true)
    b = pack_problem(problem); n, m, N = b.n, b.m, b.N; P = size(x0, 2)
    cpv = vec(Float64.(cost_params)); count = cost_params isa AbstractMatrix ? size(cost_params, 2) : 1
    U = reduce(hcat, u_array)
    thi = Float64.(θ_high_init); tli = Float64.(θ_low_init); ch = Float64.(c_high); cl = Float64.(c_low); hc = Int32.(has_c)
    θ_opt = zeros(P); value = zeros(P); iters = zeros(Int32, P); evals = zeros(Int32, P); status = zeros(Int32, P)
    x = want_traj ? zeros(n, N + 1, P) : zeros(0); l = want_traj ? zeros(m, N, P) : zeros(0)
    L = want_traj ? zeros(m, n, N, P) : zeros(0)
    GC.@preserve b cpv U x0 thi tli ch cl hc θ_opt value iters evals status x l L begin
        desc = desc_of(b, cpv, count)
        opts = bilevel_opts(nm)
        no = NmOpts(nm.α, nm.β, nm.γ, nm.ϵ, nm.λ, nm.iter_max)
        out = IleqgOut(want_traj ? pointer(x) : C_NULL, want_traj ? pointer(l) : C_NULL, want_traj ? pointer(L) : C_NULL,
                       C_NULL, pointer(status), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0)
        rc = ccall((:ratilqr_nm_solve_fleet, LIB), Int32,
                   (Ptr{Cvoid}, Ref{ProblemDesc}, Ref{IleqgOpts}, Ref{NmOpts}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Int32,
                    Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Int32}, Ptr{Int32}, Ref{IleqgOut}),
                   b200_context(), desc, opts, no, P, pointer(x0), P, pointer(U), 1, kl_bound, pointer(thi), pointer(tli),
                   pointer(ch), pointer(cl), pointer(hc), pointer(θ_opt), pointer(value), pointer(iters), pointer(evals), out)
        rc == 0 || error("ratilqr_nm_solve_fleet failed ($rc): $(last_error())")
    end
    (θ_opt=θ_opt, value=value, nm_iters=iters, n_evals=evals, status=status, θ_high_init=thi, θ_low_init=tli,
     c_high=ch, c_low=cl, has_c=hc, x=x, l=l, L=L)
end

# solve!(::NelderMeadBilevelOptimizationSolver, ...) nelder_mead...jl:276-352 for ONE problem (fleet call with P = 1).
# θ_high_init / θ_low_init and the vertex costs c_high / c_low persist in the solver struct like in the reference
# (incl. its quirk that initialize! does not reset them, :164-168 vs :283,294).
function device_solve!(nm::NelderMeadBilevelOptimizationSolver, problem::RSProblem, x_0::Vector{Float64},
                       u_array::Vector{Vector{Float64}}; kl_bound::Float64, verbose=true)
    @assert kl_bound >= 0 "KL Divergence Bound must be non-negative"   # :280
    hc = Int32[nm.c_high === nothing ? 0 : 1, nm.c_low === nothing ? 0 : 1]
    r = nm_solve_fleet(nm, problem, reshape(x_0, :, 1), u_array, kl_bound;
                       c_high=[nm.c_high === nothing ? 0.0 : nm.c_high], c_low=[nm.c_low === nothing ? 0.0 : nm.c_low],
                       has_c=reshape(hc, 2, 1))
    r.status[1] == 0 || throw(status_error(r.status[1]))   # the final solve has no try/catch (:334-346)
    nm.θ_high_init, nm.θ_low_init = r.θ_high_init[1], r.θ_low_init[1]
    nm.c_high = r.has_c[1, 1] != 0 ? r.c_high[1] : nothing
    nm.c_low = r.has_c[2, 1] != 0 ? r.c_low[1] : nothing
    nm.iter_current = r.nm_iters[1]
    xa, la, La = unpack_traj(r.x, r.l, r.L, problem.N)
    return r.θ_opt[1], xa, la, La, r.value[1]
end
RATiLQR.solve!(nm::NelderMeadBilevelOptimizationSolver, p::DeviceProblem, x_0::Vector{Float64}, u_array::Vector{Vector{Float64}};
               kl_bound::Float64, verbose=true) = device_solve!(nm, p.problem, x_0, u_array; kl_bound=kl_bound, verbose=verbose)

# ---- noisy closed-loop rollouts + cost (ileqg.jl:94-109, :115-124): ratilqr_mc_rollout -------------------------------
# noise: n x N x n_samples injected w tensor, or nothing -> Philox(seed) coloured with chol(W).  Returns J and
# [mean, var, entropic risk] of the policy (x_array, l_array, L_array).
function mc_rollout(problem::RSProblem, x_array, l_array, L_array, n_samples::Integer; noise=nothing, seed::UInt64=UInt64(0),
                    θ_risk::Float64=0.0)
    b = pack_problem(problem); n, m, N = b.n, b.m, b.N
    X = reduce(hcat, x_array); Lm = reduce(hcat, l_array); LL = cat(L_array...; dims=3)
    J = zeros(n_samples); stats = zeros(3)
    GC.@preserve b X Lm LL J stats noise begin
        desc = desc_of(b)
        rc = ccall((:ratilqr_mc_rollout, LIB), Int32,
                   (Ptr{Cvoid}, Ref{ProblemDesc}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, UInt64,
                    Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   b200_context(), desc, 1, pointer(X), pointer(Lm), pointer(LL), n_samples,
                   noise === nothing ? Ptr{Float64}(C_NULL) : pointer(noise), seed, θ_risk, pointer(J), pointer(stats),
                   Ptr{Float64}(C_NULL))
        rc == 0 || error("ratilqr_mc_rollout failed ($rc): $(last_error())")
    end
    (J=J, mean=stats[1], var=stats[2], entropic_risk=stats[3])
end

# ---- PETS: solve!(::CrossEntropyDirectOptimizationSolver, ...) pets.jl:270-281 -> ratilqr_pets_solve -------------------
# The whole CEM loop (sample, roll out num_control_samples x num_trajectory_samples particles, elite refit) runs on the
# device with Philox noise keyed by a seed drawn from the caller's rng (Julia's MersenneTwister stream cannot be
# reproduced on the device: results match the reference statistically, SURVEY.md 8c).
function device_solve!(pets::CrossEntropyDirectOptimizationSolver, problem::GenProblem, x_0::Vector{Float64}, rng::AbstractRNG;
                       use_true_model=false, verbose=true, serial=true)
    fs = problem.f_stochastic; n, m = dims_of(fs.f); N = problem.N
    @assert N == pets.N
    RATiLQR.initialize!(pets)                                         # pets.jl:70-74
    μ = reduce(hcat, pets.μ_array); Σ = cat(pets.Σ_array...; dims=3)  # m x N, m x m x N
    mp = fs.f.params; cp = problem.c.params; Wbuf = vec(fs.W); ens = fs.ensemble_params
    seed = rand(rng, UInt64)
    GC.@preserve mp cp Wbuf ens x_0 μ Σ begin
        desc = ProblemDesc(fs.f.model_id, problem.c.id, n, m, N, pointer(mp), length(mp), pointer(cp), length(cp), 1, pointer(Wbuf), 0)
        gen = GenerativeDesc(fs.noise_kind, fs.noise_scale, max(size(ens, 2), 1), size(ens, 2) > 1 ? pointer(ens) : C_NULL,
                             C_NULL, 0)
        rc = ccall((:ratilqr_pets_solve, LIB), Int32,
                   (Ptr{Cvoid}, Ref{ProblemDesc}, Ref{GenerativeDesc}, Ptr{Float64}, Int32, Int32, Int32, Int32, Float64,
                    Ptr{Float64}, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Float64}),
                   b200_context(), desc, gen, pointer(x_0), pets.num_control_samples, pets.num_trajectory_samples, pets.num_elite,
                   pets.iter_max, pets.smoothing_factor, Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL), seed, pointer(μ), pointer(Σ))
        rc == 0 || error("ratilqr_pets_solve failed ($rc): $(last_error())")
    end
    pets.μ_array = [μ[:, t] for t in 1:N]; pets.Σ_array = [Σ[:, :, t] for t in 1:N]
    pets.iter_current = pets.iter_max
    return copy(pets.μ_array), copy(pets.Σ_array)
end
RATiLQR.solve!(pets::CrossEntropyDirectOptimizationSolver, p::DeviceGenerativeProblem, x_0::Vector{Float64}, rng::AbstractRNG;
               use_true_model=false, verbose=true, serial=true) =
    device_solve!(pets, p.problem, x_0, rng; use_true_model=use_true_model, verbose=verbose, serial=serial)

# ---- optional: make PLAIN reference problems with device callables take the device path --------------------------------
# Re-defines the reference's five entry points at RUN TIME (never in a precompiled module body).  A non-device problem is
# forwarded to the method that existed BEFORE the hook, looked up in the world age recorded here -- Base.invoke_in_world
# (Julia >= 1.6) -- so the hook can never call itself.
const HOOK_WORLD = Ref{UInt}(0)
function install_hooks!()
    HOOK_WORLD[] != 0 && return nothing
    isdefined(Base, :invoke_in_world) || error("install_hooks! needs Julia >= 1.6 (Base.invoke_in_world); use device(problem) instead")
    HOOK_WORLD[] = Base.get_world_counter()
    @eval begin
        function RATiLQR.solve!(ileqg::ILEQGSolver, problem::RSProblem, x_0::Vector{Float64}, u_array::Vector{Vector{Float64}};
                                θ::Float64, verbose=true)
            is_device(problem) && return device_solve!(ileqg, problem, x_0, u_array; θ=θ, verbose=verbose)
            Base.invoke_in_world(HOOK_WORLD[], RATiLQR.solve!, ileqg, problem, x_0, u_array; θ=θ, verbose=verbose)
        end
        function RATiLQR.compute_cost(ce::CrossEntropyBilevelOptimizationSolver, problem::RSProblem, x::Vector{Float64},
                                      u_array::Vector{Vector{Float64}}, θ_array::Vector{Float64}, kl_bound::Float64)
            is_device(problem) && return device_costs(ce, problem, x, u_array, θ_array, kl_bound)
            Base.invoke_in_world(HOOK_WORLD[], RATiLQR.compute_cost, ce, problem, x, u_array, θ_array, kl_bound)
        end
        function RATiLQR.compute_cost_serial(ce::CrossEntropyBilevelOptimizationSolver, problem::RSProblem, x::Vector{Float64},
                                             u_array::Vector{Vector{Float64}}, θ_array::Vector{Float64}, kl_bound::Float64)
            is_device(problem) && (@assert length(θ_array) == ce.num_samples; return device_costs(ce, problem, x, u_array, θ_array, kl_bound))
            Base.invoke_in_world(HOOK_WORLD[], RATiLQR.compute_cost_serial, ce, problem, x, u_array, θ_array, kl_bound)
        end
        function RATiLQR.compute_cost_worker(nm::NelderMeadBilevelOptimizationSolver, problem::RSProblem, x::Vector{Float64},
                                             u_array::Vector{Vector{Float64}}, θ::Float64, kl_bound::Float64)
            is_device(problem) && return device_costs(nm, problem, x, u_array, [θ], kl_bound)[1]
            Base.invoke_in_world(HOOK_WORLD[], RATiLQR.compute_cost_worker, nm, problem, x, u_array, θ, kl_bound)
        end
        function RATiLQR.solve!(pets::CrossEntropyDirectOptimizationSolver, problem::GenProblem, x_0::Vector{Float64},
                                rng::AbstractRNG; use_true_model=false, verbose=true, serial=true)
            is_device(problem) && return device_solve!(pets, problem, x_0, rng; use_true_model=use_true_model, verbose=verbose, serial=serial)
            Base.invoke_in_world(HOOK_WORLD[], RATiLQR.solve!, pets, problem, x_0, rng; use_true_model=use_true_model,
                                 verbose=verbose, serial=serial)
        end
    end
    return nothing
end

end # module
