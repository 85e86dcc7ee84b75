/*
 * ratilqr.h -- C ABI of libratilqr_b200.so
 *
 * B200-native (sm_100a) implementation of the data-parallel hot path of
 * StanfordMSL/RATiLQR.jl: batched iLEQG solves fanned out over risk-sensitivity
 * samples theta, noisy closed-loop Monte Carlo rollouts, and PETS CEM rollouts.
 *
 * The reference is pure Julia and has no FFI seam of its own: the
 * boundary it offers is its exported Julia API.  Every entry point below therefore
 * cites the reference function (file:line under the reference tree) that a Julia
 * `ccall` shim -- or the Python ctypes mirror in ratilqr.jl_b200/ -- replaces with it.
 *
 * Conventions
 *   - plain C, no exceptions; every function returns int32 (0 = ok, <0 = API misuse /
 *     CUDA error; message from ratilqr_last_error()).
 *   - NUMERICAL failure is not an error: it is reported per instance in status[b]
 *     (RATILQR_ST_*) and mapped to value +Inf, exactly where the reference's
 *     try/catch does (cross_entropy_bilevel_optimization.jl:161-165).
 *   - all matrices column-major (what Julia produces), instance index slowest.
 *   - caller owns every host array; the library never keeps a host pointer past return.
 *   - calls are blocking; one caller thread per ctx; one ctx drives one CUDA device.
 *   - stage index k is 0-based (optimal_control_problems.jl:28,35).
 */
#ifndef RATILQR_H
#define RATILQR_H

#ifndef __CUDACC_RTC__ /* NVRTC: the fixed-width types come from rl_core.cuh */
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------
 * Registered device models (user closures cannot run on the GPU, so the
 * accelerated path covers a registered set).  Equations: DESIGN.md "Model registry".
 * ------------------------------------------------------------------------------- */
enum {
  RATILQR_MODEL_SINGLE_INTEGRATOR = 1, /* n=2 m=2  x+ = x + dt*u            params [dt]            (test/ileqg_test.jl:12) */
  RATILQR_MODEL_POWER_LAW         = 2, /* n=2 m=2  x+ = x.^a + u.^b         params [a,b]           (test/ileqg_test.jl:151) */
  RATILQR_MODEL_DOUBLE_INTEGRATOR = 3, /* n=4 m=2  p+=p+dt*v, v+=v+dt*u     params [dt] */
  RATILQR_MODEL_PENDULUM          = 4, /* n=2 m=1                           params [dt,g,len,mass,damping] */
  RATILQR_MODEL_CARTPOLE          = 5, /* n=4 m=1                           params [dt,m_cart,m_pole,len,g] */
  RATILQR_MODEL_UNICYCLE          = 6, /* n=4 m=2  (px,py,psi,v ; a,omega)  params [dt] */
  RATILQR_MODEL_QUADROTOR         = 7  /* n=12 m=4                          params [dt,mass,g,Ixx,Iyy,Izz] */
};

enum {
  /* c(k,x,u) = (ws0+ws1*k)*(1/2 dx'Q dx + 1/2 u'R u + dx'Pc u) + c0 + c1*k,  dx = x - xg
   * h(x)     = 1/2 dx'Qf dx + h0
   * params   = [ws0, ws1, c0, c1, h0, xg(n), Q(n*n), R(m*m), Pc(n*m), Qf(n*n)]  (Q,R,Qf symmetric)
   * covers optimal_control_problems.jl:59-61, test/ileqg_test.jl:13-14,53-54,68-69 */
  RATILQR_COST_QUADRATIC  = 1,
  /* c = sum(x.^p) + sum(u.^p), h = h0 ; params [p, h0]  (test/ileqg_test.jl:152-153) */
  RATILQR_COST_POWER_LAW  = 2,
  /* c = sum(abs(u)), h = h0 ; params [h0] ; rollout-only, not differentiable (test/pets_test.jl:16-17) */
  RATILQR_COST_L1_CONTROL = 3
};

/* per-instance status (replaces Julia exceptions) */
enum {
  RATILQR_ST_OK              = 0,
  RATILQR_ST_M_NOT_PD_INIT   = 1, /* @assert isposdef(M) inside initialize!           ileqg.jl:234,440 */
  RATILQR_ST_M_NOT_PD_OPT    = 2, /* @assert isposdef(M) inside solve_approximate_dp! ileqg.jl:366 */
  RATILQR_ST_DOMAIN          = 3, /* DomainError in the model/cost (negative base of a real power) */
  RATILQR_ST_LINESEARCH_HANG = 4, /* reference would loop forever: ileqg.jl:526-535 has no eps_min test */
  RATILQR_ST_MU_OVERFLOW     = 5  /* regularisation restart loop ileqg.jl:359-401 did not terminate */
};

typedef struct ratilqr_ctx ratilqr_ctx;

/* "True model" process noise: a Gaussian mixture  sum_c weights[c] N(means[:,c], covs[:,:,c])
 * -- the accurate GMM of the reference's generative example, selected there by
 * f_stochastic(x, u, rng, use_true_model=true) (src/optimal_control_problems.jl:85-86,
 * 102-115), against the single Gaussian the planner assumes. */
typedef struct {
  int32_t n_components;   /* >= 1 */
  const double* weights;  /* n_components, > 0 (normalised by the library) */
  const double* means;    /* n * n_components, column-major */
  const double* covs;     /* n*n * n_components, each positive definite */
} ratilqr_noise_mixture;

/* f, c, h, W, N of FiniteHorizonRiskSensitiveOptimalControlProblem
 * (optimal_control_problems.jl:67-73), restricted to registered models. */
typedef struct {
  int32_t model_id, cost_id;
  int32_t n, m, N;
  const double* model_params; int32_t n_model_params;
  const double* cost_params;  int32_t n_cost_params;     /* length of ONE parameter block */
  int32_t cost_params_count;                             /* 1 = shared, P = one block per problem */
  const double* W;                                       /* n*n, or n*n*N when W_time_varying */
  int32_t W_time_varying;
} ratilqr_problem_desc;

/* ILEQGSolver keyword arguments, 1:1 with ileqg.jl:165-175,191-194 */
typedef struct {
  double  mu_min;            /* 1e-6 */
  double  delta_0;           /* 2.0  */
  double  lambda;            /* 0.5  */
  double  d;                 /* 1e-2 */
  int32_t iter_max;          /* 100  */
  int32_t adaptive_eps_init; /* 0    */
  double  eps_init;          /* 1.0  */
  double  eps_min;           /* 1e-6 */
  int32_t f_returns_jacobian;/* accepted for API parity; device Jacobians are always analytic */
} ratilqr_ileqg_opts;

/* Batch shape: P problems x K theta samples per problem, instance b = p*K + j.
 * x0 is n*x0_count, u_init is m*N*u_count, with the counts 1 (shared) or P. */
typedef struct {
  int32_t P, K;
  const double* x0;     int32_t x0_count;
  const double* u_init; int32_t u_count;
  const double* theta;  /* P*K */
} ratilqr_batch_in;

/* Any pointer may be NULL (output skipped, nothing copied back). B = P*K. */
typedef struct {
  double*  x;          /* n*(N+1)*B   x_array          ileqg.jl:655 */
  double*  l;          /* m*N*B       l_array */
  double*  L;          /* m*n*N*B     L_array */
  double*  value;      /* B           value_current (+Inf when status != 0) */
  int32_t* status;     /* B */
  int32_t* iters;      /* B  iter_current */
  int32_t* trials;     /* B  line-search merit evaluations (entries of eps_history) */
  int32_t* restarts;   /* B  calls of increase_mu_and_delta! */
  double*  mu;         /* B  final mu */
  double*  d_current;  /* B  final d_current */
  double*  eps_hist;   /* 2*eps_hist_cap*B  (eps, new-cur) pairs     ileqg.jl:537 */
  int32_t  eps_hist_cap;
} ratilqr_ileqg_out;

/* ---- context ------------------------------------------------------------------ */
int32_t ratilqr_create(ratilqr_ctx** ctx, int32_t device_id);
int32_t ratilqr_destroy(ratilqr_ctx* ctx);
const char* ratilqr_last_error(const ratilqr_ctx* ctx);
int32_t ratilqr_version(void);
/* static description of a registered model: dims and parameter counts (0 if unknown id) */
int32_t ratilqr_model_dims(int32_t model_id, int32_t* n, int32_t* m, int32_t* n_params);
int32_t ratilqr_cost_param_count(int32_t cost_id, int32_t n, int32_t m);

/* ---- the hot path: solve!(::ILEQGSolver, ...) for a whole batch ------------------
 * replaces ileqg.jl:635-659 (initialize! :214-236, step! :598-613, solve_approximate_dp!
 * :341-406, line_search! :494-592) run once per instance; one persistent kernel. */
int32_t ratilqr_ileqg_solve_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc,
                                  const ratilqr_ileqg_opts* opts,
                                  const ratilqr_batch_in* in, ratilqr_ileqg_out* out);

/* compute_cost / compute_cost_serial of RAT iLQR (cross_entropy_bilevel_optimization.jl
 * :173-227) and compute_cost_worker of RAT iLQR++ (nelder_mead_bilevel_optimization.jl
 * :134-158):  cost[b] = value[b] + kl_bound/theta[b], +Inf on any failure. */
int32_t ratilqr_ce_costs(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc,
                         const ratilqr_ileqg_opts* opts, const ratilqr_batch_in* in,
                         double kl_bound, double* cost, int32_t* status);

/* solve!(::CrossEntropyBilevelOptimizationSolver, ...) (cross_entropy_bilevel_optimization.jl:364-415) for a
 * FLEET of P independent problems advanced in lock-step rounds: per round every still-active problem draws
 * num_samples positive theta (get_positive_samples :233-246), ONE batched launch solves all P*num_samples
 * instances, then the feasibility / redraw logic (:291-311), theta_min/max bookkeeping (:314-324) and the
 * elite refit (:326-334) run per problem on the device.  Final solve with the retry rule of :390-414.
 * z_inject: P*nz standard normals (row p = the stream rand(rng, Normal) of problem p), or NULL -> Philox(seed).
 * mu_init / sigma_init: P values, in-out (they persist across solve! calls, :66-68).  Outputs are P long;
 * `final` (B = P) receives the trajectories / status of the final solve. */
typedef struct {
  int32_t num_samples, num_elite, iter_max; double lambda; int32_t use_theta_max;
} ratilqr_ce_opts;
int32_t ratilqr_ce_solve_fleet(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                               const ratilqr_ce_opts* ce, int32_t P, const double* x0, int32_t x0_count,
                               const double* u_init, int32_t u_count, double kl_bound,
                               const double* z_inject, int64_t nz, uint64_t seed,
                               double* mu_init, double* sigma_init,
                               double* theta_opt, double* value, double* theta_min, double* theta_max,
                               double* mu, double* sigma, int64_t* nz_used, int32_t* rounds,
                               ratilqr_ileqg_out* final);

/* Receding-horizon RAT iLQR for a fleet of P independent systems, `steps` MPC steps WITHOUT leaving the device -- the
 * caller the reference does not ship (its solve! is one planning step).  Per step and problem:
 * plan with solve!(::CrossEntropyBilevelOptimizationSolver) from the current state, warm-started with the shifted
 * previous plan (ratilqr_ce_solve_fleet's on-device loop); apply the first nominal control to the TRUE system
 * x+ = f(x, l_0) + w; shift the plan (last control repeated).  States, plans and disturbances never visit the host;
 * mu_init / sigma_init (P, in-out) persist across the steps and the call exactly like the fields of the reference's solver
 * struct (cross_entropy_bilevel_optimization.jl:66-68, 297-301).
 * Disturbance w: `noise` (n*steps*P injected: w of problem p at step t is noise[(p*steps + t)*n ...]) or NULL ->
 * Philox(noise_seed) drawn from N(0, W) or, when true_noise != NULL, from the true-model Gaussian mixture.
 * z_inject: steps * P * nz standard normals for the theta draws (step slowest, then problem) or NULL -> Philox(seed + t).
 * Outputs (problem slowest): x_traj n*(steps+1)*P, u_traj m*steps*P, theta_traj / value_traj steps*P (theta_opt and
 * value + kl/theta_opt of every plan), step_ms steps (host wall time per step, max over the concurrent sub-fleets). */
typedef struct {
  int32_t steps;
  const double* noise;
  uint64_t noise_seed;
  const ratilqr_noise_mixture* true_noise;
} ratilqr_mpc_opts;
int32_t ratilqr_mpc_fleet_run(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                              const ratilqr_ce_opts* ce, const ratilqr_mpc_opts* mpc, int32_t P, const double* x0,
                              const double* u_init, int32_t u_count, double kl_bound,
                              const double* z_inject, int64_t nz, uint64_t seed,
                              double* mu_init, double* sigma_init,
                              double* x_traj, double* u_traj, double* theta_traj, double* value_traj,
                              float* step_ms, int32_t* rounds_total);

/* The same solve! for ONE problem -- the reference's own call shape (cross_entropy_bilevel_optimization.jl:364-367):
 * x0 n, u_init m*N, scalars in / out; whole CE loop on the device incl. the elite selection over the theta population
 * (one CTA ranks a population of up to thousands of samples, e.g. configs[1]'s 1024).  Equivalent to the fleet call
 * with P = 1. */
int32_t ratilqr_ce_solve(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                         const ratilqr_ce_opts* ce, const double* x0, const double* u_init, double kl_bound,
                         const double* z_inject, int64_t nz, uint64_t seed,
                         double* mu_init, double* sigma_init,
                         double* theta_opt, double* value, double* theta_min, double* theta_max,
                         double* mu, double* sigma, int64_t* nz_used, int32_t* rounds,
                         ratilqr_ileqg_out* final);

/* solve!(::NelderMeadBilevelOptimizationSolver, ...) (nelder_mead_bilevel_optimization.jl:276-352) for a fleet of P
 * independent problems in lock-step rounds.  Evaluations are pure functions of theta, so every round solves, in ONE
 * batched launch, all candidates the decision tree of step! (:174-252) can ask for (theta_r, theta_e, both possible
 * contractions, both shrink points) and then replays the tree per problem on the device: visited vertices and
 * returned values equal the serial order.  theta_high_init / theta_low_init (P, in-out) and the vertex costs
 * c_high / c_low with their has_c flags (P, in-out) persist across calls like the fields of the Julia struct. */
typedef struct {
  double alpha, beta, gamma, eps, lambda; int32_t iter_max;
} ratilqr_nm_opts;
int32_t ratilqr_nm_solve_fleet(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                               const ratilqr_nm_opts* nm, int32_t P, const double* x0, int32_t x0_count,
                               const double* u_init, int32_t u_count, double kl_bound,
                               double* theta_high_init, double* theta_low_init, double* c_high, double* c_low,
                               int32_t* has_c, double* theta_opt, double* value, int32_t* nm_iters, int32_t* n_evals,
                               ratilqr_ileqg_out* final);

/* The same solve! for ONE problem (nelder_mead_bilevel_optimization.jl:276-279); has_c: 2 flags (c_high, c_low known). */
int32_t ratilqr_nm_solve(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                         const ratilqr_nm_opts* nm, const double* x0, const double* u_init, double kl_bound,
                         double* theta_high_init, double* theta_low_init, double* c_high, double* c_low,
                         int32_t* has_c, double* theta_opt, double* value, int32_t* nm_iters, int32_t* n_evals,
                         ratilqr_ileqg_out* final);

/* Device-resident variant used for throughput measurement: stage once, run many times.
 * stage = H2D of inputs; run = the solve kernel only, `reps` launches back to back on the
 * ctx stream, bracketed by CUDA events recorded on that stream (ms_total out);
 * fetch = D2H of whatever `out` asks for. */
int32_t ratilqr_ileqg_stage(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc,
                            const ratilqr_ileqg_opts* opts, const ratilqr_batch_in* in);
int32_t ratilqr_ileqg_run(ratilqr_ctx* ctx, int32_t reps, float* ms_total);
int32_t ratilqr_ileqg_fetch(ratilqr_ctx* ctx, ratilqr_ileqg_out* out);

/* ---- component kernels, exposed for unit parity -------------------- */
/* simulate_dynamics open loop (ileqg.jl:18-38): x0 n*B, u m*N*B -> x n*(N+1)*B */
int32_t ratilqr_rollout_open_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t B,
                                   const double* x0, const double* u, double* x, int32_t* status);
/* simulate_dynamics closed loop (ileqg.jl:62-87): -> x_new n*(N+1)*B, u_new m*N*B */
int32_t ratilqr_rollout_closed_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t B,
                                     const double* xbar, const double* l, const double* L,
                                     double* x_new, double* u_new, int32_t* status);
/* integrate_cost (ileqg.jl:115-124) */
int32_t ratilqr_integrate_cost_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t B,
                                     const double* x, const double* u, double* cost, int32_t* status);
/* approximate_model (ileqg.jl:258-322): per instance q (N+1), qv n*(N+1), Q n*n*(N+1),
 * r m*N, R m*m*N, Pm m*n*N, A n*n*N, Bm n*m*N */
int32_t ratilqr_linearize_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t B,
                                const double* x, const double* u,
                                double* q, double* qv, double* Q, double* r, double* R,
                                double* Pm, double* A, double* Bm, int32_t* status);
/* the two Riccati passes on caller-supplied approximations (any n<=12, m<=4 registered size):
 * optimise != 0 : solve_approximate_dp! (ileqg.jl:341-406); L, dl are outputs; mu/delta in-out
 * optimise == 0 : solve_approximate_dp  (ileqg.jl:412-465); L input, dl input or NULL (zeros)
 * W is n*n (shared). outputs s (N+1)*B, sv n*(N+1)*B, S n*n*(N+1)*B. */
int32_t ratilqr_riccati_batch(ratilqr_ctx* ctx, int32_t n, int32_t m, int32_t N, int32_t B,
                              int32_t optimise,
                              const double* q, const double* qv, const double* Q, const double* r,
                              const double* R, const double* Pm, const double* A, const double* Bm,
                              const double* W, const double* theta,
                              double mu_min, double delta_0, double* mu, double* delta,
                              double* L, double* dl,
                              double* s, double* sv, double* S, int32_t* status, int32_t* restarts);
/* same with a time-varying noise covariance: W is n*n*N when W_time_varying != 0 and stage k uses
 * W(k) = W[:, :, k] exactly as ileqg.jl:364,438 index approx_result.W_array[ii] */
int32_t ratilqr_riccati_batch_tv(ratilqr_ctx* ctx, int32_t n, int32_t m, int32_t N, int32_t B,
                                 int32_t optimise,
                                 const double* q, const double* qv, const double* Q, const double* r,
                                 const double* R, const double* Pm, const double* A, const double* Bm,
                                 const double* W, int32_t W_time_varying, const double* theta,
                                 double mu_min, double delta_0, double* mu, double* delta,
                                 double* L, double* dl,
                                 double* s, double* sv, double* S, int32_t* status, int32_t* restarts);

/* ---- noisy closed-loop Monte Carlo rollouts ----------------------------------------
 * simulate_dynamics(problem, x_array, l_array, L_array, rng) + integrate_cost
 * (ileqg.jl:94-109,115-124) for n_samples noise realisations of ONE policy per problem.
 * noise: injected w tensor n*N*n_samples (per problem: n*N*n_samples*P) or NULL -> Philox(seed)
 * coloured with chol(W).  J n_samples*P.  stats per problem: [mean, var(unbiased),
 * entropic risk 1/theta*log mean exp(theta J) (theta_risk>0)]. */
int32_t ratilqr_mc_rollout(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t P,
                           const double* xbar, const double* l, const double* L,
                           int32_t n_samples, const double* noise, uint64_t seed,
                           double theta_risk, double* J, double* stats, double* x_out);

/* Closed-loop Monte Carlo evaluation under the TRUE noise model: same as
 * ratilqr_mc_rollout in Philox mode, but w_k is drawn from `true_noise` instead of N(0, W(k)):
 * E[J], Var[J] and the entropic risk of a policy that was optimised under the Gaussian model. */
int32_t ratilqr_mc_rollout_true_model(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t P,
                                      const double* xbar, const double* l, const double* L,
                                      int32_t n_samples, const ratilqr_noise_mixture* true_noise,
                                      uint64_t seed, double theta_risk, double* J, double* stats,
                                      double* x_out);

/* ---- PETS (pets.jl) ---------------------------------------------------------------- */
typedef struct {
  int32_t noise_kind;       /* 0 gaussian chol(W)*z ; 1 uniform[0,1)*scale (test/pets_test.jl:15) */
  double  noise_scale;
  int32_t n_ensemble;       /* model parameter sets; particle kk uses set kk / (particles/n_ensemble) */
  const double* ensemble_params; /* n_model_params * n_ensemble, or NULL -> desc.model_params */
  /* f_stochastic(x, u, rng, use_true_model) (optimal_control_problems.jl:82-87): when use_true_model != 0
   * and true_model != NULL, on-device (Philox) noise is drawn from the mixture instead of noise_kind */
  const ratilqr_noise_mixture* true_model;
  int32_t use_true_model;
} ratilqr_generative_desc;

/* compute_cost_serial (pets.jl:128-157): controls m*N*C, noise n*N*particles*C or NULL->Philox */
int32_t ratilqr_pets_costs(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc,
                           const ratilqr_generative_desc* gen, const double* x0,
                           const double* controls, int32_t C, int32_t particles,
                           const double* noise, uint64_t seed, double* cost);
/* get_elite_samples + compute_new_distribution (pets.jl:159-191) on device */
int32_t ratilqr_pets_refit(ratilqr_ctx* ctx, int32_t m, int32_t N, int32_t C, int32_t num_elite,
                           double smoothing, const double* controls, const double* cost,
                           double* mu, double* Sigma, int32_t* elite_idx);
/* step!/solve! (pets.jl:193-245,270-281): whole CEM loop on device. z_inject: m*N*C*iter_max
 * standard-normal draws or NULL -> Philox(seed). mu m*N, Sigma m*m*N are in-out. */
int32_t ratilqr_pets_solve(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc,
                           const ratilqr_generative_desc* gen, const double* x0,
                           int32_t C, int32_t particles, int32_t num_elite, int32_t iter_max,
                           double smoothing, const double* z_inject, const double* noise,
                           uint64_t seed, double* mu, double* Sigma);

/* ---- user-extensible device models ---------------------------------
 * The reference takes arbitrary Julia closures for f, c, h and differentiates them with
 * ForwardDiff (src/ileqg.jl:265-273: cx, cxx, cu, cuu, cux, fx, fu).  The device analogue: a
 * CUDA C++ snippet that defines the model once as a template over the scalar type,
 *   dynamics_src:  template <class T> void dynamics(const double* p, const T* x, const T* u, T* xn);
 *   cost_src:      template <class T> T stage_cost(const double* cp, int k, const T* x, const T* u);
 *                  template <class T> T terminal_cost(const double* cp, const T* x);
 * (k 0-based as in optimal_control_problems.jl:28,35; sin cos tan exp log sqrt tanh atan fabs
 * pow(T,double) square and + - * / < > work for every T).  The library compiles it at run time
 * (NVRTC, sm_100a) against its own kernels, instantiating T = double for rollouts and
 * forward-mode dual numbers (first order for A, B; second order for the cost's gradients and
 * Hessians) for approximate_model.  A NULL snippet selects a registered model / cost instead,
 * so user dynamics can be paired with RATILQR_COST_QUADRATIC and vice versa.
 * After registration, describe problems with  desc.model_id = *model_id_out,
 * desc.cost_id = RATILQR_COST_USER (cost_src given) or base_cost_id, and the declared n, m,
 * n_model_params (<= 8), n_cost_params.  Every entry point taking a ratilqr_problem_desc accepts
 * it (one thread per iLEQG instance; n <= 16, m <= 4).  A NaN/Inf produced by a snippet is that
 * instance's RATILQR_ST_DOMAIN (Julia: DomainError). */
#define RATILQR_MODEL_USER_BASE 1000
#define RATILQR_COST_USER 100
typedef struct {
  int32_t n, m;
  const char* dynamics_src; /* NULL -> registered model base_model_id */
  int32_t base_model_id;
  int32_t n_model_params;
  const char* cost_src;     /* NULL -> registered cost base_cost_id */
  int32_t base_cost_id;
  int32_t n_cost_params;    /* user cost: length of cp */
  /* Optional structure declarations (NULL = dense): one entry per matrix element, column-major,
   * 0 = identically zero, 1 = identically one, 2 = general.  a_kind n*n (df/dx), b_kind n*m (df/du) for
   * user dynamics; q_kind n*n (cxx and the terminal Hessian), r_kind m*m (cuu), p_kind m*n (cux) for a
   * user cost.  The kernels then skip / simplify those terms at compile time exactly like the registered
   * models do (bit-identical results when the declaration is true).  ratilqr_user_model_register
   * checks every declaration against the dual-number derivatives at random points and refuses a
   * model whose declared zero / one is violated there. */
  const int8_t* a_kind;
  const int8_t* b_kind;
  const int8_t* q_kind;
  const int8_t* r_kind;
  const int8_t* p_kind;
} ratilqr_user_model_desc;
/* compile only (NVRTC; no GPU, no ctx): 0 ok, -1 bad description, -20 NVRTC not loadable,
 * -21 compilation failed.  The compiler log is copied to log (NUL-terminated, truncated). */
int32_t ratilqr_user_model_check(const ratilqr_user_model_desc* um, char* log, int64_t log_cap);
/* compile + load into ctx; ids are per ctx and stay valid until ratilqr_destroy */
int32_t ratilqr_user_model_register(ratilqr_ctx* ctx, const ratilqr_user_model_desc* um,
                                    int32_t* model_id_out, char* log, int64_t log_cap);

/* ---- multi-GPU ---------------------------------------------------------------------------------
 * The reference's only parallelism is a scatter of independent samples + a gather of one cost per sample
 * (remotecall_fetch, cross_entropy_bilevel_optimization.jl:180-193, pets.jl:108-125).  Here a context can own an NCCL
 * communicator (loaded at run time); the *_sharded calls are collective: EVERY rank of the communicator calls them with
 * the same arguments, solves / rolls out its contiguous block of the population, and one ncclAllGather on DEVICE buffers
 * hands every rank the whole cost vector; elite selection then runs redundantly and deterministically on every rank, so
 * the CE state stays replicated without a broadcast.  Fleets of independent problems need no collective at all.
 *   multi-process (one process per GPU):  rank 0 calls ratilqr_nccl_unique_id, the 128 bytes travel to the other ranks
 *                                         by any means, every rank calls ratilqr_attach_comm on its ctx;
 *   single process, several GPUs:         ratilqr_create_multi (ncclCommInitAll) and the ratilqr_multi_* calls, which run
 *                                         the sharded call of every device on its own host thread. */
int32_t ratilqr_nccl_unique_id(uint8_t* id128);
int32_t ratilqr_attach_comm(ratilqr_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t world);
int32_t ratilqr_comm_info(const ratilqr_ctx* ctx, int32_t* rank, int32_t* world);
/* compute_cost for ONE problem's K-sample population (x0 n, u_init m*N, theta K): cost K, status K on every rank */
int32_t ratilqr_ce_costs_sharded(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                                 const double* x0, const double* u_init, const double* theta, int32_t K,
                                 double kl_bound, double* cost, int32_t* status);
/* ratilqr_ce_solve with the theta population of every CE iteration sharded (num_samples >= world size) */
int32_t ratilqr_ce_solve_sharded(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                                 const ratilqr_ce_opts* ce, const double* x0, const double* u_init, double kl_bound,
                                 const double* z_inject, int64_t nz, uint64_t seed,
                                 double* mu_init, double* sigma_init,
                                 double* theta_opt, double* value, double* theta_min, double* theta_max,
                                 double* mu, double* sigma, int64_t* nz_used, int32_t* rounds,
                                 ratilqr_ileqg_out* final);
/* ratilqr_pets_costs with the C action sequences sharded; Philox streams are indexed by the global sequence / particle
 * number, so the costs do not depend on the world size */
int32_t ratilqr_pets_costs_sharded(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc,
                                   const ratilqr_generative_desc* gen, const double* x0,
                                   const double* controls, int32_t C, int32_t particles,
                                   const double* noise, uint64_t seed, double* cost);

typedef struct ratilqr_multi ratilqr_multi;
int32_t ratilqr_create_multi(ratilqr_multi** out, const int32_t* device_ids, int32_t n_dev);
int32_t ratilqr_destroy_multi(ratilqr_multi* mg);
int32_t ratilqr_multi_size(const ratilqr_multi* mg);
ratilqr_ctx* ratilqr_multi_ctx(ratilqr_multi* mg, int32_t i);
const char* ratilqr_multi_last_error(const ratilqr_multi* mg);
int32_t ratilqr_multi_ce_costs(ratilqr_multi* mg, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                               const double* x0, const double* u_init, const double* theta, int32_t K,
                               double kl_bound, double* cost, int32_t* status);
int32_t ratilqr_multi_ce_solve(ratilqr_multi* mg, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                               const ratilqr_ce_opts* ce, const double* x0, const double* u_init, double kl_bound,
                               const double* z_inject, int64_t nz, uint64_t seed,
                               double* mu_init, double* sigma_init,
                               double* theta_opt, double* value, double* theta_min, double* theta_max,
                               double* mu, double* sigma, int64_t* nz_used, int32_t* rounds,
                               ratilqr_ileqg_out* final);
int32_t ratilqr_multi_pets_costs(ratilqr_multi* mg, const ratilqr_problem_desc* desc,
                                 const ratilqr_generative_desc* gen, const double* x0,
                                 const double* controls, int32_t C, int32_t particles,
                                 const double* noise, uint64_t seed, double* cost);
/* fleet of P independent RAT iLQR problems block-partitioned over the devices, no collective; l_out m*N*P or NULL */
int32_t ratilqr_multi_ce_solve_fleet(ratilqr_multi* mg, const ratilqr_problem_desc* desc,
                                     const ratilqr_ileqg_opts* opts, const ratilqr_ce_opts* ce, int32_t P,
                                     const double* x0, const double* u_init, int32_t u_count, double kl_bound,
                                     uint64_t seed, double* mu_init, double* sigma_init,
                                     double* theta_opt, double* value, double* l_out);

/* ---- measurement utilities --------------------------------------------------------- */
/* dependent-chain-free DFMA loop on every SM: returns achieved TFLOP/s (FP64, non-tensor) */
int32_t ratilqr_fp64_peak_probe(ratilqr_ctx* ctx, double* tflops, float* ms);
/* same loop launched back to back for `seconds` (power-capped steady state): the sustained FP64 figure */
int32_t ratilqr_fp64_peak_probe_sustained(ratilqr_ctx* ctx, double seconds, double* tflops);
/* number of kernel launches issued by this ctx since creation */
int64_t ratilqr_launch_count(const ratilqr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RATILQR_H */
