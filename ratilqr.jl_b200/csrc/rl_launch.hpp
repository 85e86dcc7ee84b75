// rl_launch.hpp -- launchers implemented in the rl_kernels_*.cu translation units.
// All pointers are DEVICE pointers.  Every launcher returns 0, or -1 if the (model, cost)
// pair is not compiled in; CUDA errors are checked by the caller (cudaGetLastError).
#pragma once
#include <cuda_runtime.h>

#include "rl_args.hpp"
#include "rl_core.cuh"

namespace rll {

// persistent one-thread-per-instance iLEQG solve (rl_kernels_solve.cu)
int launch_solve(int model_id, int cost_id, const rl::SolveParams& P, cudaStream_t st);

// warp-cooperative variant (rl_coop.cuh): one warp per instance; traj_global is the fallback trajectory storage
// (B * coop_traj_doubles) used only when matrices + trajectories exceed the shared memory of an SM
int coop_smem_query(int model_id, int cost_id, int N, size_t* smem_bytes);
int launch_solve_coop(int model_id, int cost_id, const rl::SolveParams& P, double* traj_global, cudaStream_t st);

// speculative latency kernel for small batches (rl_spec.cuh): G = 2, 4 or 8 lanes per instance; the workspace holds one
// column per LANE (B * G columns) with double-buffered Lg / DL; x, l, L are written in host layout to P.xo / lo / Lo
bool spec_supported(int model_id, int cost_id);
int launch_solve_spec(int model_id, int cost_id, int G, const rl::SolveParams& P, cudaStream_t st);

// two-warp speculative variant of the warp-cooperative kernel (rl_coop2.cuh): latency path of the n > 6 models
size_t coop2_smem_query(int model_id, int cost_id, int N);
int launch_solve_coop2(int model_id, int cost_id, const rl::SolveParams& P, cudaStream_t st);

// SoA workspace -> host-layout outputs (x, l, L), tile transpose through shared memory
// (X, U, Lg: sections of the tiled workspace, `rec` = elements per slot record, see rl::SolveParams)
void launch_gather(int n, int m, int N, int B, const double* X, const double* U, const double* Lg, size_t rec, const int32_t* cur,
                   const int32_t* perm, double* x_out, double* l_out, double* L_out, cudaStream_t st);

// per-problem ascending sort of theta -> slot-to-instance permutation, problems laid out in `order` (device, P entries,
// or null = natural order); returns -1 when not applicable (caller keeps the identity)
int launch_sort_theta(const double* theta, int P, int K, const int32_t* order, int32_t* perm, cudaStream_t st);
// key[p] = max iterations over the K instances of problem p (device -> device)
void launch_problem_work(const int32_t* iters, int P, int K, int32_t* key, cudaStream_t st);

int launch_rollout_open(const CompArgs& a, cudaStream_t st);
int launch_rollout_closed(const CompArgs& a, cudaStream_t st);
int launch_integrate_cost(const CompArgs& a, cudaStream_t st);
int launch_linearize(const CompArgs& a, cudaStream_t st);

struct RiccatiArgs {
  int n, m, N, B, optimise;
  const double *q, *qv, *Q, *r, *R, *Pm, *A, *Bm, *W, *Winv;
  const double* detW;  // device: 1 or N determinants
  int W_tv;
  const double* theta;
  double mu_min, delta_0;
  double *mu, *delta, *L, *dl, *s, *sv, *S;
  int has_dl;
  int32_t *status, *restarts;
};
int launch_riccati(const RiccatiArgs& a, cudaStream_t st);

int launch_mc_rollout(const McArgs& a, cudaStream_t st);
// per-problem mean / unbiased variance / entropic risk of J (deterministic single-block reduction)
// scratch: mc_stats_scratch_doubles(n_samples, P) doubles (0 = the one-CTA-per-problem kernel is used)
void launch_mc_stats(const double* J, int n_samples, int P, double theta_risk, double* stats, double* scratch, cudaStream_t st);
size_t mc_stats_scratch_doubles(int n_samples, int P);

int launch_pets_costs(const PetsArgs& a, cudaStream_t st);
// u = mu_t + chol_lower(Sigma_t) z ; z injected (m*N*C) or Philox. returns via err[0] != 0 if a Sigma_t is not PD
void launch_pets_sample(int m, int N, int C, const double* mu, const double* Sigma, const double* z, uint64_t seed,
                        uint64_t stream_offset, double* controls, int32_t* err, cudaStream_t st);
// stable top-k elites + smoothed refit (pets.jl:159-191)
void launch_pets_refit(int m, int N, int C, int num_elite, double smoothing, const double* controls,
                       const double* cost, double* mu, double* Sigma, int32_t* elite_idx, int32_t* sort_ws,
                       cudaStream_t st);

// ---- RAT iLQR (CE over theta) for a fleet: per-problem state lives on the device ------------------------
struct CeFleet {
  int P, S, num_elite, iter_max, use_theta_max;
  double lambda, kl;
  const double* z; long long nz; unsigned long long seed;   // injected normals (P*nz) or Philox
  long long p0;             // global index of this block's first problem (Philox stream = p0 + p): sub-fleets draw what the whole fleet would
  double *mu_init, *sigma_init, *mu, *sigma, *theta_min, *theta_max, *theta_opt, *value_out;
  long long* cursor; int32_t *iter, *active, *err;
  double* theta;            // P*S draws (device), consumed by the solve kernel
  const double* value; const int32_t* status;   // results of the batched solve (P*S)
  int32_t* n_active;        // single counter
};
void launch_ce_draw(const CeFleet& c, cudaStream_t st);
void launch_ce_update(const CeFleet& c, double* elite_ws /* P * num_elite doubles */, cudaStream_t st);
void launch_ce_pick_theta(const CeFleet& c, double* theta_final, cudaStream_t st);
void launch_ce_final_update(const CeFleet& c, double* theta_final, const double* value, const int32_t* status, cudaStream_t st);

// ---- RAT iLQR++ (Nelder-Mead over theta) for a fleet ------------------------------------------------------------
struct NmFleet {
  int P, iter_max;
  double alpha, beta, gamma, eps, lambda, kl;
  double *th_high, *th_low, *th_high_init, *th_low_init, *c_high, *c_low, *theta_opt, *value_out;
  int32_t *has_c /* 2 per problem */, *iter, *active, *evals, *phase /* 0 = initial vertices, 1 = stepping */;
  double* theta;                               // P*6 candidate slots consumed by the solve kernel
  const double* value; const int32_t* status;  // its results
  int32_t* n_active;
};
void launch_nm_candidates(const NmFleet& c, cudaStream_t st);
void launch_nm_decide(const NmFleet& c, cudaStream_t st);
void launch_nm_final(const NmFleet& c, double* theta_final, const double* value, const int32_t* status, int stage, cudaStream_t st);

// ---- receding-horizon MPC driver (ratilqr_mpc_fleet_run): true-system step + plan shift, one thread per problem -----
struct MpcArgs {
  int model_id, N, P, steps, t;
  long long p0;                 // global index of this block's first problem (Philox stream)
  double mp[8];
  const double* plan;           // m*N*P  l_array of the plans just computed (host layout, device memory)
  double* x;                    // n*P    current true states (in/out)
  double* u_init;               // m*N*P  warm start of the next step (out)
  const double* noise;          // n*steps*P injected disturbances (per problem: n x steps), or null
  const double* cholW; uint64_t seed; rl::MixtureView mix;
  const double *theta_opt, *value;        // per-problem results of this step's plan
  double *x_traj, *u_traj, *theta_traj, *value_traj;  // n*(steps+1)*P, m*steps*P, steps*P, steps*P (problem slowest)
  int32_t* err;
};
int launch_mpc_advance(const MpcArgs& a, cudaStream_t st);

// DFMA throughput probe: returns total flops issued
double launch_fp64_probe(double* sink, int iters, cudaStream_t st);

}  // namespace rll
