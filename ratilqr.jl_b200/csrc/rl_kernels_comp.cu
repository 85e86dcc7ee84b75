// rl_kernels_comp.cu -- component kernels (unit-parity API), Monte Carlo rollouts and PETS.
#include "rl_kernels_model.cuh"
#include "rl_host.hpp"
#include "rl_launch.hpp"

namespace rll {

using namespace rl;


#define RL_BLOCKS(total, th) (((total) + (th)-1) / (th))

int launch_rollout_open(const CompArgs& a, cudaStream_t st) {
#define X(MID, CID) if (a.model_id == MID) { k_rollout_open<Dyn<MID>><<<RL_BLOCKS(a.B, 64), 64, 0, st>>>(a); return 0; }
  RL_FOR_EACH_ILEQG_COMBO(X)
#undef X
  return -1;
}

// ---- receding-horizon driver: apply the first control of each new plan to the TRUE system and shift the plan ----------
// thread = problem.  x+ = f(x, l_0) + w with w injected (n per problem and step), Philox N(0, W) or the true-model
// Gaussian mixture; the plan l (m x N, host layout) becomes the next warm start: u_init[k] = l[k+1], last control repeated.
template <class D>
__global__ void k_mpc_advance(MpcArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.P) return;
  constexpr int n = D::n, m = D::m;
  const int N = a.N;
  const double* l = a.plan + (size_t)p * m * N;
  double x[n], u[m], xn[n], w[n];
  for (int i = 0; i < n; ++i) x[i] = a.x[(size_t)p * n + i];
  for (int j = 0; j < m; ++j) u[j] = l[j];
  int st = D::f(a.mp, x, u, xn) ? 0 : RATILQR_ST_DOMAIN;
  if (a.noise) { for (int i = 0; i < n; ++i) w[i] = a.noise[((size_t)p * a.steps + a.t) * n + i]; }
  else if (a.mix.k > 0) philox_mixture_noise<n>(a.seed, (uint64_t)(a.p0 + p), (uint32_t)a.t, a.mix, w);
  else philox_noise<n>(a.seed, (uint64_t)(a.p0 + p), (uint32_t)a.t, 0, 1.0, a.cholW, w);
  for (int i = 0; i < n; ++i) {
    const double v = st ? (double)NAN : xn[i] + w[i];
    a.x[(size_t)p * n + i] = v;
    a.x_traj[((size_t)p * (a.steps + 1) + a.t + 1) * n + i] = v;
  }
  for (int j = 0; j < m; ++j) a.u_traj[((size_t)p * a.steps + a.t) * m + j] = u[j];
  double* ui = a.u_init + (size_t)p * m * N;
  for (int k = 0; k < N; ++k)
    for (int j = 0; j < m; ++j) ui[(size_t)k * m + j] = l[(size_t)(k + 1 < N ? k + 1 : N - 1) * m + j];
  a.theta_traj[(size_t)p * a.steps + a.t] = a.theta_opt[p];
  a.value_traj[(size_t)p * a.steps + a.t] = a.value[p];
  if (st) atomicMax(a.err, st);
}

int launch_mpc_advance(const MpcArgs& a, cudaStream_t st) {
#define X(MID, CID) if (a.model_id == MID) { k_mpc_advance<Dyn<MID>><<<RL_BLOCKS(a.P, 64), 64, 0, st>>>(a); return 0; }
  RL_FOR_EACH_ILEQG_COMBO(X)
#undef X
  return -1;
}

int launch_rollout_closed(const CompArgs& a, cudaStream_t st) {
#define X(MID, CID) if (a.model_id == MID && a.cost_id == CID) { k_rollout_closed<Dyn<MID>, Cost<CID, Dyn<MID>::n, Dyn<MID>::m>><<<RL_BLOCKS(a.B, 64), 64, 0, st>>>(a); return 0; }
  RL_FOR_EACH_ILEQG_COMBO(X)
  RL_FOR_EACH_ROLLOUT_ONLY_COMBO(X)
#undef X
  return -1;
}

int launch_integrate_cost(const CompArgs& a, cudaStream_t st) {
#define X(MID, CID) if (a.model_id == MID && a.cost_id == CID) { k_integrate_cost<Dyn<MID>, Cost<CID, Dyn<MID>::n, Dyn<MID>::m>><<<RL_BLOCKS(a.B, 64), 64, 0, st>>>(a); return 0; }
  RL_FOR_EACH_ILEQG_COMBO(X)
  RL_FOR_EACH_ROLLOUT_ONLY_COMBO(X)
#undef X
  return -1;
}

int launch_linearize(const CompArgs& a, cudaStream_t st) {
#define X(MID, CID) if (a.model_id == MID && a.cost_id == CID) { k_linearize<Dyn<MID>, Cost<CID, Dyn<MID>::n, Dyn<MID>::m>><<<RL_BLOCKS(a.B * (a.N + 1), 64), 64, 0, st>>>(a); return 0; }
  RL_FOR_EACH_ILEQG_COMBO(X)
#undef X
  return -1;
}

// ---- the two Riccati passes on supplied approximations: thread = instance ----------------------
template <int n, int m>
__global__ void __launch_bounds__(64) k_riccati(RiccatiArgs a) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  int N = a.N;
  int32_t nr = 0;
  double mu = a.mu[b], delta = a.delta[b];
  int st = comp_riccati<n, m>(N, a.optimise, a.q + (size_t)b * (N + 1), a.qv + (size_t)b * n * (N + 1),
                              a.Q + (size_t)b * n * n * (N + 1), a.r + (size_t)b * m * N, a.R + (size_t)b * m * m * N,
                              a.Pm + (size_t)b * m * n * N, a.A + (size_t)b * n * n * N, a.Bm + (size_t)b * n * m * N,
                              a.W, a.Winv, a.detW, a.W_tv, a.theta[b], a.mu_min, a.delta_0, &mu, &delta,
                              a.L + (size_t)b * m * n * N, (a.optimise || a.has_dl) ? a.dl + (size_t)b * m * N : nullptr,
                              a.s + (size_t)b * (N + 1), a.sv + (size_t)b * n * (N + 1),
                              a.S + (size_t)b * n * n * (N + 1), &nr);
  a.mu[b] = mu;
  a.delta[b] = delta;
  if (a.status) a.status[b] = st;
  if (a.restarts) a.restarts[b] = nr;
}

int launch_riccati(const RiccatiArgs& a, cudaStream_t st) {
#define R(NN, MM) if (a.n == NN && a.m == MM) { k_riccati<NN, MM><<<RL_BLOCKS(a.B, 64), 64, 0, st>>>(a); return 0; }
  R(2, 2) R(2, 1) R(4, 2) R(4, 1) R(12, 4)
#undef R
  return -1;
}

int launch_mc_rollout(const McArgs& a, cudaStream_t st) {
  dim3 grid(RL_BLOCKS(a.n_samples, 128), a.P);
#define X(MID, CID) if (a.model_id == MID && a.cost_id == CID) { k_mc_rollout<Dyn<MID>, Cost<CID, Dyn<MID>::n, Dyn<MID>::m>><<<grid, 128, 0, st>>>(a); return 0; }
  RL_FOR_EACH_ILEQG_COMBO(X)
  RL_FOR_EACH_ROLLOUT_ONLY_COMBO(X)
#undef X
  return -1;
}

__global__ void __launch_bounds__(256) k_mc_stats(const double* J, int n_samples, double theta_risk, double* stats) {
  __shared__ double sh[33];
  int p = blockIdx.x;
  const double* Jp = J + (size_t)p * n_samples;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_samples; i += blockDim.x) acc += Jp[i];
  double mean = block_reduce(acc, OpAdd(), 0.0, sh) / n_samples;
  acc = 0.0;
  for (int i = threadIdx.x; i < n_samples; i += blockDim.x) { double d = Jp[i] - mean; acc += d * d; }
  double var = block_reduce(acc, OpAdd(), 0.0, sh);
  var = n_samples > 1 ? var / (n_samples - 1) : 0.0;
  double risk = mean;
  if (theta_risk > 0.0) {
    double mx = -HUGE_VAL;
    for (int i = threadIdx.x; i < n_samples; i += blockDim.x) mx = fmax(mx, theta_risk * Jp[i]);
    mx = block_reduce(mx, OpMax(), -HUGE_VAL, sh);
    acc = 0.0;
    for (int i = threadIdx.x; i < n_samples; i += blockDim.x) acc += exp(theta_risk * Jp[i] - mx);
    double se = block_reduce(acc, OpAdd(), 0.0, sh);
    risk = (mx + log(se / n_samples)) / theta_risk;
  }
  if (threadIdx.x == 0) { stats[3 * p] = mean; stats[3 * p + 1] = var; stats[3 * p + 2] = risk; }
}

// Large sample counts (one problem x 2^20 samples took 2 ms in the single CTA above): the same four reductions in chunks of
// RL_MC_CHUNK samples, one CTA per (chunk, problem); every CTA first combines the previous stage's per-chunk partials in
// chunk order (<= a few hundred values), so the result is a fixed function of (n_samples, data): deterministic, independent
// of the launch.  mode 0: sum x | 1: sum (x - mean)^2 | 2: max theta x | 3: sum exp(theta x - max) | 4: final combine.
constexpr int RL_MC_CHUNK = 8192;
__global__ void __launch_bounds__(256) k_mc_stats_chunked(const double* J, int n_samples, int n_chunks, double theta_risk, int mode,
                                                          double* part /* [P][4][n_chunks] */, double* stats) {
  __shared__ double sh[33];
  const int p = blockIdx.y, c = blockIdx.x;
  const double* Jp = J + (size_t)p * n_samples;
  double* pp = part + (size_t)p * 4 * n_chunks;
  auto combine = [&](int which, bool is_max) {  // every thread: the partials of stage `which` in chunk order
    double a = is_max ? -HUGE_VAL : 0.0;
    for (int k = 0; k < n_chunks; ++k) a = is_max ? fmax(a, pp[which * n_chunks + k]) : a + pp[which * n_chunks + k];
    return a;
  };
  const int lo = c * RL_MC_CHUNK, hi = min(n_samples, lo + RL_MC_CHUNK);
  if (mode == 4) {
    if (threadIdx.x != 0 || c != 0) return;
    const double mean = combine(0, false) / n_samples;
    const double ss = combine(1, false);
    double risk = mean;
    if (theta_risk > 0.0) { const double mx = combine(2, true); risk = (mx + log(combine(3, false) / n_samples)) / theta_risk; }
    stats[3 * p] = mean; stats[3 * p + 1] = n_samples > 1 ? ss / (n_samples - 1) : 0.0; stats[3 * p + 2] = risk;
    return;
  }
  double acc = (mode == 2) ? -HUGE_VAL : 0.0;
  if (mode == 0) { for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += Jp[i]; }
  else if (mode == 1) { const double mean = combine(0, false) / n_samples; for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) { double d = Jp[i] - mean; acc += d * d; } }
  else if (mode == 2) { for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) acc = fmax(acc, theta_risk * Jp[i]); }
  else { const double mx = combine(2, true); for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += exp(theta_risk * Jp[i] - mx); }
  const double r = (mode == 2) ? block_reduce(acc, OpMax(), -HUGE_VAL, sh) : block_reduce(acc, OpAdd(), 0.0, sh);
  if (threadIdx.x == 0) pp[mode * n_chunks + c] = r;
}

void launch_mc_stats(const double* J, int n_samples, int P, double theta_risk, double* stats, double* scratch, cudaStream_t st) {
  const int n_chunks = (n_samples + RL_MC_CHUNK - 1) / RL_MC_CHUNK;
  if (n_chunks < 4 || !scratch) { k_mc_stats<<<P, 256, 0, st>>>(J, n_samples, theta_risk, stats); return; }
  const dim3 grid(n_chunks, P);
  for (int mode = 0; mode < 4; ++mode) {
    if (mode >= 2 && !(theta_risk > 0.0)) continue;
    k_mc_stats_chunked<<<grid, 256, 0, st>>>(J, n_samples, n_chunks, theta_risk, mode, scratch, stats);
  }
  k_mc_stats_chunked<<<dim3(1, P), 32, 0, st>>>(J, n_samples, n_chunks, theta_risk, 4, scratch, stats);
}
size_t mc_stats_scratch_doubles(int n_samples, int P) {
  const int n_chunks = (n_samples + RL_MC_CHUNK - 1) / RL_MC_CHUNK;
  return n_chunks < 4 ? 0 : (size_t)P * 4 * n_chunks;
}

int launch_pets_costs(const PetsArgs& a, cudaStream_t st) {
  int th = a.particles >= 256 ? 256 : ((a.particles + 31) / 32) * 32;
#define X(MID, CID) if (a.model_id == MID && a.cost_id == CID) { k_pets_costs<Dyn<MID>, Cost<CID, Dyn<MID>::n, Dyn<MID>::m>><<<a.C, th, 0, st>>>(a); return 0; }
  RL_FOR_EACH_ILEQG_COMBO(X)
  RL_FOR_EACH_ROLLOUT_ONLY_COMBO(X)
#undef X
  return -1;
}

// u = mu_t + chol_lower(Sigma_t) z  (rand(rng, MvNormal(mu_t, Sigma_t)), pets.jl:212-213); thread = (tt, ii)
__global__ void k_pets_sample(int m, int N, int C, const double* mu, const double* Sigma, const double* z,
                              uint64_t seed, uint64_t stream_offset, double* controls, int32_t* err) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * C) return;
  int ii = t / N, tt = t % N;
  double Cs[16], zz[5];
  const double* Sg = Sigma + (size_t)tt * m * m;
  // runtime-m Cholesky (m <= 4)
  bool ok = true;
  for (int j = 0; j < m && ok; ++j) {
    double d = Sg[j + j * m];
    for (int k = 0; k < j; ++k) d = fma(-Cs[j + k * m], Cs[j + k * m], d);
    if (!(d > 0.0)) { ok = false; break; }
    double cjj = sqrt(d), inv = 1.0 / cjj;
    Cs[j + j * m] = cjj;
    for (int i = j + 1; i < m; ++i) {
      double a = Sg[j + i * m];
      for (int k = 0; k < j; ++k) a = fma(-Cs[i + k * m], Cs[j + k * m], a);
      Cs[i + j * m] = a * inv;
    }
  }
  if (!ok) { atomicExch(err, 1); return; }
  if (z) { for (int j = 0; j < m; ++j) zz[j] = z[(size_t)ii * m * N + (size_t)tt * m + j]; }
  else { for (int j = 0; j < m; j += 2) rl::philox_normal2(seed, stream_offset + (uint64_t)ii, (uint32_t)tt, (uint32_t)(j >> 1), &zz[j], &zz[j + 1]); }
  for (int j = 0; j < m; ++j) {
    double a = mu[(size_t)tt * m + j];
    for (int k = 0; k <= j; ++k) a = fma(Cs[j + k * m], zz[k], a);
    controls[(size_t)ii * m * N + (size_t)tt * m + j] = a;
  }
}

void launch_pets_sample(int m, int N, int C, const double* mu, const double* Sigma, const double* z, uint64_t seed,
                        uint64_t stream_offset, double* controls, int32_t* err, cudaStream_t st) {
  k_pets_sample<<<RL_BLOCKS(N * C, 128), 128, 0, st>>>(m, N, C, mu, Sigma, z, seed, stream_offset, controls, err);
}

// get_elite_samples (pets.jl:159-171): stable ascending sort by cost == total order on (cost, index),
// NaN last.  Rank-by-counting in one block (C <= 65536): rank_i = #{j : key_j < key_i}; the first
// num_elite ranks are the elites in sorted order.  O(C^2 / threads) compares, C = 4096 => 16.7M.
__device__ __forceinline__ bool key_less(double ca, int ia, double cb, int ib) {
  bool na = ca != ca, nb = cb != cb;
  if (na || nb) { if (na && nb) return ia < ib; return nb; }
  if (ca < cb) return true;
  if (cb < ca) return false;
  return ia < ib;
}

// rank by counting on the total order (cost, index) = the position a stable sort would give; one WARP per candidate, lanes
// stride over the other candidates (integer counts: exact whatever the reduction shape)
__global__ void __launch_bounds__(256) k_pets_rank(int C, int num_elite, const double* cost, int32_t* elite_idx) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= C) return;
  const double ci = cost[i];
  int rank = 0;
  for (int j = lane; j < C; j += 32) rank += key_less(cost[j], j, ci, i) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
  if (lane == 0 && rank < num_elite) elite_idx[rank] = i;
}

// compute_new_distribution (pets.jl:173-191): one warp per (tt, j).  The lanes gather the elites' controls into shared
// memory; lane 0 then sums them in sorted order (the same sequence of additions as a serial loop over the elites)
__global__ void __launch_bounds__(32) k_pets_refit(int m, int N, int num_elite, double smoothing, const double* controls,
                                                   const int32_t* elite_idx, double* mu, double* Sigma) {
  extern __shared__ double elite_vals[];
  const int t = blockIdx.x, lane = threadIdx.x;
  if (t >= N * m) return;
  const int tt = t / m, j = t % m;
  for (int e = lane; e < num_elite; e += 32) elite_vals[e] = controls[(size_t)elite_idx[e] * m * N + (size_t)tt * m + j];
  __syncwarp();
  if (lane != 0) return;
  double sum = 0.0;
  for (int e = 0; e < num_elite; ++e) sum += elite_vals[e];
  double mean = sum / num_elite;
  double ss = 0.0;
  for (int e = 0; e < num_elite; ++e) {
    double d = elite_vals[e] - mean;
    ss += d * d;
  }
  double var = ss / (num_elite - 1);  // Julia var(): unbiased
  mu[(size_t)tt * m + j] = (1.0 - smoothing) * mean + smoothing * mu[(size_t)tt * m + j];
  double* Sg = Sigma + (size_t)tt * m * m;
  for (int i = 0; i < m; ++i) {
    double cv = (i == j) ? var : 0.0;
    Sg[i + j * m] = (1.0 - smoothing) * cv + smoothing * Sg[i + j * m];
  }
}

void launch_pets_refit(int m, int N, int C, int num_elite, double smoothing, const double* controls,
                       const double* cost, double* mu, double* Sigma, int32_t* elite_idx, int32_t*,
                       cudaStream_t st) {
  k_pets_rank<<<(unsigned)(((size_t)C * 32 + 255) / 256), 256, 0, st>>>(C, num_elite, cost, elite_idx);
  const size_t smem = (size_t)num_elite * sizeof(double);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_pets_refit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_pets_refit<<<N * m, 32, smem, st>>>(m, N, num_elite, smoothing, controls, elite_idx, mu, Sigma);
}


// ---- RAT iLQR fleet kernels: thread = problem (the per-problem logic is sequential and tiny) ---------------
// get_positive_samples (cross_entropy_bilevel_optimization.jl:233-246): sequential rejection sampling
__global__ void k_ce_draw(CeFleet c) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= c.P || !c.active[p]) return;
  const bool first = c.iter[p] == 1;                       // :266-279
  const double mm = first ? c.mu_init[p] : c.mu[p], ss = first ? c.sigma_init[p] : c.sigma[p];
  long long cur = c.cursor[p];
  int cnt = 0;
  long long guard = 0;
  while (cnt < c.S) {
    if ((c.z && cur >= c.nz) || ++guard > 1000000) { c.err[p] = 1; c.active[p] = 0; break; }
    double z0, z1;
    if (c.z) z0 = c.z[(size_t)p * c.nz + cur];
    else rl::philox_normal2(c.seed, (uint64_t)(c.p0 + p), (uint32_t)cur, (uint32_t)(cur >> 32), &z0, &z1);
    cur++;
    double t = mm + ss * z0;                                // rand(rng, Normal(mu, sigma))
    if (t > 0.0) c.theta[(size_t)p * c.S + cnt++] = t;
  }
  c.cursor[p] = cur;
}

// feasibility logic, theta_min/max bookkeeping, stable sort, elite refit (step! :291-334)
__global__ void k_ce_update(CeFleet c) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= c.P || !c.active[p]) return;
  const int S = c.S;
  const double* th = c.theta + (size_t)p * S;
  const double* val = c.value + (size_t)p * S;
  const int32_t* st = c.status + (size_t)p * S;
  auto cost = [&](int i) { return st[i] == 0 ? val[i] + c.kl / th[i] : HUGE_VAL; };  // :193, exceptions -> Inf
  int num_inf = 0;
  for (int i = 0; i < S; ++i) { double ci = cost(i); num_inf += (ci == HUGE_VAL || ci == -HUGE_VAL) ? 1 : 0; }
  const int num_valid = S - num_inf;
  const double thr = fmax((double)c.num_elite, S * c.lambda);
  const bool first = c.iter[p] == 1;
  if (first && num_valid < thr) { c.mu_init[p] *= c.lambda; c.sigma_init[p] *= c.lambda; atomicAdd(c.n_active, 1); return; }  // :293-298 redraw
  else if (first && num_valid == S) { c.mu_init[p] /= c.lambda; c.sigma_init[p] /= c.lambda; }                                 // :299-305
  else if (!(num_valid >= thr)) { atomicAdd(c.n_active, 1); return; }                                                         // redraw, unchanged mu/sigma
  double tmin = c.theta_min[p], tmax = c.theta_max[p];
  for (int i = 0; i < S; ++i) {  // :314-324 (if / elseif quirk kept)
    double ci = cost(i);
    if (ci == HUGE_VAL || ci == -HUGE_VAL) continue;
    if (th[i] < tmin) tmin = th[i];
    else if (th[i] > tmax) tmax = th[i];
  }
  c.theta_min[p] = tmin; c.theta_max[p] = tmax;
  // elites = first num_elite of the stable ascending sort by cost (NaN last): selection by rank
  double sum = 0.0;
  for (int e = 0; e < c.num_elite; ++e) {
    for (int i = 0; i < S; ++i) {
      double ci = cost(i);
      int rank = 0;
      for (int j = 0; j < S; ++j) rank += key_less(cost(j), j, ci, i) ? 1 : 0;
      if (rank == e) { sum += th[i]; break; }
    }
  }
  const double mu_new = sum / c.num_elite;  // :329
  double ss = 0.0;
  for (int e = 0; e < c.num_elite; ++e) {
    for (int i = 0; i < S; ++i) {
      double ci = cost(i);
      int rank = 0;
      for (int j = 0; j < S; ++j) rank += key_less(cost(j), j, ci, i) ? 1 : 0;
      if (rank == e) { ss += (th[i] - mu_new) * (th[i] - mu_new); break; }
    }
  }
  c.mu[p] = mu_new;
  c.sigma[p] = sqrt(ss / c.num_elite);  // :330 population std
  if (c.iter[p] >= c.iter_max) c.active[p] = 0;   // CE finished for this problem
  else { c.iter[p] += 1; atomicAdd(c.n_active, 1); }
}

// ---- the same two steps for a LARGE theta population (configs[1]: one problem x 1024 theta): one CTA per problem ------
// k_ce_draw_cta: candidate t_i = mu + sigma z_i is a pure function of the stream position i, so a chunk of candidates is
// evaluated in parallel and the positive ones are compacted in stream order by a block-wide prefix sum -- exactly the
// samples, in the order, that the sequential rejection loop of get_positive_samples (:233-246) keeps.
constexpr int CE_CTA = 256;
__global__ void __launch_bounds__(CE_CTA) k_ce_draw_cta(CeFleet c) {
  const int p = blockIdx.x;
  if (!c.active[p]) return;
  __shared__ int wsum[CE_CTA / 32];
  __shared__ int s_cnt, s_stop;
  __shared__ long long s_cur;
  const bool first = c.iter[p] == 1;
  const double mm = first ? c.mu_init[p] : c.mu[p], ss = first ? c.sigma_init[p] : c.sigma[p];
  if (threadIdx.x == 0) { s_cnt = 0; s_cur = c.cursor[p]; s_stop = 0; }
  __syncthreads();
  long long guard = 0;
  while (true) {
    const long long base = s_cur;
    const int cnt0 = s_cnt;
    const long long i = base + threadIdx.x;
    bool in_range = !(c.z && i >= c.nz);
    double t = -1.0;
    if (in_range) {
      double z0, z1;
      if (c.z) z0 = c.z[(size_t)p * c.nz + i];
      else rl::philox_normal2(c.seed, (uint64_t)(c.p0 + p), (uint32_t)i, (uint32_t)(i >> 32), &z0, &z1);
      t = mm + ss * z0;
    }
    const bool pos = in_range && t > 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, pos);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int before = __popc(bal & ((1u << lane) - 1));
    for (int k = 0; k < w; ++k) before += wsum[k];
    const int slot = cnt0 + before;
    if (pos && slot < c.S) c.theta[(size_t)p * c.S + slot] = t;
    if (pos && slot == c.S - 1) { s_cur = i + 1; s_stop = 1; }  // the S-th positive sample: the stream stops right after it
    __syncthreads();
    if (threadIdx.x == 0 && !s_stop) {
      int tot = 0;
      for (int k = 0; k < CE_CTA / 32; ++k) tot += wsum[k];
      s_cnt = cnt0 + tot;
      s_cur = base + CE_CTA;
      guard += CE_CTA;
      if ((c.z && s_cur >= c.nz) || guard > 1000000) { c.err[p] = 1; c.active[p] = 0; s_stop = 2; }  // stream exhausted
    }
    __syncthreads();
    if (s_stop) break;
  }
  if (threadIdx.x == 0) c.cursor[p] = s_cur;
}

// k_ce_update_cta: step! :291-334 with the O(S^2) stable rank computed by the whole CTA; the elite mean / std are summed
// by one thread in rank order, i.e. in the order the reference's sorted array is summed.
__global__ void __launch_bounds__(CE_CTA) k_ce_update_cta(CeFleet c, double* elite_ws) {
  const int p = blockIdx.x;
  if (!c.active[p]) return;
  const int S = c.S;
  const double* th = c.theta + (size_t)p * S;
  const double* val = c.value + (size_t)p * S;
  const int32_t* st = c.status + (size_t)p * S;
  double* elite = elite_ws + (size_t)p * c.num_elite;
  auto cost = [&](int i) { return st[i] == 0 ? val[i] + c.kl / th[i] : HUGE_VAL; };
  __shared__ int s_inf, s_go;
  if (threadIdx.x == 0) { s_inf = 0; s_go = 0; }
  __syncthreads();
  int ninf = 0;
  for (int i = threadIdx.x; i < S; i += CE_CTA) { double ci = cost(i); ninf += (ci == HUGE_VAL || ci == -HUGE_VAL) ? 1 : 0; }
  if (ninf) atomicAdd(&s_inf, ninf);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int num_valid = S - s_inf;
    const double thr = fmax((double)c.num_elite, S * c.lambda);
    const bool first = c.iter[p] == 1;
    if (first && num_valid < thr) { c.mu_init[p] *= c.lambda; c.sigma_init[p] *= c.lambda; atomicAdd(c.n_active, 1); }
    else if (first && num_valid == S) { c.mu_init[p] /= c.lambda; c.sigma_init[p] /= c.lambda; s_go = 1; }
    else if (!(num_valid >= thr)) atomicAdd(c.n_active, 1);
    else s_go = 1;
    if (s_go) {  // :314-324, sequential because of the if / elseif quirk
      double tmin = c.theta_min[p], tmax = c.theta_max[p];
      for (int i = 0; i < S; ++i) {
        double ci = cost(i);
        if (ci == HUGE_VAL || ci == -HUGE_VAL) continue;
        if (th[i] < tmin) tmin = th[i];
        else if (th[i] > tmax) tmax = th[i];
      }
      c.theta_min[p] = tmin; c.theta_max[p] = tmax;
    }
  }
  __syncthreads();
  if (!s_go) return;
  for (int i = threadIdx.x; i < S; i += CE_CTA) {
    const double ci = cost(i);
    int rank = 0;
    for (int j = 0; j < S; ++j) rank += key_less(cost(j), j, ci, i) ? 1 : 0;
    if (rank < c.num_elite) elite[rank] = th[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
    for (int e = 0; e < c.num_elite; ++e) sum += elite[e];
    const double mu_new = sum / c.num_elite;  // :329
    double ss = 0.0;
    for (int e = 0; e < c.num_elite; ++e) ss += (elite[e] - mu_new) * (elite[e] - mu_new);
    c.mu[p] = mu_new;
    c.sigma[p] = sqrt(ss / c.num_elite);  // :330 population std
    if (c.iter[p] >= c.iter_max) c.active[p] = 0;
    else { c.iter[p] += 1; atomicAdd(c.n_active, 1); }
  }
}

__global__ void k_ce_pick_theta(CeFleet c, double* theta_final) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= c.P) return;
  double t = c.kl > 0 ? (c.use_theta_max ? c.theta_max[p] : c.mu[p]) : 0.0;  // :374-389
  theta_final[p] = t;
  c.theta_opt[p] = t;
  c.active[p] = c.err[p] ? 0 : 1;
}

// final solve bookkeeping: success -> value (+ kl/theta); failure -> theta_opt = max(0, theta_opt - sigma), retry (:405-413)
__global__ void k_ce_final_update(CeFleet c, double* theta_final, const double* value, const int32_t* status) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= c.P || !c.active[p]) return;
  if (status[p] == 0) {
    c.value_out[p] = c.kl > 0 ? value[p] + c.kl / theta_final[p] : value[p];
    c.theta_opt[p] = theta_final[p];
    c.active[p] = 0;
  } else {
    double t = fmax(0.0, theta_final[p] - c.sigma[p]);
    if (t == theta_final[p] && t == 0.0) { c.err[p] = 2; c.active[p] = 0; c.value_out[p] = HUGE_VAL; return; }  // even iLQG failed
    theta_final[p] = t;
    atomicAdd(c.n_active, 1);
  }
}

// ---- RAT iLQR++ fleet kernels (thread = problem) ---------------------------------------------------------------
// candidate slots per problem: phase 0 -> [theta_high, theta_low, -, -, -, -]; phase 1 -> [r, e, c1, c2, s1, s2]
__global__ void k_nm_candidates(NmFleet c) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= c.P || !c.active[p]) return;
  double* th = c.theta + (size_t)p * 6;
  if (c.phase[p] == 0) {
    th[0] = c.th_high[p]; th[1] = c.th_low[p];
    for (int i = 2; i < 6; ++i) th[i] = th[1];  // unused slots repeat a cheap, valid theta
    return;
  }
  if (c.c_high[p] < c.c_low[p]) {  // step! :184-187
    double t = c.th_low[p]; c.th_low[p] = c.th_high[p]; c.th_high[p] = t;
    t = c.c_low[p]; c.c_low[p] = c.c_high[p]; c.c_high[p] = t;
  }
  const double lo = c.th_low_init[p], m_ = c.th_low[p], hi = c.th_high[p];
  const double r = fmax(lo, m_ + c.alpha * (m_ - hi));          // :195-196
  th[0] = r;
  th[1] = fmax(lo, m_ + c.beta * (r - m_));                     // expansion :204-205
  th[2] = fmax(lo, m_ + c.gamma * (r - m_));                    // contraction if theta_high <- theta_r (:228-233)
  th[3] = fmax(lo, m_ + c.gamma * (hi - m_));                   // contraction if theta_high kept
  th[4] = (r + m_) / 2;                                         // shrink points :239
  th[5] = (hi + m_) / 2;
}

__global__ void k_nm_decide(NmFleet c) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= c.P || !c.active[p]) return;
  const double* th = c.theta + (size_t)p * 6;
  const double* val = c.value + (size_t)p * 6;
  const int32_t* st = c.status + (size_t)p * 6;
  auto cost = [&](int i) { return st[i] == 0 ? val[i] + c.kl / th[i] : HUGE_VAL; };  // compute_cost_worker :134-158
  if (c.phase[p] == 0) {  // solve! :283-304: find finite vertex costs, halving theta (and the persistent *_init) on Inf
    bool done = true;
    if (!c.has_c[2 * p]) {
      double ch = cost(0);
      c.evals[p] += 1;
      if (ch == HUGE_VAL) { c.th_high[p] *= c.lambda; c.th_high_init[p] *= c.lambda; done = false; }
      else { c.c_high[p] = ch; c.has_c[2 * p] = 1; }
    }
    if (!c.has_c[2 * p + 1]) {
      // the reference evaluates c_low only after c_high is finite; evaluations are pure, so using the speculative
      // result is equivalent -- but the evaluation COUNT follows the reference: count it once c_high is known
      double cl = cost(1);
      if (c.has_c[2 * p]) {
        c.evals[p] += 1;
        if (cl == HUGE_VAL) { c.th_low[p] *= c.lambda; c.th_low_init[p] *= c.lambda; done = false; }
        else { c.c_low[p] = cl; c.has_c[2 * p + 1] = 1; }
      } else done = false;
    }
    if (done) c.phase[p] = 1;
    atomicAdd(c.n_active, 1);
    return;
  }
  // step! :174-252, replayed on the speculative results
  c.iter[p] += 1;
  const double c_r = cost(0);
  c.evals[p] += 1;
  if (c_r < c.c_low[p]) {
    const double c_e = cost(1);
    c.evals[p] += 1;
    if (c_e < c_r) { c.th_high[p] = th[1]; c.c_high[p] = c_e; } else { c.th_high[p] = th[0]; c.c_high[p] = c_r; }
  } else {
    bool moved = false;
    if (c_r < c.c_high[p]) { c.th_high[p] = th[0]; c.c_high[p] = c_r; moved = true; }
    const int ic = moved ? 2 : 3;
    const double c_c = cost(ic);
    c.evals[p] += 1;
    if (c_c > c.c_high[p]) {
      const int is = moved ? 4 : 5;
      c.th_high[p] = th[is];            // (theta_high + theta_low)/2 :239
      c.c_high[p] = cost(is);
      c.evals[p] += 1;
    } else { c.th_high[p] = th[ic]; c.c_high[p] = c_c; }
  }
  const double c_mean = (c.c_low[p] + c.c_high[p]) / 2;  // :309-310
  const double d1 = c.c_high[p] - c_mean, d2 = c.c_low[p] - c_mean;
  const double stdev = sqrt(0.5 * (d1 * d1 + d2 * d2));
  if (stdev < c.eps || c.iter[p] == c.iter_max) c.active[p] = 0;
  else atomicAdd(c.n_active, 1);
}

// stage 0: theta_final <- theta_low (or 0 if kl == 0), all problems active; stage 1: collect the final solve (no retry, :334-350)
__global__ void k_nm_final(NmFleet c, double* theta_final, const double* value, const int32_t* status, int stage) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= c.P) return;
  if (stage == 0) {
    double t = c.kl > 0 ? c.th_low[p] : 0.0;  // :325 / :332
    theta_final[p] = t; c.theta_opt[p] = t; c.active[p] = 1;
  } else {
    c.value_out[p] = status[p] == 0 ? (c.kl > 0 ? value[p] + c.kl / theta_final[p] : value[p]) : HUGE_VAL;
  }
}

void launch_nm_candidates(const NmFleet& c, cudaStream_t st) { k_nm_candidates<<<RL_BLOCKS(c.P, 64), 64, 0, st>>>(c); }
void launch_nm_decide(const NmFleet& c, cudaStream_t st) { k_nm_decide<<<RL_BLOCKS(c.P, 64), 64, 0, st>>>(c); }
void launch_nm_final(const NmFleet& c, double* tf, const double* v, const int32_t* s, int stage, cudaStream_t st) {
  k_nm_final<<<RL_BLOCKS(c.P, 64), 64, 0, st>>>(c, tf, v, s, stage);
}

// thread per problem for the reference's small populations (10 theta), CTA per problem beyond 32 samples
void launch_ce_draw(const CeFleet& c, cudaStream_t st) {
  if (c.S > 32) k_ce_draw_cta<<<c.P, CE_CTA, 0, st>>>(c);
  else k_ce_draw<<<RL_BLOCKS(c.P, 64), 64, 0, st>>>(c);
}
void launch_ce_update(const CeFleet& c, double* elite_ws, cudaStream_t st) {
  if (c.S > 32) k_ce_update_cta<<<c.P, CE_CTA, 0, st>>>(c, elite_ws);
  else k_ce_update<<<RL_BLOCKS(c.P, 64), 64, 0, st>>>(c);
}
void launch_ce_pick_theta(const CeFleet& c, double* tf, cudaStream_t st) { k_ce_pick_theta<<<RL_BLOCKS(c.P, 64), 64, 0, st>>>(c, tf); }
void launch_ce_final_update(const CeFleet& c, double* tf, const double* v, const int32_t* s, cudaStream_t st) {
  k_ce_final_update<<<RL_BLOCKS(c.P, 64), 64, 0, st>>>(c, tf, v, s);
}

}  // namespace rll
