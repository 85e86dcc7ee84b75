// rl_kernels_model.cuh -- the __global__ kernels that are templates over the (dynamics, cost) pair.
//
// They live in a header because two compilers instantiate them: nvcc, for the registered models of the
// library build (rl_kernels_solve.cu, rl_kernels_comp.cu), and NVRTC at run time, for user-supplied
// dynamics / cost snippets (rl_user.cuh, ratilqr_user_model_register).  Every kernel takes ONE by-value
// argument block (rl_args.hpp / rl::SolveParams), so a driver-API launch passes a single pointer.
#pragma once
#include "rl_args.hpp"
#include "rl_components.cuh"

namespace rll {

using namespace rl;

// ---- the hot kernel: one persistent thread per iLEQG instance (solve!, ileqg.jl:635-659) ------------------------
// WC: W(k) is constant and small: inv(W) is read from the argument block (constant bank), see SolveParams::Winvc
template <class D, class CT, int THREADS, int MINB, bool WC = false>
__global__ void __launch_bounds__(THREADS, MINB) k_ileqg_solve(const __grid_constant__ SolveParams P) {
  extern __shared__ double stage_area[];  // [2][RL_STAGE_NV][THREADS] doubles when staging is enabled, else empty
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Stage sg;
  sg.base = (UseStage<D>::value && P.use_stage) ? stage_area + threadIdx.x : nullptr;
  sg.stride = THREADS;
#if defined(RL_ENABLE_DYNAMIC)  // opt-in build of a negative result (profiles/r01_dynamic_refill_ab.jsonl): persistent threads
  if (!WC && P.queue) { solve_dynamic<D, CT>(P, b, sg); return; }  // that keep pulling instances from a queue
#endif
  if (b < (size_t)P.B) {
    if (sg.base) solve_instance<D, CT, WC, 1>(P, b, sg);
    else solve_instance<D, CT, WC, 0>(P, b, sg);
  }
}

// the same kernel under an explicit register cap (__maxnreg__): launch bounds only offer the caps 65536 / (THREADS * MINB)
// rounded down by ptxas to 128 / 168 / 255 here; tuning shapes of rl_kernels_solve.cu (RL_TUNE_SHAPES)
#if defined(RL_TUNE_SHAPES)
template <class D, class CT, int THREADS, int MAXREG>
__global__ void __launch_bounds__(THREADS) __maxnreg__(MAXREG) k_ileqg_solve_r(const __grid_constant__ SolveParams P) {
  extern __shared__ double stage_area[];
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Stage sg;
  sg.base = stage_area + threadIdx.x;
  sg.stride = THREADS;
  if (b < (size_t)P.B) solve_instance<D, CT, false, 1>(P, b, sg);
}
#endif

// ---- rollouts / cost / linearize: thread = instance (host layout, instance slowest) ----------
template <class D>
__global__ void k_rollout_open(CompArgs a) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  constexpr int n = D::n, m = D::m;
  const int N = a.N;
  int st = comp_rollout_open<D>(a.mp, N, a.x0 + (size_t)b * n, a.u + (size_t)b * m * N, a.x + (size_t)b * n * (N + 1));
  if (a.status) a.status[b] = st;
}

template <class D, class CT>
__global__ void k_rollout_closed(CompArgs a) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  constexpr int n = D::n, m = D::m;
  const int N = a.N;
  int st = comp_rollout_closed<D, CT>(a.mp, a.cp, N, a.xbar + (size_t)b * n * (N + 1), a.l + (size_t)b * m * N,
                                      a.L + (size_t)b * m * n * N, nullptr, a.x + (size_t)b * n * (N + 1),
                                      a.u_new + (size_t)b * m * N, nullptr);
  if (a.status) a.status[b] = st;
}

template <class D, class CT>
__global__ void k_integrate_cost(CompArgs a) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  constexpr int n = D::n, m = D::m;
  const int N = a.N;
  double c = HUGE_VAL;
  int st = comp_integrate_cost<D, CT>(a.cp, N, a.x + (size_t)b * n * (N + 1), a.u + (size_t)b * m * N, &c);
  a.cost[b] = st ? HUGE_VAL : c;
  if (a.status) a.status[b] = st;
}

// approximate_model is embarrassingly parallel over stages (ileqg.jl:293): thread = (stage, instance)
template <class D, class CT>
__global__ void k_linearize(CompArgs a) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = a.N;
  if (t >= a.B * (N + 1)) return;
  constexpr int n = D::n, m = D::m;
  int b = t / (N + 1), k = t % (N + 1);
  int st = comp_linearize_stage<D, CT>(a.mp, a.cp, N, k, a.x + (size_t)b * n * (N + 1), a.u + (size_t)b * m * N,
                                       a.q + (size_t)b * (N + 1), a.qv + (size_t)b * n * (N + 1),
                                       a.Q + (size_t)b * n * n * (N + 1), a.r + (size_t)b * m * N, a.R + (size_t)b * m * m * N,
                                       a.Pm + (size_t)b * m * n * N, a.A + (size_t)b * n * n * N, a.Bm + (size_t)b * n * m * N);
  if (st && a.status) atomicMax(&a.status[b], st);
}

#ifndef RL_MC_MINB
#define RL_MC_MINB(n) 1
#endif
// ---- Monte Carlo closed-loop rollouts (ileqg.jl:94-109 + :115-124): thread = sample -------------
// The policy (xbar, l, L) of a problem is read by every thread of the block through the
// read-only path (same address across the warp => one broadcast transaction); the injected
// noise row of a sample is a contiguous n*N block (each 32-byte sector fully used).
template <class D, class CT>
__global__ void __launch_bounds__(128, RL_MC_MINB(D::n)) k_mc_rollout(McArgs a) {
  constexpr int n = D::n, m = D::m;
  int p = blockIdx.y;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n_samples) return;
  size_t gi = (size_t)p * a.n_samples + s;
  int N = a.N;
  const double* cp = a.cp + (a.cp_count > 1 ? (size_t)p * a.ncp : 0);
  const double* xbar = a.xbar + (size_t)p * n * (N + 1);
  const double* l = a.l + (size_t)p * m * N;
  const double* L = a.L + (size_t)p * m * n * N;
  double cost = HUGE_VAL;
  int st;
  if (a.noise) {
    st = comp_rollout_closed<D, CT>(a.mp, cp, N, xbar, l, L, a.noise + gi * n * N,
                                    a.x_out ? a.x_out + gi * n * (N + 1) : nullptr, nullptr, &cost);
  } else {
    // Philox mode: same loop with generated noise
    double x[n], xn[n], u[m], w[n];
    for (int i = 0; i < n; ++i) { x[i] = xbar[i]; if (a.x_out) a.x_out[gi * n * (N + 1) + i] = x[i]; }
    double J = 0.0;
    st = 0;
    for (int k = 0; k < N && !st; ++k) {
      double dx[n];
      for (int i = 0; i < n; ++i) dx[i] = x[i] - xbar[(size_t)k * n + i];
      const double* Lk = L + (size_t)k * m * n;
      for (int j = 0; j < m; ++j) {
        double acc = Lk[j] * dx[0];
        for (int i = 1; i < n; ++i) acc = rl_fma(Lk[j + i * m], dx[i], acc);
        u[j] = l[(size_t)k * m + j] + acc;
      }
      double q;
      if (!CT::stage(cp, k, x, u, false, q, nullptr, nullptr, nullptr, nullptr, nullptr)) { st = RATILQR_ST_DOMAIN; break; }
      J += q;
      if (!D::f(a.mp, x, u, xn)) { st = RATILQR_ST_DOMAIN; break; }
      if (a.mix.k > 0) philox_mixture_noise<n>(a.seed, gi, (uint32_t)k, a.mix, w);
      else philox_noise<n>(a.seed, gi, (uint32_t)k, 0, 1.0, a.cholW + (a.W_tv ? (size_t)k * n * n : 0), w);
      for (int i = 0; i < n; ++i) { x[i] = xn[i] + w[i]; if (a.x_out) a.x_out[gi * n * (N + 1) + (size_t)(k + 1) * n + i] = x[i]; }
    }
    if (!st) {
      double q;
      if (!CT::terminal(cp, x, false, q, nullptr, nullptr)) st = RATILQR_ST_DOMAIN; else cost = J + q;
    }
  }
  a.J[gi] = st ? HUGE_VAL : cost;
}

// fixed-shape block reductions (deterministic): one block per problem
template <class Op>
__device__ double block_reduce(double v, Op op, double neutral, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_down_sync(0xffffffffu, v, o));
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  if (w == 0) {
    double acc = (lane < nw) ? sh[lane] : neutral;
    for (int o = 16; o > 0; o >>= 1) acc = op(acc, __shfl_down_sync(0xffffffffu, acc, o));
    if (lane == 0) sh[32] = acc;
  }
  __syncthreads();
  double r = sh[32];
  __syncthreads();
  return r;
}
struct OpAdd { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMax { __device__ double operator()(double a, double b) const { return fmax(a, b); } };

// register cap of the PETS rollout kernel: small systems do not need 164 registers; 2-3 resident 256-thread CTAs per SM
// hide the latency of the FP64 transcendental chains (sincos, log, sqrt of the dynamics and of Box-Muller)
#ifndef RL_PETS_MINB
#define RL_PETS_MINB(n) ((n) <= 4 ? 3 : 1)
#endif
// ---- PETS: compute_cost_serial (pets.jl:128-157): block = sequence, thread = particle -------------
template <class D, class CT>
__global__ void __launch_bounds__(256, RL_PETS_MINB(D::n)) k_pets_costs(PetsArgs a) {
  constexpr int n = D::n, m = D::m;
  __shared__ double sh[33];
  int ii = blockIdx.x;
  double acc = 0.0;
  int per = a.n_ens > 1 ? max(a.particles / a.n_ens, 1) : a.particles;
  for (int kk = threadIdx.x; kk < a.particles; kk += blockDim.x) {
    const double* mpp = a.mp;
    if (a.ens_params && a.n_ens > 1) mpp = a.ens_params + (size_t)min(kk / per, a.n_ens - 1) * a.n_mp;
    size_t gi = (size_t)ii * a.particles + kk;
    double c = comp_pets_particle<D, CT>(mpp, a.cp, a.N, a.x0, a.controls + (size_t)ii * m * a.N,
                                         a.noise ? a.noise + gi * n * a.N : nullptr, a.seed, a.stream_offset + gi,
                                         a.noise_kind, a.noise_scale, a.cholW, &a.mix);
    acc += c;
  }
  double tot = block_reduce(acc, OpAdd(), 0.0, sh);
  if (threadIdx.x == 0) a.cost[ii] = tot / a.particles;  // mean over particles (pets.jl:154)
}

}  // namespace rll
