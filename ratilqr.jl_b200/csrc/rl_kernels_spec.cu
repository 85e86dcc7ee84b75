// rl_kernels_spec.cu -- the latency kernel for small batches: G lanes per iLEQG instance used as speculative workers
// (trial speculation x pass speculation, see rl_spec.cuh).  One warp per CTA, 32 / G instances per warp.
#include <cstdio>
#include <cstdlib>

#include "rl_coop2.cuh"
#include "rl_spec.cuh"
#include "rl_host.hpp"
#include "rl_launch.hpp"

namespace rll {

using namespace rl;

template <class D, class CT, int G>
__global__ void __launch_bounds__(32) k_ileqg_solve_spec(const __grid_constant__ SolveParams P) {
  extern __shared__ double stage_area[];  // [2][RL_STAGE_NV][32]
  constexpr int n = D::n, m = D::m;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x;
  const int g = lane % G;
  const size_t slot = ((size_t)blockIdx.x * 32 + lane) / G;  // thread slot of the instance; workspace columns are per LANE
  bool alive = slot < (size_t)P.B;
  const size_t inst = alive ? (P.perm ? (size_t)P.perm[slot] : slot) : 0;
  const size_t p = inst / (size_t)P.K;
  if (alive && P.active && !P.active[p]) alive = false;
  const double* cp = P.cost_params + (P.cp_count > 1 ? p * (size_t)P.ncp : 0);
  const double theta = alive ? P.theta[inst] : 0.0;
  Stage sg;
  sg.base = (UseStage<D>::value && P.use_stage) ? stage_area + lane : nullptr;
  sg.stride = 32;
  SpecCols C;
  C.P = &P; C.tile = blockIdx.x; C.lane0 = (size_t)(lane - g); C.n = n; C.m = m; C.N = P.N;
  SpecState S;
  spec_state_init(P, S);
  while (true) {
    const bool work = alive && !S.done;
    SpecLaneRes r;
    r.st_roll = 0; r.rc = 0; r.nw = 0.0; r.dmax = 0.0; r.mu = 0.0; r.delta = 0.0; r.restarts = 0;
    if (work) r = sg.base ? spec_lane_work<D, CT, 1>(P, C, g, S, cp, theta, p, sg) : spec_lane_work<D, CT, 0>(P, C, g, S, cp, theta, p, sg);
    __syncwarp(full);  // candidates / speculative policies written above are read by the other lanes of the group below
    SpecLaneRes res[G];
#pragma unroll
    for (int s = 0; s < G; ++s) {
      res[s].st_roll = __shfl_sync(full, r.st_roll, s, G);
      res[s].rc = __shfl_sync(full, r.rc, s, G);
      res[s].nw = __shfl_sync(full, r.nw, s, G);
      res[s].dmax = __shfl_sync(full, r.dmax, s, G);
      res[s].mu = __shfl_sync(full, r.mu, s, G);
      res[s].delta = __shfl_sync(full, r.delta, s, G);
      res[s].restarts = __shfl_sync(full, r.restarts, s, G);
    }
    if (work) spec_decide<G>(P, S, res, inst, g == 0);
    if (!__any_sync(full, alive && !S.done)) break;
  }
  if (alive) spec_write_outputs(P, C, S, inst, g, G);
}

template <int MID, int CID, int G>
static void launch_spec_one(const SolveParams& P, cudaStream_t st) {
  using D = Dyn<MID>;
  using CT = Cost<CID, D::n, D::m>;
  const size_t smem = UseStage<D>::value ? (size_t)2 * RL_STAGE_NV * 32 * sizeof(double) : 0;
  const unsigned blocks = (unsigned)(((size_t)P.B * G + 31) / 32);
  k_ileqg_solve_spec<D, CT, G><<<blocks, 32, smem, st>>>(P);
}

// pairs served by the speculative kernel: per-thread matrices must fit the register file (n <= 6)
#define RL_FOR_EACH_SPEC_COMBO(X)                                    \
  X(RATILQR_MODEL_SINGLE_INTEGRATOR, RATILQR_COST_QUADRATIC)         \
  X(RATILQR_MODEL_POWER_LAW, RATILQR_COST_POWER_LAW)                 \
  X(RATILQR_MODEL_POWER_LAW, RATILQR_COST_QUADRATIC)                 \
  X(RATILQR_MODEL_DOUBLE_INTEGRATOR, RATILQR_COST_QUADRATIC)         \
  X(RATILQR_MODEL_PENDULUM, RATILQR_COST_QUADRATIC)                  \
  X(RATILQR_MODEL_CARTPOLE, RATILQR_COST_QUADRATIC)                  \
  X(RATILQR_MODEL_UNICYCLE, RATILQR_COST_QUADRATIC)                  \
  RL_FOR_EACH_DIAG_COMBO(X)

bool spec_supported(int model_id, int cost_id) {
#define X(MID, CID) if (model_id == MID && cost_id == CID) return true;
  RL_FOR_EACH_SPEC_COMBO(X)
#undef X
  return false;
}

int launch_solve_spec(int model_id, int cost_id, int G, const SolveParams& P, cudaStream_t st) {
#define X(MID, CID)                                                        \
  if (model_id == MID && cost_id == CID) {                                 \
    if (G == 8) { launch_spec_one<MID, CID, 8>(P, st); return 0; }         \
    if (G == 4) { launch_spec_one<MID, CID, 4>(P, st); return 0; }         \
    if (G == 2) { launch_spec_one<MID, CID, 2>(P, st); return 0; }         \
    return -1;                                                             \
  }
  RL_FOR_EACH_SPEC_COMBO(X)
#undef X
  return -1;
}

// ---- two-warp speculative variant of the warp-cooperative kernel (rl_coop2.cuh): one CTA of 64 threads per instance ----
template <class D, class CT>
__global__ void __launch_bounds__(64) k_ileqg_solve_coop2(const __grid_constant__ SolveParams P) {
  extern __shared__ double coop2_smem[];
  __shared__ SpecLaneRes res[2];
  constexpr int n = D::n, m = D::m;
  const int tid = threadIdx.x, g = tid >> 5, lane = tid & 31;
  constexpr size_t wsd = (sizeof(CoopWs<n, m>) + 7) / 8;
  CoopWs<n, m>& w = *reinterpret_cast<CoopWs<n, m>*>(coop2_smem + (size_t)g * wsd);
  Coop2Traj t;
  t.n = n; t.m = m; t.N = P.N;
  t.X = coop2_smem + 2 * wsd; t.U = t.X + (size_t)3 * (P.N + 1) * n; t.Lg = t.U + (size_t)3 * P.N * m; t.DL = t.Lg + (size_t)2 * P.N * m * n;
  const size_t inst = blockIdx.x;
  const size_t p = inst / (size_t)P.K;
  if (P.active && !P.active[p]) return;
  const double* cp = P.cost_params + (P.cp_count > 1 ? p * (size_t)P.ncp : 0);
  const double theta = P.theta[inst];
  SpecState S;
  spec_state_init(P, S);
  coop_ws_init(lane, w);
  while (!S.done) {
    const SpecLaneRes r = coop2_warp_work<D, CT>(lane, g, P, S, cp, theta, p, w, t);
    if (lane == 0) res[g] = r;
    __syncthreads();
    SpecLaneRes rr[2];
    rr[0] = res[0]; rr[1] = res[1];
    spec_decide<2>(P, S, rr, inst, tid == 0);
    __syncthreads();
  }
  coop2_write_outputs(P, t, S, inst, tid, 64);
}

template <class D, class CT>
static size_t coop2_smem_bytes(int N) {
  return (2 * ((sizeof(CoopWs<D::n, D::m>) + 7) / 8) + coop2_traj_doubles(D::n, D::m, N)) * sizeof(double);
}

#define RL_FOR_EACH_COOP2_COMBO(X) X(RATILQR_MODEL_QUADROTOR, RATILQR_COST_QUADRATIC) X(RATILQR_MODEL_QUADROTOR, RL_COST_QUAD_DIAG)

// bytes of dynamic shared memory one instance needs, or 0 if the pair is not served by this kernel
size_t coop2_smem_query(int model_id, int cost_id, int N) {
#define X(MID, CID) if (model_id == MID && cost_id == CID) return coop2_smem_bytes<Dyn<MID>, Cost<CID, Dyn<MID>::n, Dyn<MID>::m>>(N);
  RL_FOR_EACH_COOP2_COMBO(X)
#undef X
  return 0;
}

int launch_solve_coop2(int model_id, int cost_id, const SolveParams& P, cudaStream_t st) {
#define X(MID, CID)                                                                                      \
  if (model_id == MID && cost_id == CID) {                                                               \
    using D = Dyn<MID>;                                                                                  \
    using CT = Cost<CID, D::n, D::m>;                                                                    \
    const size_t smem = coop2_smem_bytes<D, CT>(P.N);                                                    \
    auto kfn = k_ileqg_solve_coop2<D, CT>;                                                               \
    static size_t configured_dev[64] = {0};                                                              \
    int dev = 0;                                                                                         \
    cudaGetDevice(&dev);                                                                                 \
    if (smem > configured_dev[dev & 63]) {                                                               \
      cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
      cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);                    \
      configured_dev[dev & 63] = smem;                                                                   \
    }                                                                                                    \
    kfn<<<P.B, 64, smem, st>>>(P);                                                                       \
    return 0;                                                                                            \
  }
  RL_FOR_EACH_COOP2_COMBO(X)
#undef X
  return -1;
}

}  // namespace rll
