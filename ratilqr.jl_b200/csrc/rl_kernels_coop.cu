// rl_kernels_coop.cu -- the warp-cooperative solve kernel (rl_coop.cuh): one warp (= one CTA) per instance.  Own translation
// unit so that it builds beside the thread-per-instance kernels (rl_kernels_solve.cu).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rl_coop.cuh"
#include "rl_kernels_model.cuh"
#include "rl_host.hpp"
#include "rl_launch.hpp"

namespace rll {

using namespace rl;

// ---- warp-cooperative variant: one warp (= one CTA) per instance, matrices + trajectories in shared memory -----
template <class D, class CT>
__global__ void __launch_bounds__(32) k_ileqg_solve_coop(const __grid_constant__ SolveParams P, double* traj_global) {
  extern __shared__ double coop_smem[];
  constexpr int n = D::n, m = D::m;
  CoopWs<n, m>& w = *reinterpret_cast<CoopWs<n, m>*>(coop_smem);
  const size_t inst = blockIdx.x;
  const size_t td = coop_traj_doubles(n, m, P.N);
  double* base = traj_global ? traj_global + inst * td : coop_smem + (sizeof(CoopWs<n, m>) + 7) / 8;
  CoopTraj tj;
  tj.X = base; tj.U = tj.X + (size_t)2 * (P.N + 1) * n; tj.Lg = tj.U + (size_t)2 * P.N * m; tj.DL = tj.Lg + (size_t)P.N * m * n;
  int cur = 0;
  const int lane = threadIdx.x;
  if (coop_solve_instance<D, CT>(lane, P, inst, w, tj, cur)) coop_write_outputs<n, m>(lane, P, inst, tj, cur);
}

template <class D, class CT>
size_t coop_smem_bytes(int N, bool traj_in_smem) {
  size_t b = ((sizeof(CoopWs<D::n, D::m>) + 7) / 8) * 8;
  if (traj_in_smem) b += coop_traj_doubles(D::n, D::m, N) * 8;
  return b;
}

template <int MID, int CID>
static int launch_coop_one(const SolveParams& P, double* traj_global, bool query_only, size_t* smem_out, cudaStream_t st) {
  using D = Dyn<MID>;
  using CT = Cost<CID, D::n, D::m>;
  size_t smem = coop_smem_bytes<D, CT>(P.N, true);
  bool in_smem = smem <= 200 * 1024;
  if (in_smem && !query_only && traj_global) {
    // Trajectories in shared memory minimise latency, but cap residency (5 warps/SM for the quadrotor).  When the batch
    // exceeds one resident wave, keep only the per-stage matrices in shared memory and the trajectories in HBM
    // (contiguous per instance, read once per stage): ~3x more resident warps to hide the shared-memory latency chains.
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t per_sm = (227 * 1024) / (smem + 1024);
    const char* e = getenv("RATILQR_COOP_TRAJ");  // smem / global: force (tuning)
    if (e ? (e[0] == 'g') : ((size_t)P.B > per_sm * (size_t)sms)) in_smem = false;
  }
  if (!in_smem) smem = coop_smem_bytes<D, CT>(P.N, false);
  if (smem_out) *smem_out = smem;
  if (query_only) return 0;
  auto kfn = k_ileqg_solve_coop<D, CT>;
  static size_t configured_dev[64] = {0};  // per device (function attributes do not carry over to another GPU)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  size_t& configured = configured_dev[cur_dev & 63];
  if (smem > configured) {
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = smem;
  }
  kfn<<<P.B, 32, smem, st>>>(P, in_smem ? nullptr : traj_global);
  return 0;
}

// coop kernels exist for every pair the serial kernel supports (the quadrotor also gets the diagonal-cost variant)
#define RL_FOR_EACH_COOP_COMBO(X) RL_FOR_EACH_ILEQG_COMBO(X) RL_FOR_EACH_DIAG_COMBO(X) X(RATILQR_MODEL_QUADROTOR, RL_COST_QUAD_DIAG)

int coop_smem_query(int model_id, int cost_id, int N, size_t* smem) {
  SolveParams P; memset(&P, 0, sizeof(P)); P.N = N;
#define X(MID, CID) if (model_id == MID && cost_id == CID) return launch_coop_one<MID, CID>(P, nullptr, true, smem, 0);
  RL_FOR_EACH_COOP_COMBO(X)
#undef X
  return -1;
}

int launch_solve_coop(int model_id, int cost_id, const SolveParams& P, double* traj_global, cudaStream_t st) {
#define X(MID, CID) if (model_id == MID && cost_id == CID) return launch_coop_one<MID, CID>(P, traj_global, false, nullptr, st);
  RL_FOR_EACH_COOP_COMBO(X)
#undef X
  return -1;
}

}  // namespace rll
