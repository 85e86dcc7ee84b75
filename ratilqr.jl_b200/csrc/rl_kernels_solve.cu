// rl_kernels_solve.cu -- the hot kernel: one persistent thread per iLEQG instance.
//
// k_ileqg_solve runs solve!(::ILEQGSolver) (ileqg.jl:635-659) for instance b entirely on the
// device: open-loop rollout -> evaluation pass -> { optimising Riccati pass (fused with the
// linearisation) -> line-search trials (closed-loop rollout + evaluation pass) } until the stop
// rule fires.  No host round trips, no tensor cores (tiny FP64 matrices), per-stage matrices in
// registers, trajectories in an SoA workspace (instance index fastest => every load/store of a
// warp is one coalesced 256-byte transaction).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "rl_kernels_model.cuh"
#include "rl_host.hpp"
#include "rl_launch.hpp"

namespace rll {

using namespace rl;

template <class D, class CT, int THREADS, int MINB, bool WC = false>
static void launch_shape(const SolveParams& P, cudaStream_t st) {
  int blocks = (P.B + THREADS - 1) / THREADS;
  const size_t smem = UseStage<D>::value ? (size_t)2 * RL_STAGE_NV * THREADS * sizeof(double) : 0;
  static std::mutex cfg_mutex;  // sub-fleets launch from several host threads (ratilqr_ce_solve_fleet)
  static bool configured_dev[64] = {false};  // function attributes are per device (ratilqr_create_multi: several in one process)
  static int resident = MINB, sms = 148;
  std::lock_guard<std::mutex> cfg_lock(cfg_mutex);
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  if (!configured) {
    configured = true;
    auto kfn = k_ileqg_solve<D, CT, THREADS, MINB, WC>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (smem > 0) {
      // the default L1/shared split admits only a few staging CTAs per SM: ask for what MINB resident CTAs need
      int pct = (int)((MINB * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 5;
      cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kfn, THREADS, smem);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (getenv("RATILQR_DEBUG")) {
      int nb = resident;
      fprintf(stderr, "[ratilqr] k_ileqg_solve threads=%d minb=%d smem=%zu -> %d resident CTAs/SM\n", THREADS, MINB, smem, nb);
    }
  }
  if (P.queue && blocks > resident * sms) blocks = resident * sms;  // persistent grid: exactly one resident wave
  k_ileqg_solve<D, CT, THREADS, MINB, WC><<<blocks, THREADS, smem, st>>>(P);
}

#if defined(RL_TUNE_SHAPES)
template <class D, class CT, int THREADS, int MAXREG>
static void launch_shape_r(const SolveParams& P, cudaStream_t st) {
  const int blocks = (P.B + THREADS - 1) / THREADS;
  const size_t smem = (size_t)2 * RL_STAGE_NV * THREADS * sizeof(double);
  constexpr int resident = 65536 / (MAXREG * THREADS);
  auto kfn = k_ileqg_solve_r<D, CT, THREADS, MAXREG>;
  static bool configured = false;
  if (!configured) {
    configured = true;
    int pct = (int)((resident * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 5;
    cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kfn, THREADS, smem);
    if (getenv("RATILQR_DEBUG")) fprintf(stderr, "[ratilqr] k_ileqg_solve_r threads=%d maxreg=%d smem=%zu -> %d resident CTAs/SM\n", THREADS, MAXREG, smem, nb);
  }
  kfn<<<blocks, THREADS, smem, st>>>(P);
}
#endif

// launch shape = (threads per CTA, min resident CTAs per SM => register cap).  The default was chosen from
// the sweep recorded in profiles/; RATILQR_SOLVE_SHAPE=<idx> overrides it for tuning runs.
static int shape_override() {
  static const int v = [] { const char* e = getenv("RATILQR_SOLVE_SHAPE"); return e ? atoi(e) : -1; }();  // thread-safe init
  return v;
}

template <int MID, int CID>
static void launch_one(const SolveParams& P, cudaStream_t st) {
  using D = Dyn<MID>;
  using CT = Cost<CID, D::n, D::m>;
  if constexpr (MID == RATILQR_MODEL_UNICYCLE && CID == RL_COST_QUAD_DIAG) {
    switch (shape_override()) {
      // (the other shapes of the round-1 sweeps -- 64x{5,7,8}, 32x{8..20} -- lost everywhere and were removed: profiles/r01_tune_*.jsonl)
      case 0: launch_shape<D, CT, 64, 4>(P, st); return;    // 255 regs,  8 warps/SM
      case 11: launch_shape<D, CT, 128, 3>(P, st); return;  // 168 regs, 128-thread CTAs
#if defined(RL_TUNE_SHAPES)
      case 20: launch_shape_r<D, CT, 64, 144>(P, st); return;   // 14 warps/SM
      case 21: launch_shape_r<D, CT, 32, 152>(P, st); return;   // 13 warps/SM
      case 22: launch_shape_r<D, CT, 96, 136>(P, st); return;   // 15 warps/SM
      case 23: launch_shape_r<D, CT, 64, 160>(P, st); return;   // 12 warps/SM, 64-thread CTAs
      case 24: launch_shape<D, CT, 128, 4>(P, st); return;      // 128 regs, 16 warps/SM
      case 25: launch_shape<D, CT, 64, 8>(P, st); return;       // 128 regs, 16 warps/SM, 64-thread CTAs
#endif
      default: {
        // Throughput shape (168 registers, 12 warps/SM) once the batch exceeds what it keeps resident at a time; below
        // that every instance is resident anyway and the 255-register build wins on latency (fewer spills, more ILP per
        // thread): 7.3 vs 8.2 ms for 10-1024 solves, 39 vs 43 ms for a 56,830-instance CE round, equal at 82k
        // (profiles/r01_shape_vs_batch.jsonl).
        static const int sms = [] { int d = 0, v = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
        // inv(W) from the constant bank instead of L1: measured 0.8 % SLOWER at full load (4.67 vs 4.71 M solves/s, same box,
        // two alternating runs each: profiles/r02_wconst_ab.txt) and equal in the latency regime, so it stays opt-in
        // (those kernels are only compiled with -DRL_ENABLE_WCONST)
#if defined(RL_ENABLE_WCONST)
        static const bool wc_on = [] { const char* e = getenv("RATILQR_WCONST"); return e && e[0] == '1'; }();
        const bool wc = wc_on && P.w_const && !P.queue;
        if (wc) {
          if ((size_t)P.B <= (size_t)sms * 384) launch_shape<D, Cost<RL_COST_QUAD_DIAG_PIN, D::n, D::m>, 64, 4, true>(P, st);
          else launch_shape<D, CT, 128, 3, true>(P, st);
          return;
        }
#endif
        if ((size_t)P.B <= (size_t)sms * 384 && !P.queue) launch_shape<D, Cost<RL_COST_QUAD_DIAG_PIN, D::n, D::m>, 64, 4>(P, st);  // + constants pinned in L1
#if defined(RL_PIN_THROUGHPUT)  // tuning build: the throughput shape with the cost constants on the L1 evict-last path too
        else launch_shape<D, Cost<RL_COST_QUAD_DIAG_PIN, D::n, D::m>, 128, 3>(P, st);
#else
        else launch_shape<D, CT, 128, 3>(P, st);   // best of the sweeps in profiles/r01_tune_*.jsonl
#endif
        return;
      }
    }
  } else {
    launch_shape<D, CT, 64, 4>(P, st);
  }
}

int launch_solve(int model_id, int cost_id, const SolveParams& P, cudaStream_t st) {
#define X(MID, CID) if (model_id == MID && cost_id == CID) { launch_one<MID, CID>(P, st); return 0; }
  RL_FOR_EACH_ILEQG_COMBO(X)
  RL_FOR_EACH_DIAG_COMBO(X)
#undef X
  return -1;
}

// ---- outputs: warp-tiled SoA workspace -> host layout -------------------------------------------
// src element e of the instance in slot b: src[((b/32) * Etot + buf_b * E + e) * 32 + b%32]  (buf_b = cur[b] for the
// double-buffered arrays, Etot = elements per slot of the whole workspace record).  dst: dst[instance * E + e].
// 32x32 tiles through shared memory: coalesced on both sides.
__global__ void k_gather(const double* __restrict__ src, const int32_t* __restrict__ cur,
                         const int32_t* __restrict__ perm, int E, size_t Etot, int B, double* __restrict__ dst) {
  __shared__ double tile[32][33];
  int b0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int e = e0 + r, b = b0 + threadIdx.x;
    if (e < E && b < B) {
      size_t buf = cur ? (size_t)cur[b] : 0;
      tile[r][threadIdx.x] = src[((size_t)blockIdx.x * Etot + buf * (size_t)E + e) * 32 + threadIdx.x];
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int b = b0 + r, e = e0 + threadIdx.x;
    if (e < E && b < B) dst[(size_t)(perm ? perm[b] : b) * E + e] = tile[threadIdx.x][r];
  }
}

void launch_gather(int n, int m, int N, int B, const double* X, const double* U, const double* Lg, size_t rec,
                   const int32_t* cur, const int32_t* perm, double* x_out, double* l_out, double* L_out, cudaStream_t st) {
  dim3 th(32, 8);
  if (x_out) { int E = n * (N + 1); k_gather<<<dim3((B + 31) / 32, (E + 31) / 32), th, 0, st>>>(X, cur, perm, E, rec, B, x_out); }
  if (l_out) { int E = m * N; k_gather<<<dim3((B + 31) / 32, (E + 31) / 32), th, 0, st>>>(U, cur, perm, E, rec, B, l_out); }
  if (L_out) { int E = m * n * N; k_gather<<<dim3((B + 31) / 32, (E + 31) / 32), th, 0, st>>>(Lg, nullptr, perm, E, rec, B, L_out); }
}

// ---- theta sort: one CTA per problem, bitonic sort of (theta, index) in shared memory ---------------
// perm[r*K + j] = p*K + (index of the j-th smallest theta of problem p); ties broken by index (total order).
// p = order[r] when a problem order is given (heaviest problems first, see rl_capi.cu), else p = r.
__global__ void __launch_bounds__(1024) k_sort_theta(const double* __restrict__ theta, int K, int Kpad,
                                                     const int32_t* __restrict__ order, int32_t* __restrict__ perm) {
  extern __shared__ unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);
  int32_t* idx = reinterpret_cast<int32_t*>(key + Kpad);
  const size_t base = (size_t)(order ? order[blockIdx.x] : (int)blockIdx.x) * K;
  const size_t slot0 = (size_t)blockIdx.x * K;
  for (int i = threadIdx.x; i < Kpad; i += blockDim.x) {
    double t = i < K ? theta[base + i] : HUGE_VAL;
    key[i] = (t != t) ? HUGE_VAL : t;  // NaN sorts last
    idx[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= Kpad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < Kpad; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          bool up = (i & k) == 0;
          double a = key[i], b = key[ixj];
          int ia = idx[i], ib = idx[ixj];
          bool a_gt_b = (a > b) || (a == b && ia > ib);
          if (a_gt_b == up) { key[i] = b; key[ixj] = a; idx[i] = ib; idx[ixj] = ia; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < K; i += blockDim.x) perm[slot0 + i] = (int32_t)(base + idx[i]);
}

// slots of problem rank r <- instances of problem order[r], in instance order (K = 1, or K too large for one CTA)
__global__ void k_order_slots(const int32_t* __restrict__ order, int P, int K, int32_t* __restrict__ perm) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= (size_t)P * K) return;
  perm[s] = (int32_t)((size_t)order[s / K] * K + s % K);
}

// key[p] = max iterations over the K instances of problem p in the launch that just finished (slot-ordering profile)
__global__ void k_problem_work(const int32_t* __restrict__ iters, int P, int K, int32_t* __restrict__ key) {
  int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= P) return;
  int mx = 0;
  for (int j = lane; j < K; j += 32) mx = max(mx, iters[(size_t)p * K + j]);
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) key[p] = mx;
}
void launch_problem_work(const int32_t* iters, int P, int K, int32_t* key, cudaStream_t st) {
  k_problem_work<<<(P + 3) / 4, 128, 0, st>>>(iters, P, K, key);
}

int launch_sort_theta(const double* theta, int P, int K, const int32_t* order, int32_t* perm, cudaStream_t st) {
  if (K < 2 || K > 4096) {  // nothing to sort / does not fit one CTA: identity within the problem
    if (!order) return -1;  // caller keeps the identity
    k_order_slots<<<(unsigned)(((size_t)P * K + 255) / 256), 256, 0, st>>>(order, P, K, perm);
    return 0;
  }
  int Kpad = 2;
  while (Kpad < K) Kpad <<= 1;
  int threads = Kpad / 2 < 1024 ? (Kpad / 2 < 32 ? 32 : Kpad / 2) : 1024;
  k_sort_theta<<<P, threads, (size_t)Kpad * 12, st>>>(theta, K, Kpad, order, perm);
  return 0;
}

// ---- FP64 FMA throughput probe ----------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_probe(double* sink, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double x = 0.999999, y = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) sink[0] = s;  // never true; keeps the chain alive
}

double launch_fp64_probe(double* sink, int iters, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = sms * 8, threads = 256;
  k_fp64_probe<<<blocks, threads, 0, st>>>(sink, iters);
  return 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
}

}  // namespace rll
