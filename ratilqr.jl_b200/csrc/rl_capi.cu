// rl_capi.cu -- the C ABI of libratilqr_b200.so (include/ratilqr.h): context, staging of host
// arrays into the device SoA workspace, kernel launches on the ctx stream, result download.
// There is no CPU compute path in this file: every entry point ends in a kernel launch.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "rl_host.hpp"
#include "rl_launch.hpp"
#include "rl_user_host.hpp"

namespace {

struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct ratilqr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  // staged solve
  bool staged = false;
  int model_id = 0, cost_id = 0, n = 0, m = 0, N = 0, B = 0, eps_cap = 0;
  bool coop = false;          // staged solve runs on the warp-cooperative kernel
  bool coop2 = false;         // ... on its two-warp speculative variant (small batches: latency)
  bool dynamic = false;       // staged solve uses the persistent kernel with lane-level refill
  int spec_G = 0;             // > 0: staged solve runs on the speculative latency kernel with this many lanes per instance
  bool traj_retained = false; // xo/lo/Lo hold the trajectories (coop / dynamic modes)
  DBuf d_queue;
  int coop_cost_id = 0;
  DBuf d_coop_traj;
  rl::SolveParams sp;
  DBuf d_cp, d_W, d_Winv, d_detW, d_x0, d_u, d_theta, d_X;
  DBuf d_value, d_status, d_iters, d_trials, d_restarts, d_mu, d_d, d_cur, d_eps, d_perm;
  DBuf d_out1, d_out2, d_out3;  // host-layout staging for x, l, L
  // user-extensible device models (NVRTC): id = RATILQR_MODEL_USER_BASE + index
  std::vector<rlu::Module*> user_models;
  const rlu::Module* staged_user = nullptr;
  // sub-fleet contexts (own stream + buffers) for the concurrent fleet solve; they resolve user models through `parent`
  std::vector<ratilqr_ctx*> children;
  const ratilqr_ctx* parent = nullptr;
  std::vector<int32_t> fleet_key;  // per-problem work of the last fleet round (slot ordering across rounds and MPC steps)
  DBuf d_order, d_key;
  int32_t* h_key = nullptr;        // pinned
  int h_key_cap = 0, profile_pending = 0;
  // scratch for component calls
  DBuf s[16];
  DBuf d_mix[3];  // true-model noise mixture: cumulative weights, means, Cholesky factors
  DBuf d_cost;
  // multi-GPU (ratilqr_attach_comm / ratilqr_create_multi): NCCL communicator of this ctx's device; rank / world of the group
  void* comm = nullptr;
  int rank = 0, world = 1;
  DBuf d_sh[6];  // sharded population: theta, value, status (full size) and the padded all-gather buffers
  // receding-horizon driver (ratilqr_mpc_fleet_run): states / warm starts stay in d_x0 / d_u between the steps
  bool inputs_on_device = false;  // stage_internal: x0 / u_init are already in d_x0 / d_u (one block per problem)
  DBuf d_mpc[7];                  // x_traj, u_traj, theta_traj, value_traj, noise, cholW, err
};

#define CU(expr)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                     \
      return -100 - (int)e__;                                                             \
    }                                                                                     \
  } while (0)

#define FAIL(code, msg) do { ctx->err = (msg); return (code); } while (0)

static int check_launch(ratilqr_ctx* ctx, const char* what, int nlaunch = 1);
// ---- NCCL, loaded at run time (dlopen, like NVRTC): the library still loads on boxes without NCCL, and inside a process
// that already carries a libnccl.so.2 (torch) the loader hands back that copy instead of a second one ----------------------
namespace nccl {
typedef struct { char internal[128]; } UniqueId;  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef int (*GetUniqueId_t)(UniqueId*);
typedef int (*CommInitRank_t)(void**, int, UniqueId, int);
typedef int (*CommInitAll_t)(void**, int, const int*);
typedef int (*CommDestroy_t)(void*);
typedef int (*AllGather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*GetErrorString_t)(int);
static GetUniqueId_t GetUniqueId;
static CommInitRank_t CommInitRank;
static CommInitAll_t CommInitAll;
static CommDestroy_t CommDestroy;
static AllGather_t AllGather;
static GetErrorString_t GetErrorString;
constexpr int kInt32 = 2, kFloat64 = 8;  // ncclInt32, ncclFloat64
static const char* load() {
  static std::mutex mu;
  static const char* err = nullptr;
  static bool done = false;
  std::lock_guard<std::mutex> lk(mu);
  if (done) return err;
  done = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return err = "libnccl.so.2 cannot be loaded (multi-GPU entry points need NCCL)";
  GetUniqueId = (GetUniqueId_t)dlsym(h, "ncclGetUniqueId");
  CommInitRank = (CommInitRank_t)dlsym(h, "ncclCommInitRank");
  CommInitAll = (CommInitAll_t)dlsym(h, "ncclCommInitAll");
  CommDestroy = (CommDestroy_t)dlsym(h, "ncclCommDestroy");
  AllGather = (AllGather_t)dlsym(h, "ncclAllGather");
  GetErrorString = (GetErrorString_t)dlsym(h, "ncclGetErrorString");
  if (!GetUniqueId || !CommInitRank || !CommInitAll || !CommDestroy || !AllGather || !GetErrorString) return err = "libnccl.so.2 lacks an expected symbol";
  return nullptr;
}
}  // namespace nccl
#define NC(expr)                                                                                   \
  do {                                                                                             \
    int r__ = (expr);                                                                              \
    if (r__ != 0) { ctx->err = std::string(#expr) + ": " + nccl::GetErrorString(r__); return -200 - r__; } \
  } while (0)

// sharded population: rank r owns the samples [S r / W, S (r+1) / W); blocks travel padded to ceil(S / W)
__global__ void k_shard_pack(int cnt, const double* value, const int32_t* status, double* gv, int32_t* gs, int blk, int rank) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= blk) return;
  gv[(size_t)rank * blk + i] = i < cnt ? value[i] : HUGE_VAL;
  gs[(size_t)rank * blk + i] = i < cnt ? status[i] : -1;
}
__global__ void k_shard_unpack(int S, int W, int blk, const double* gv, const int32_t* gs, double* value, int32_t* status) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  int r = (int)(((long long)i * W + W - 1) / S);  // candidate owner; fix up below
  while (r > 0 && (long long)S * r / W > i) --r;
  while (r + 1 < W && (long long)S * (r + 1) / W <= i) ++r;
  const int lo = (int)((long long)S * r / W);
  value[i] = gv[(size_t)r * blk + (i - lo)];
  if (status) status[i] = gs[(size_t)r * blk + (i - lo)];
}
// all-gather of the shard's (value, status) into full-size device arrays; every rank ends up with the same S entries
static int shard_allgather(ratilqr_ctx* ctx, int S, int lo, int cnt, const double* value, const int32_t* status,
                           double* value_full, int32_t* status_full) {
  const int W = ctx->world, blk = (S + W - 1) / W;
  (void)lo;
  CU(ctx->d_sh[3].reserve((size_t)W * blk * 8)); CU(ctx->d_sh[4].reserve((size_t)W * blk * 4));
  double* gv = ctx->d_sh[3].as<double>();
  int32_t* gs = ctx->d_sh[4].as<int32_t>();
  k_shard_pack<<<(blk + 127) / 128, 128, 0, ctx->stream>>>(cnt, value, status, gv, gs, blk, ctx->rank);
  NC(nccl::AllGather(gv + (size_t)ctx->rank * blk, gv, (size_t)blk, nccl::kFloat64, ctx->comm, ctx->stream));  // in place
  NC(nccl::AllGather(gs + (size_t)ctx->rank * blk, gs, (size_t)blk, nccl::kInt32, ctx->comm, ctx->stream));
  k_shard_unpack<<<(S + 127) / 128, 128, 0, ctx->stream>>>(S, W, blk, gv, gs, value_full, status_full);
  return check_launch(ctx, "k_shard_pack/unpack", 2);
}

static int upload(ratilqr_ctx* ctx, DBuf& b, const void* src, size_t bytes) {
  CU(b.reserve(bytes ? bytes : 8));
  if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
#define UP(buf, src, bytes) do { int rc__ = upload(ctx, buf, src, bytes); if (rc__) return rc__; } while (0)

static int check_launch(ratilqr_ctx* ctx, const char* what, int nlaunch) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); return -100 - (int)e; }
  ctx->launches += nlaunch;
  return 0;
}

static const char* check_opts(const ratilqr_ileqg_opts* o) {  // the @asserts of ileqg.jl:195-201
  if (!o) return "null opts";
  if (!(o->lambda > 0 && o->lambda < 1)) return "lambda has to be in (0, 1)";
  if (!(o->d > 0)) return "d > 0 is necessary";
  if (!(o->mu_min > 0)) return "mu_min > 0 is necessary";
  if (!(o->delta_0 > 0)) return "delta_0 > 0 is necessary";
  if (!(o->eps_init > 0 && o->eps_init <= 1)) return "eps_init has to be in (0, 1]";
  if (!(o->eps_init > o->eps_min)) return "eps_init > eps_min is necessary";
  if (!(o->eps_min > 0 && o->eps_min < 1)) return "eps_min has to be in (0, 1)";
  if (o->iter_max < 1) return "iter_max must be >= 1";
  return nullptr;
}

__global__ void k_ce_cost(int B, const double* value, const int32_t* status, const double* theta, double kl, double* cost) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // cost = value + kl_bound/theta (cross_entropy_bilevel_optimization.jl:193); any exception => Inf (:161-165)
  cost[b] = status[b] == 0 ? value[b] + kl / theta[b] : HUGE_VAL;
}

// ---- work profile: per-problem iteration counts of the last launch, used to order the next one ----------------------
// The number of iLEQG iterations is a property of the PROBLEM (x0, goal) far more than of theta, and the bilevel
// optimisers call the batched solve again and again on the same problems with fresh theta populations (one call per
// CE iteration, cross_entropy_bilevel_optimization.jl:291-334; one fleet solve per MPC step).  So the iterations each
// problem needed last time predict what it needs now (round-to-round correlation 0.95 on the C5 fleet), and laying the
// problems out heaviest-first removes the launch tail formed by the ~5 % that run to iter_max (list scheduling:
// makespan 1.06x ideal in natural order, 1.001x heaviest-first).  Pure scheduling: results do not depend on it.
// RATILQR_PROFILE_ORDER=0 disables it.
static bool profile_enabled() {
  static const bool on = [] { const char* e = getenv("RATILQR_PROFILE_ORDER"); return !(e && e[0] == '0'); }();
  return on;
}
// enqueue: key[p] = max iterations of problem p -> pinned host buffer; valid after the caller's next stream sync
static int work_profile_enqueue(ratilqr_ctx* ctx, int P, int K) {
  if (P < 64 || !profile_enabled() || ctx->coop) { ctx->profile_pending = 0; return 0; }
  CU(ctx->d_key.reserve((size_t)P * 4));
  if (ctx->h_key_cap < P) {
    if (ctx->h_key) cudaFreeHost(ctx->h_key);
    ctx->h_key = nullptr; ctx->h_key_cap = 0;
    CU(cudaMallocHost(&ctx->h_key, (size_t)P * 4));
    ctx->h_key_cap = P;
  }
  rll::launch_problem_work(ctx->sp.iters, P, K, ctx->d_key.as<int32_t>(), ctx->stream);
  if (int rc = check_launch(ctx, "k_problem_work")) return rc;
  CU(cudaMemcpyAsync(ctx->h_key, ctx->d_key.p, (size_t)P * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->profile_pending = P;
  return 0;
}
static void work_profile_commit(ratilqr_ctx* ctx) {  // after a stream synchronisation
  if (ctx->profile_pending > 0) ctx->fleet_key.assign(ctx->h_key, ctx->h_key + ctx->profile_pending);
  ctx->profile_pending = 0;
}

// ---- user-extensible device models: lookup, validation, launch ------------------------------------------------
static const rlu::Module* find_user(const ratilqr_ctx* ctx, int model_id) {
  if (ctx->parent) ctx = ctx->parent;
  const int i = model_id - RATILQR_MODEL_USER_BASE;
  return (i >= 0 && i < (int)ctx->user_models.size()) ? ctx->user_models[i] : nullptr;
}

// nullptr if the description is fine, else a message; *um receives the user module (or nullptr)
static const char* check_desc_ctx(const ratilqr_ctx* ctx, const ratilqr_problem_desc* d, bool differentiable,
                                  const rlu::Module** um) {
  *um = nullptr;
  if (!d) return "null problem description";
  if (d->model_id < RATILQR_MODEL_USER_BASE) return rlh::check_desc(d, differentiable);
  const rlu::Module* u = find_user(ctx, d->model_id);
  if (!u) return "unknown user model id (ratilqr_user_model_register returns the id; ids are per ctx)";
  if (d->n != u->spec.n || d->m != u->spec.m) return "n/m do not match the registered user model";
  if (d->n_model_params != u->spec.n_model_params || (d->n_model_params > 0 && !d->model_params)) return "wrong number of model parameters";
  if (d->cost_id != u->cost_id) return "cost_id does not match the registered user model (RATILQR_COST_USER or its base cost)";
  const int ncp = u->spec.cost_src.empty() ? rlh::cost_param_count(d->cost_id, d->n, d->m) : u->spec.n_cost_params;
  if (d->n_cost_params != ncp || (ncp > 0 && !d->cost_params)) return "wrong number of cost parameters";
  if (d->N < 1) return "N must be >= 1";
  if (d->cost_params_count < 1) return "cost_params_count must be >= 1";
  if (differentiable && !u->differentiable) return "L1_CONTROL cost is rollout-only (PETS)";
  if (!d->W) return "W is null";
  *um = u;
  return nullptr;
}

static int user_launch(ratilqr_ctx* ctx, const rlu::Module* um, rlu::Kernel k, unsigned gx, unsigned gy, unsigned block,
                       size_t smem, void* args, const char* what) {
  std::string e;
  if (int rc = rlu::launch(*um, k, gx, gy, block, smem, ctx->stream, args, e)) { ctx->err = std::string(what) + " (user model): " + e; return rc; }
  ctx->launches += 1;
  return 0;
}
#define RL_GRID(total, th) ((unsigned)(((total) + (th)-1) / (th)))

static int stage_internal(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                          const ratilqr_batch_in* in, int eps_cap, bool device_theta = false, int want_traj = 0) {
  ctx->staged = false;
  const rlu::Module* um = nullptr;
  if (const char* m = check_desc_ctx(ctx, desc, true, &um)) FAIL(-1, m);
  ctx->staged_user = um;
  if (const char* m = check_opts(opts)) FAIL(-3, m);
  const bool dev_in = ctx->inputs_on_device;
  if (!in || in->P < 1 || in->K < 1 || ((!in->x0 || !in->u_init) && !dev_in) || (!in->theta && !device_theta)) FAIL(-1, "bad batch description");
  if ((in->x0_count != 1 && in->x0_count != in->P) || (in->u_count != 1 && in->u_count != in->P)) FAIL(-1, "x0_count/u_count must be 1 or P");
  if (desc->cost_params_count != 1 && desc->cost_params_count != in->P) FAIL(-1, "cost_params_count must be 1 or P");
  CU(cudaSetDevice(ctx->device));
  const int n = desc->n, m = desc->m, N = desc->N;
  const size_t B = (size_t)in->P * in->K;
  if (B > 0x7fffffffull) FAIL(-1, "batch too large");
  rlh::WPrep wp;
  if (!rlh::prep_W(n, N, desc->W, desc->W_time_varying, wp)) FAIL(-2, "W(k) is not positive definite (inv(W) / MvNormal need PD)");
  UP(ctx->d_W, wp.W.data(), wp.W.size() * 8);
  UP(ctx->d_Winv, wp.Winv.data(), wp.Winv.size() * 8);
  UP(ctx->d_detW, wp.detW.data(), wp.detW.size() * 8);
  UP(ctx->d_cp, desc->cost_params, (size_t)desc->n_cost_params * desc->cost_params_count * 8);
  if (!dev_in) {
    UP(ctx->d_x0, in->x0, (size_t)n * in->x0_count * 8);
    UP(ctx->d_u, in->u_init, (size_t)m * N * in->u_count * 8);
  }
  if (device_theta) CU(ctx->d_theta.reserve(B * 8));  // theta is produced on the device (fleet CE)
  else UP(ctx->d_theta, in->theta, B * 8);
  // Kernel choice, part 1: a batch that leaves most of the machine idle runs on the speculative latency kernel
  // (rl_spec.cuh): G lanes per instance evaluate several line-search candidates at once and run the next iteration's
  // optimising pass alongside each candidate's evaluating pass.  G = 8 while every warp still gets a scheduler of its own
  // (148 SMs x 4), G = 4 up to twice that; beyond, lanes are worth more as instances (one thread per instance).
  ctx->spec_G = 0;
  {
    const int cid = (!um && rlh::quad_is_diag(desc) && desc->model_id != RATILQR_MODEL_POWER_LAW) ? RL_COST_QUAD_DIAG : desc->cost_id;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const size_t warps1 = (size_t)sms * 4;
    const char* e = getenv("RATILQR_SPEC");  // 0 = off, 2 / 4 / 8 = force that many lanes per instance (tuning / A-B runs)
    int G = 0;
    if (e) G = atoi(e);
    else if (B * 8 <= warps1 * 32) G = 8;
    else if (B * 4 <= warps1 * 64) G = 4;
    if (G != 2 && G != 4 && G != 8) G = 0;
    const char* ed = getenv("RATILQR_DYNAMIC");
    const char* ec = getenv("RATILQR_COOP");
    if (!um && n <= 6 && G && rll::spec_supported(desc->model_id, cid) && !(ed && ed[0] == '1') && !(ec && ec[0] == '1')) ctx->spec_G = G;
  }
  const size_t cols = ctx->spec_G ? B * ctx->spec_G : B;  // workspace columns: one per thread
  const size_t pol = ctx->spec_G ? 2 : 1;                 // the speculative kernel double-buffers the policy (Lg, DL)
  const size_t Bp = (cols + 31) / 32 * 32;  // the workspace is tiled in groups of 32 thread slots
  const rl::WsLayout wl = rl::ws_layout(n, m, N, (int)pol, rl::model_naux(desc->model_id));  // one allocation of per-tile records (rl::SolveParams::X)
  CU(ctx->d_X.reserve(wl.rec * Bp * 8));
  CU(ctx->d_value.reserve(B * 8));
  CU(ctx->d_mu.reserve(B * 8));
  CU(ctx->d_d.reserve(B * 8));
  CU(ctx->d_status.reserve(B * 4));
  CU(ctx->d_iters.reserve(B * 4));
  CU(ctx->d_trials.reserve(B * 4));
  CU(ctx->d_restarts.reserve(B * 4));
  CU(ctx->d_cur.reserve(B * 4));
  if (eps_cap > 0) {
    CU(ctx->d_eps.reserve(B * eps_cap * 16));
    CU(cudaMemsetAsync(ctx->d_eps.p, 0, B * eps_cap * 16, ctx->stream));
  }
  // gains of instances that fail before their first optimising pass stay zero (initialize!: L = 0)
  if (!ctx->spec_G)
    CU(cudaMemset2DAsync(ctx->d_X.as<double>() + wl.oLg * 32, wl.rec * 32 * 8, 0, (size_t)N * m * n * 32 * 8, Bp / 32, ctx->stream));
  rl::SolveParams& P = ctx->sp;
  memset(&P, 0, sizeof(P));
  P.N = N; P.B = (int)B; P.K = in->K;
  for (int i = 0; i < 8; ++i) P.mp[i] = i < desc->n_model_params ? desc->model_params[i] : 0.0;
  P.cost_params = ctx->d_cp.as<double>(); P.ncp = desc->n_cost_params; P.cp_count = desc->cost_params_count;
  P.W = ctx->d_W.as<double>(); P.Winv = ctx->d_Winv.as<double>(); P.detW = ctx->d_detW.as<double>(); P.W_tv = desc->W_time_varying;
  P.w_const = (!desc->W_time_varying && n * n <= 36) ? 1 : 0;
  if (P.w_const) {
    for (int i = 0; i < n * n; ++i) { P.Wc[i] = wp.W[i]; P.Winvc[i] = wp.Winv[i]; }
    P.detWc = wp.detW[0];
  }
  P.x0 = ctx->d_x0.as<double>(); P.x0_count = in->x0_count;
  P.u_init = ctx->d_u.as<double>(); P.u_count = in->u_count;
  P.theta = ctx->d_theta.as<double>();
  P.mu_min = opts->mu_min; P.delta_0 = opts->delta_0; P.lambda = opts->lambda; P.d = opts->d;
  P.iter_max = opts->iter_max; P.eps_auto = opts->adaptive_eps_init; P.eps_init = opts->eps_init; P.eps_min = opts->eps_min;
  P.X = ctx->d_X.as<double>(); P.U = P.X + wl.oU * 32; P.Lg = P.X + wl.oLg * 32; P.DL = P.X + wl.oDL * 32; P.AUX = P.X + wl.oAux * 32; P.rec = wl.rec;
  P.value = ctx->d_value.as<double>(); P.status = ctx->d_status.as<int32_t>(); P.iters = ctx->d_iters.as<int32_t>();
  P.trials = ctx->d_trials.as<int32_t>(); P.restarts = ctx->d_restarts.as<int32_t>();
  P.mu_out = ctx->d_mu.as<double>(); P.d_out = ctx->d_d.as<double>(); P.cur = ctx->d_cur.as<int32_t>();
  P.eps_hist = eps_cap > 0 ? ctx->d_eps.as<double>() : nullptr; P.eps_hist_cap = eps_cap;
  { const char* e = getenv("RATILQR_NO_STAGE"); P.use_stage = (e && e[0] == '1') ? 0 : 1; }
  // warp-homogeneous scheduling: lanes of a warp get neighbouring theta of one problem
  CU(ctx->d_perm.reserve(B * 4));
  // ... and, when the previous call solved a fleet of this many problems, problems go heaviest-first (work_profile below)
  const int32_t* order = nullptr;
  if (!device_theta && (int)ctx->fleet_key.size() == in->P && in->P >= 64 && profile_enabled()) {
    std::vector<int32_t> ord(in->P);
    std::iota(ord.begin(), ord.end(), 0);
    const std::vector<int32_t>& key = ctx->fleet_key;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return key[a] > key[b]; });
    UP(ctx->d_order, ord.data(), (size_t)in->P * 4);  // pageable source: cudaMemcpyAsync returns once it has been staged
    order = ctx->d_order.as<int32_t>();
  }
  if (!device_theta && rll::launch_sort_theta(P.theta, in->P, in->K, order, ctx->d_perm.as<int32_t>(), ctx->stream) == 0) {
    if (int rc = check_launch(ctx, "k_sort_theta")) return rc;
    P.perm = ctx->d_perm.as<int32_t>();
  } else {
    P.perm = nullptr;
  }
  ctx->model_id = desc->model_id;
  // structure-specialised kernel when the quadratic cost is diagonal (bit-identical results, fewer flops)
  ctx->cost_id = (!um && rlh::quad_is_diag(desc) && desc->model_id != RATILQR_MODEL_QUADROTOR && desc->model_id != RATILQR_MODEL_POWER_LAW) ? RL_COST_QUAD_DIAG : desc->cost_id;
  ctx->n = n; ctx->m = m; ctx->N = N; ctx->B = (int)B; ctx->eps_cap = eps_cap;
  // kernel choice: the warp-cooperative kernel when the per-stage state is too big for one thread's registers
  // (n > 6: the quadrotor); one thread per instance otherwise.  (For n = 4 the ~20 synchronised shared-memory
  // phases per stage cost more than they save: 44 ms vs 21 ms for 1024 unicycle solves, profiles/r01_coop_*.)
  ctx->coop = false;
  ctx->coop_cost_id = rlh::quad_is_diag(desc) ? RL_COST_QUAD_DIAG : desc->cost_id;
  if (desc->model_id == RATILQR_MODEL_POWER_LAW) ctx->coop_cost_id = desc->cost_id;
  size_t csm = 0;
  if (!um && rll::coop_smem_query(desc->model_id, ctx->coop_cost_id, N, &csm) == 0) {  // user models: one thread per instance
    size_t per_sm = (227 * 1024) / (csm + 1024);
    if (per_sm > 32) per_sm = 32;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const size_t capacity = per_sm * (size_t)sms;
    const char* e = getenv("RATILQR_COOP");  // 1 = force, 0 = never (tuning / A-B runs)
    (void)capacity;
    ctx->coop = e ? (e[0] == '1') : (n > 6);
  }
  if (ctx->spec_G) {  // the speculative kernel writes x, l, L in host layout itself
    ctx->coop = false;
    ctx->cost_id = (rlh::quad_is_diag(desc) && desc->model_id != RATILQR_MODEL_POWER_LAW) ? RL_COST_QUAD_DIAG : desc->cost_id;
    CU(ctx->d_out1.reserve((size_t)n * (N + 1) * B * 8));
    CU(ctx->d_out2.reserve((size_t)m * N * B * 8));
    CU(ctx->d_out3.reserve((size_t)m * n * N * B * 8));
    P.xo = ctx->d_out1.as<double>(); P.lo = ctx->d_out2.as<double>(); P.Lo = ctx->d_out3.as<double>();
  }
  ctx->dynamic = false;
  ctx->traj_retained = ctx->spec_G > 0;
  if (!ctx->coop && !ctx->spec_G) {
    // Lane-level refill (persistent kernel pulling instances from a queue) is an opt-in experiment: measured
    // SLOWER than the static theta-sorted assignment (148 vs 120 ms on the bench fleet, profiles/r01_dynamic_refill_ab.jsonl)
    // because refilled lanes fall out of phase with their warp and most trips then carry a partially used optimising pass.
    // (the device code is only there in a -DRL_ENABLE_DYNAMIC build; the host emulation keeps testing the logic)
#if defined(RL_ENABLE_DYNAMIC)
    const char* e2 = getenv("RATILQR_DYNAMIC");
    ctx->dynamic = (e2 && e2[0] == '1');
#endif
    if (ctx->dynamic) {
      CU(ctx->d_queue.reserve(8));
      P.queue = ctx->d_queue.as<unsigned int>();
      if (want_traj & 1) { CU(ctx->d_out1.reserve((size_t)n * (N + 1) * B * 8)); P.xo = ctx->d_out1.as<double>(); }
      if (want_traj & 2) { CU(ctx->d_out2.reserve((size_t)m * N * B * 8)); P.lo = ctx->d_out2.as<double>(); }
      if (want_traj & 4) { CU(ctx->d_out3.reserve((size_t)m * n * N * B * 8)); P.Lo = ctx->d_out3.as<double>(); }
      ctx->traj_retained = true;
    }
  }
  ctx->coop2 = false;
  if (ctx->coop) {
    // small batches: the two-warp variant (pass speculation) while every instance still gets an SM share of its own
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const size_t c2 = rll::coop2_smem_query(desc->model_id, ctx->coop_cost_id, N);
    const char* e = getenv("RATILQR_COOP2");  // 0 = never, 1 = whenever it fits (tuning / A-B runs)
    const bool fits = c2 > 0 && c2 <= 200 * 1024;
    ctx->coop2 = fits && (e ? e[0] == '1' : B <= (size_t)3 * sms);
  }
  if (ctx->coop) {
    ctx->traj_retained = true;
    P.perm = nullptr;
    CU(ctx->d_out1.reserve((size_t)n * (N + 1) * B * 8));
    CU(ctx->d_out2.reserve((size_t)m * N * B * 8));
    CU(ctx->d_out3.reserve((size_t)m * n * N * B * 8));
    P.xo = ctx->d_out1.as<double>(); P.lo = ctx->d_out2.as<double>(); P.Lo = ctx->d_out3.as<double>();
    CU(cudaMemsetAsync(P.Lo, 0, (size_t)m * n * N * B * 8, ctx->stream));
    if (csm > 200 * 1024 || csm == 0) {}  // (trajectories in shared memory)
    CU(ctx->d_coop_traj.reserve(rl::coop_traj_doubles(n, m, N) * B * 8));
  }
  ctx->staged = true;
  return 0;
}

static int run_internal(ratilqr_ctx* ctx, int reps, float* ms_total) {
  if (!ctx->staged) FAIL(-4, "nothing staged: call ratilqr_ileqg_stage first");
  if (reps < 1) reps = 1;
  CU(cudaSetDevice(ctx->device));
  if (ms_total) CU(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int r = 0; r < reps; ++r) {
    if (ctx->coop && ctx->coop2) {
      if (rll::launch_solve_coop2(ctx->model_id, ctx->coop_cost_id, ctx->sp, ctx->stream)) FAIL(-5, "this (model, cost) pair is not compiled in");
      if (int rc = check_launch(ctx, "k_ileqg_solve_coop2")) return rc;
      continue;
    }
    if (ctx->coop) {
      if (rll::launch_solve_coop(ctx->model_id, ctx->coop_cost_id, ctx->sp, ctx->d_coop_traj.as<double>(), ctx->stream)) FAIL(-5, "this (model, cost) pair is not compiled in");
      if (int rc = check_launch(ctx, "k_ileqg_solve_coop")) return rc;
      continue;
    }
    if (ctx->spec_G) {
      if (rll::launch_solve_spec(ctx->model_id, ctx->cost_id, ctx->spec_G, ctx->sp, ctx->stream)) FAIL(-5, "this (model, cost) pair is not compiled in");
      if (int rc = check_launch(ctx, "k_ileqg_solve_spec")) return rc;
      continue;
    }
    if (ctx->dynamic) CU(cudaMemsetAsync(ctx->sp.queue, 0, 4, ctx->stream));
    if (ctx->staged_user) {
      const rlu::Module* um = ctx->staged_user;
      if (int rc = user_launch(ctx, um, rlu::K_SOLVE, RL_GRID(ctx->sp.B, um->solve_threads), 1, um->solve_threads, um->solve_smem, &ctx->sp, "k_ileqg_solve")) return rc;
      continue;
    }
    if (rll::launch_solve(ctx->model_id, ctx->cost_id, ctx->sp, ctx->stream)) FAIL(-5, "this (model, cost) pair is not compiled in");
    if (int rc = check_launch(ctx, "k_ileqg_solve")) return rc;
  }
  if (ms_total) {
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev1));
    CU(cudaEventElapsedTime(ms_total, ctx->ev0, ctx->ev1));
  }
  return 0;
}

// Fleet scheduling: order the thread slots by the work each PROBLEM needed in the previous launch (max iterations
// over its instances), so that the 32 lanes of a warp hold problems of similar length.  Host-side argsort of P keys
// (the fleet loop synchronises once per round anyway).
// slot r*K + j <- instance order[r]*K + j, problems ordered by `key` (their predicted work), HEAVIEST FIRST: CTAs are
// dispatched in grid order, so the few problems that run to iter_max (the mu quirk of ileqg.jl:471-488; ~5 % of a
// fleet, 5x the median work) start with the launch instead of forming its tail (longest-processing-time-first).
static int apply_slot_order(ratilqr_ctx* ctx, const std::vector<int32_t>& key, int P, int K) {
  static const bool ascending = [] { const char* e = getenv("RATILQR_FLEET_ORDER"); return e && e[0] == 'a'; }();  // A/B runs
  const size_t B = (size_t)P * K;
  std::vector<int32_t> order(P), perm(B);
  std::iota(order.begin(), order.end(), 0);
  if (ascending) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  else std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] > key[b]; });
  for (int r = 0; r < P; ++r) for (int j = 0; j < K; ++j) perm[(size_t)r * K + j] = (int32_t)((size_t)order[r] * K + j);
  CU(ctx->d_perm.reserve(B * 4));
  // (pageable source: the copy call returns once the vector has been staged for DMA, so it may go out of scope)
  CU(cudaMemcpyAsync(ctx->d_perm.p, perm.data(), B * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->sp.perm = ctx->d_perm.as<int32_t>();
  return 0;
}

// key[p] = max iterations over the K instances of problem p in the launch that just finished.  Iteration counts are
// a property of the problem far more than of theta (round-to-round correlation 0.95 on the C5 fleet), so the key also
// predicts the next round -- and, kept in the context, the first round of the next MPC step of the same fleet.
static int resort_slots_by_last_iters(ratilqr_ctx* ctx, int P, int K) {
  const size_t B = (size_t)P * K;
  std::vector<int32_t> iters(B);
  CU(cudaMemcpyAsync(iters.data(), ctx->sp.iters, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t>& key = ctx->fleet_key;
  key.assign(P, 0);
  for (int p = 0; p < P; ++p) { int32_t mx = 0; for (int j = 0; j < K; ++j) mx = std::max(mx, iters[(size_t)p * K + j]); key[p] = mx; }
  return apply_slot_order(ctx, key, P, K);
}

// device pointer to l_array (m*N*B, host layout) of the staged solve that just ran: the kernels that keep trajectories in
// host layout have it already, the thread-per-instance kernel's workspace goes through the tile-transpose gather
static int device_plan(ratilqr_ctx* ctx, const double** plan) {
  if (!ctx->staged) FAIL(-4, "nothing staged");
  if (ctx->traj_retained) {
    if (!ctx->sp.lo) FAIL(-4, "l_array was not retained by this staged solve");
    *plan = ctx->sp.lo;
    return 0;
  }
  const size_t B = (size_t)ctx->B;
  CU(ctx->d_out2.reserve((size_t)ctx->m * ctx->N * B * 8));
  rll::launch_gather(ctx->n, ctx->m, ctx->N, (int)B, ctx->sp.X, ctx->sp.U, ctx->sp.Lg, ctx->sp.rec, ctx->sp.cur, ctx->sp.perm, nullptr,
                     ctx->d_out2.as<double>(), nullptr, ctx->stream);
  if (int rc = check_launch(ctx, "k_gather")) return rc;
  *plan = ctx->d_out2.as<double>();
  return 0;
}

static int fetch_internal(ratilqr_ctx* ctx, ratilqr_ileqg_out* out) {
  if (!ctx->staged) FAIL(-4, "nothing staged");
  if (!out) FAIL(-1, "null out");
  CU(cudaSetDevice(ctx->device));
  const int n = ctx->n, m = ctx->m, N = ctx->N;
  const size_t B = (size_t)ctx->B;
  cudaStream_t st = ctx->stream;
  double *dx = nullptr, *dl = nullptr, *dL = nullptr;
  if (!ctx->traj_retained) {
    if (out->x) { CU(ctx->d_out1.reserve((size_t)n * (N + 1) * B * 8)); dx = ctx->d_out1.as<double>(); }
    if (out->l) { CU(ctx->d_out2.reserve((size_t)m * N * B * 8)); dl = ctx->d_out2.as<double>(); }
    if (out->L) { CU(ctx->d_out3.reserve((size_t)m * n * N * B * 8)); dL = ctx->d_out3.as<double>(); }
  }
  if (ctx->traj_retained) {  // the cooperative / persistent kernels wrote x, l, L in host layout themselves
    if ((out->x && !ctx->sp.xo) || (out->l && !ctx->sp.lo) || (out->L && !ctx->sp.Lo))
      FAIL(-4, "trajectories were not retained by this staged solve: request them through ratilqr_ileqg_solve_batch");
    dx = out->x ? ctx->sp.xo : nullptr; dl = out->l ? ctx->sp.lo : nullptr; dL = out->L ? ctx->sp.Lo : nullptr;
  } else if (dx || dl || dL) {
    rll::launch_gather(n, m, N, (int)B, ctx->sp.X, ctx->sp.U, ctx->sp.Lg, ctx->sp.rec, ctx->sp.cur, ctx->sp.perm, dx, dl, dL, st);
    if (int rc = check_launch(ctx, "k_gather", (dx ? 1 : 0) + (dl ? 1 : 0) + (dL ? 1 : 0))) return rc;
  }
#define DOWN(dst, src, bytes) if (dst) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st))
  DOWN(out->x, dx, (size_t)n * (N + 1) * B * 8);
  DOWN(out->l, dl, (size_t)m * N * B * 8);
  DOWN(out->L, dL, (size_t)m * n * N * B * 8);
  DOWN(out->value, ctx->sp.value, B * 8);
  DOWN(out->status, ctx->sp.status, B * 4);
  DOWN(out->iters, ctx->sp.iters, B * 4);
  DOWN(out->trials, ctx->sp.trials, B * 4);
  DOWN(out->restarts, ctx->sp.restarts, B * 4);
  DOWN(out->mu, ctx->sp.mu_out, B * 8);
  DOWN(out->d_current, ctx->sp.d_out, B * 8);
  if (out->eps_hist && out->eps_hist_cap > 0) {
    if (out->eps_hist_cap != ctx->eps_cap) FAIL(-1, "eps_hist_cap differs from the staged capacity");
    DOWN(out->eps_hist, ctx->sp.eps_hist, B * ctx->eps_cap * 16);
  }
#undef DOWN
  CU(cudaStreamSynchronize(st));
  return 0;
}

extern "C" {

int32_t ratilqr_version(void) { return 100; }

int32_t ratilqr_model_dims(int32_t model_id, int32_t* n, int32_t* m, int32_t* n_params) {
  int a, b, c;
  if (!rlh::model_dims(model_id, &a, &b, &c)) return -1;
  if (n) *n = a;
  if (m) *m = b;
  if (n_params) *n_params = c;
  return 0;
}

int32_t ratilqr_cost_param_count(int32_t cost_id, int32_t n, int32_t m) { return rlh::cost_param_count(cost_id, n, m); }

int32_t ratilqr_create(ratilqr_ctx** out, int32_t device_id) {
  if (!out) return -1;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count < 1) return -10;  // no CUDA device: there is no CPU fallback
  if (device_id < 0 || device_id >= count) return -11;
  ratilqr_ctx* ctx = new ratilqr_ctx();
  ctx->device = device_id;
  if (cudaSetDevice(device_id) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
    delete ctx;
    return -12;
  }
  *out = ctx;
  return 0;
}

int32_t ratilqr_destroy(ratilqr_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  DBuf* all[] = {&ctx->d_cp, &ctx->d_W, &ctx->d_Winv, &ctx->d_detW, &ctx->d_x0, &ctx->d_u, &ctx->d_theta, &ctx->d_X,
                 &ctx->d_value, &ctx->d_status, &ctx->d_iters, &ctx->d_trials,
                 &ctx->d_restarts, &ctx->d_mu, &ctx->d_d, &ctx->d_cur, &ctx->d_eps, &ctx->d_perm, &ctx->d_out1, &ctx->d_out2,
                 &ctx->d_out3, &ctx->d_cost, &ctx->d_coop_traj, &ctx->d_queue};
  for (DBuf* b : all) b->release();
  for (DBuf& b : ctx->s) b.release();
  for (DBuf& b : ctx->d_mix) b.release();
  for (DBuf& b : ctx->d_mpc) b.release();
  for (DBuf& b : ctx->d_sh) b.release();
  if (ctx->comm && nccl::CommDestroy) nccl::CommDestroy(ctx->comm);
  ctx->d_order.release(); ctx->d_key.release();
  if (ctx->h_key) cudaFreeHost(ctx->h_key);
  for (rlu::Module* um : ctx->user_models) { rlu::unload(*um); delete um; }
  for (ratilqr_ctx* ch : ctx->children) ratilqr_destroy(ch);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

static void copy_log(const std::string& src, char* log, int64_t cap) {
  if (!log || cap < 1) return;
  size_t k = src.size() < (size_t)cap - 1 ? src.size() : (size_t)cap - 1;
  memcpy(log, src.data(), k);
  log[k] = 0;
}

static int user_spec_from(const ratilqr_user_model_desc* um, rlu::Spec& sp) {
  if (!um) return -1;
  sp.n = um->n; sp.m = um->m; sp.n_model_params = um->n_model_params; sp.n_cost_params = um->n_cost_params;
  sp.base_model_id = um->base_model_id; sp.base_cost_id = um->base_cost_id;
  sp.dynamics_src = um->dynamics_src ? um->dynamics_src : "";
  sp.cost_src = um->cost_src ? um->cost_src : "";
  if (sp.dynamics_src.empty()) {  // registered dynamics: its own parameter count
    int n, m, np;
    if (rlh::model_dims(sp.base_model_id, &n, &m, &np)) sp.n_model_params = np;
  }
  if (sp.cost_src.empty()) sp.n_cost_params = rlh::cost_param_count(sp.base_cost_id, sp.n, sp.m);
  if (sp.n < 1 || sp.n > 16 || sp.m < 1 || sp.m > 4) return 0;  // rejected by rlu::compile with a message
  const size_t n = sp.n, m = sp.m;
  if (!sp.dynamics_src.empty()) {
    if (um->a_kind) sp.a_kind.assign(um->a_kind, um->a_kind + n * n);
    if (um->b_kind) sp.b_kind.assign(um->b_kind, um->b_kind + n * m);
  }
  if (!sp.cost_src.empty()) {
    if (um->q_kind) sp.q_kind.assign(um->q_kind, um->q_kind + n * n);
    if (um->r_kind) sp.r_kind.assign(um->r_kind, um->r_kind + m * m);
    if (um->p_kind) sp.p_kind.assign(um->p_kind, um->p_kind + m * n);
  }
  return 0;
}

// Checks the declared structure of a freshly loaded user model against its dual-number derivatives at a few generic
// points (generic parameters too: a structural zero / one holds for every parameter value, so this can only reject
// false declarations).  Points where the model leaves its domain are skipped.  Returns "" or what is wrong.
static std::string verify_user_structure(ratilqr_ctx* ctx, const rlu::Module& mod) {
  const rlu::Spec& sp = mod.spec;
  if (!mod.differentiable) return "";
  if (sp.a_kind.empty() && sp.b_kind.empty() && sp.q_kind.empty() && sp.r_kind.empty() && sp.p_kind.empty()) return "";
  const int n = sp.n, m = sp.m, N = 1, B = 6;
  uint64_t lcg = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return (double)(lcg >> 11) * (1.0 / 9007199254740992.0); };
  std::vector<double> mp(std::max(sp.n_model_params, 1)), cp(std::max(sp.n_cost_params, 1)), W((size_t)n * n, 0.0);
  for (auto& v : mp) v = 0.3 + 0.6 * rnd();
  for (auto& v : cp) v = 0.3 + 0.6 * rnd();
  for (int i = 0; i < n; ++i) W[i + (size_t)i * n] = 1.0;
  std::vector<double> x((size_t)n * (N + 1) * B), u((size_t)m * N * B);
  for (auto& v : x) v = 0.2 + 1.3 * rnd();
  for (auto& v : u) v = -1.0 + 2.0 * rnd();
  ratilqr_problem_desc d;
  memset(&d, 0, sizeof(d));
  d.model_id = mod.id; d.cost_id = mod.cost_id; d.n = n; d.m = m; d.N = N;
  d.model_params = mp.data(); d.n_model_params = sp.n_model_params;
  d.cost_params = cp.data(); d.n_cost_params = sp.n_cost_params; d.cost_params_count = 1;
  d.W = W.data(); d.W_time_varying = 0;
  std::vector<double> q((size_t)(N + 1) * B), qv((size_t)n * (N + 1) * B), Q((size_t)n * n * (N + 1) * B), r((size_t)m * N * B),
      R((size_t)m * m * N * B), Pm((size_t)m * n * N * B), A((size_t)n * n * N * B), Bm((size_t)n * m * N * B);
  std::vector<int32_t> st(B, 0);
  if (ratilqr_linearize_batch(ctx, &d, B, x.data(), u.data(), q.data(), qv.data(), Q.data(), r.data(), R.data(), Pm.data(),
                              A.data(), Bm.data(), st.data()))
    return "structure check could not run: " + ctx->err;
  auto bad = [&](const char* name, const std::vector<signed char>& k, const double* M, int rows, int cols, size_t per_inst,
                 size_t stage_off) -> std::string {
    if (k.empty()) return "";
    for (int b = 0; b < B; ++b) {
      if (st[b]) continue;
      const double* Mb = M + per_inst * b + stage_off;
      for (int j = 0; j < cols; ++j)
        for (int i = 0; i < rows; ++i) {
          const int kind = k[(size_t)i + (size_t)j * rows];
          const double v = Mb[(size_t)i + (size_t)j * rows];
          if ((kind == 0 && v != 0.0) || (kind == 1 && v != 1.0))
            return std::string(name) + "[" + std::to_string(i) + "," + std::to_string(j) + "] is declared " +
                   (kind == 0 ? "zero" : "one") + " but evaluates to " + std::to_string(v);
        }
    }
    return "";
  };
  std::string e;
  if (!(e = bad("a_kind", sp.a_kind, A.data(), n, n, (size_t)n * n * N, 0)).empty()) return e;
  if (!(e = bad("b_kind", sp.b_kind, Bm.data(), n, m, (size_t)n * m * N, 0)).empty()) return e;
  if (!(e = bad("q_kind", sp.q_kind, Q.data(), n, n, (size_t)n * n * (N + 1), 0)).empty()) return e;
  if (!(e = bad("q_kind (terminal cost)", sp.q_kind, Q.data(), n, n, (size_t)n * n * (N + 1), (size_t)n * n * N)).empty()) return e;
  if (!(e = bad("r_kind", sp.r_kind, R.data(), m, m, (size_t)m * m * N, 0)).empty()) return e;
  if (!(e = bad("p_kind", sp.p_kind, Pm.data(), m, n, (size_t)m * n * N, 0)).empty()) return e;
  return "";
}

int32_t ratilqr_user_model_check(const ratilqr_user_model_desc* um, char* log, int64_t log_cap) {
  rlu::Spec sp;
  if (user_spec_from(um, sp)) return -1;
  rlu::Compiled c;
  int rc = rlu::compile(sp, c);
  copy_log(c.log, log, log_cap);
  return rc;
}

int32_t ratilqr_user_model_register(ratilqr_ctx* ctx, const ratilqr_user_model_desc* um, int32_t* model_id_out, char* log,
                                    int64_t log_cap) {
  if (!ctx) return -1;
  if (!model_id_out) FAIL(-1, "model_id_out is null");
  rlu::Spec sp;
  if (user_spec_from(um, sp)) FAIL(-1, "null user model description");
  rlu::Compiled c;
  int rc = rlu::compile(sp, c);
  copy_log(c.log, log, log_cap);
  if (rc) FAIL(rc, "user model: " + (c.log.empty() ? std::string("compilation failed") : c.log));
  CU(cudaSetDevice(ctx->device));
  CU(cudaFree(0));  // make sure the primary context is current before the driver-API load
  rlu::Module* mod = new rlu::Module();
  mod->spec = sp;
  mod->id = RATILQR_MODEL_USER_BASE + (int)ctx->user_models.size();
  mod->cost_id = sp.cost_src.empty() ? sp.base_cost_id : RATILQR_COST_USER;
  mod->differentiable = !sp.cost_src.empty() || sp.base_cost_id != RATILQR_COST_L1_CONTROL;
  std::string e;
  if ((rc = rlu::load(c, *mod, e))) { delete mod; FAIL(rc, "user model: " + e); }
  ctx->user_models.push_back(mod);
  const std::string wrong = verify_user_structure(ctx, *mod);
  if (!wrong.empty()) {
    ctx->user_models.pop_back();
    rlu::unload(*mod);
    delete mod;
    copy_log("declared structure is wrong: " + wrong, log, log_cap);
    FAIL(-25, "user model: declared structure is wrong: " + wrong);
  }
  *model_id_out = mod->id;
  return 0;
}

const char* ratilqr_last_error(const ratilqr_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
int64_t ratilqr_launch_count(const ratilqr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t ratilqr_ileqg_stage(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                            const ratilqr_batch_in* in) {
  if (!ctx) return -1;
  int rc = stage_internal(ctx, desc, opts, in, 0);
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t ratilqr_ileqg_run(ratilqr_ctx* ctx, int32_t reps, float* ms_total) {
  if (!ctx) return -1;
  int rc = run_internal(ctx, reps, ms_total);
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t ratilqr_ileqg_fetch(ratilqr_ctx* ctx, ratilqr_ileqg_out* out) {
  if (!ctx) return -1;
  return fetch_internal(ctx, out);
}

int32_t ratilqr_ileqg_solve_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                                  const ratilqr_batch_in* in, ratilqr_ileqg_out* out) {
  if (!ctx) return -1;
  if (!out) FAIL(-1, "null out");
  const int want = (out->x ? 1 : 0) | (out->l ? 2 : 0) | (out->L ? 4 : 0);
  int rc = stage_internal(ctx, desc, opts, in, out->eps_hist ? out->eps_hist_cap : 0, false, want);
  if (rc) return rc;
  rc = run_internal(ctx, 1, nullptr);
  if (rc) return rc;
  return fetch_internal(ctx, out);
}

int32_t ratilqr_ce_costs(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                         const ratilqr_batch_in* in, double kl_bound, double* cost, int32_t* status) {
  if (!ctx) return -1;
  if (!cost) FAIL(-1, "null cost");
  int rc = stage_internal(ctx, desc, opts, in, 0);
  if (rc) return rc;
  rc = run_internal(ctx, 1, nullptr);
  if (rc) return rc;
  const int B = ctx->B;
  CU(ctx->d_cost.reserve((size_t)B * 8));
  k_ce_cost<<<(B + 127) / 128, 128, 0, ctx->stream>>>(B, ctx->sp.value, ctx->sp.status, ctx->sp.theta, kl_bound, ctx->d_cost.as<double>());
  if ((rc = check_launch(ctx, "k_ce_cost"))) return rc;
  CU(cudaMemcpyAsync(cost, ctx->d_cost.p, (size_t)B * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CU(cudaMemcpyAsync(status, ctx->sp.status, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = work_profile_enqueue(ctx, in->P, in->K))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  work_profile_commit(ctx);
  return 0;
}

// true-model noise mixture -> device view (rl::MixtureView); k = 0 when mx is null
static int upload_mixture(ratilqr_ctx* ctx, int n, const ratilqr_noise_mixture* mx, rl::MixtureView& v) {
  v.k = 0; v.cumw = v.mean = v.chol = nullptr;
  if (!mx) return 0;
  rlh::MixPrep mp;
  if (const char* msg = rlh::prep_mixture(n, mx, mp)) FAIL(-2, msg);
  UP(ctx->d_mix[0], mp.cumw.data(), mp.cumw.size() * 8);
  UP(ctx->d_mix[1], mp.mean.data(), mp.mean.size() * 8);
  UP(ctx->d_mix[2], mp.chol.data(), mp.chol.size() * 8);
  v.k = mp.k; v.cumw = ctx->d_mix[0].as<double>(); v.mean = ctx->d_mix[1].as<double>(); v.chol = ctx->d_mix[2].as<double>();
  return 0;
}

// ---- component entry points ---------------------------------------------------------------------
static int fill_comp(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int B, rll::CompArgs& a, bool differentiable,
                     const rlu::Module** um) {
  ctx->staged = false;  // shares d_cp / d_W with a staged solve
  if (const char* m = check_desc_ctx(ctx, desc, differentiable, um)) FAIL(-1, m);
  if (B < 1) FAIL(-1, "B must be >= 1");
  CU(cudaSetDevice(ctx->device));
  memset(&a, 0, sizeof(a));
  a.model_id = desc->model_id; a.cost_id = desc->cost_id; a.n = desc->n; a.m = desc->m; a.N = desc->N; a.B = B;
  for (int i = 0; i < 8; ++i) a.mp[i] = i < desc->n_model_params ? desc->model_params[i] : 0.0;
  UP(ctx->d_cp, desc->cost_params, (size_t)desc->n_cost_params * 8);
  a.cp = ctx->d_cp.as<double>();
  return 0;
}

#define DOWNSYNC(dst, src, bytes) do { if (dst) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream)); } while (0)

int32_t ratilqr_rollout_open_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t B, const double* x0,
                                   const double* u, double* x, int32_t* status) {
  if (!ctx) return -1;
  rll::CompArgs a;
  const rlu::Module* um = nullptr;
  if (int rc = fill_comp(ctx, desc, B, a, false, &um)) return rc;
  const size_t n = a.n, m = a.m, N = a.N;
  UP(ctx->s[0], x0, n * B * 8); UP(ctx->s[1], u, m * N * B * 8);
  CU(ctx->s[2].reserve(n * (N + 1) * B * 8)); CU(ctx->s[3].reserve((size_t)B * 4));
  CU(cudaMemsetAsync(ctx->s[2].p, 0, n * (N + 1) * B * 8, ctx->stream));
  a.x0 = ctx->s[0].as<double>(); a.u = ctx->s[1].as<double>(); a.x = ctx->s[2].as<double>(); a.status = ctx->s[3].as<int32_t>();
  if (um) { if (int rc = user_launch(ctx, um, rlu::K_ROLLOUT_OPEN, RL_GRID(a.B, 64), 1, 64, 0, &a, "k_rollout_open")) return rc; }
  else if (rll::launch_rollout_open(a, ctx->stream)) FAIL(-5, "model not compiled in");
  if (int rc = check_launch(ctx, "k_rollout_open", um ? 0 : 1)) return rc;
  DOWNSYNC(x, a.x, n * (N + 1) * B * 8); DOWNSYNC(status, a.status, (size_t)B * 4);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t ratilqr_rollout_closed_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t B, const double* xbar,
                                     const double* l, const double* L, double* x_new, double* u_new, int32_t* status) {
  if (!ctx) return -1;
  rll::CompArgs a;
  const rlu::Module* um = nullptr;
  if (int rc = fill_comp(ctx, desc, B, a, false, &um)) return rc;
  const size_t n = a.n, m = a.m, N = a.N;
  UP(ctx->s[0], xbar, n * (N + 1) * B * 8); UP(ctx->s[1], l, m * N * B * 8); UP(ctx->s[2], L, m * n * N * B * 8);
  CU(ctx->s[3].reserve(n * (N + 1) * B * 8)); CU(ctx->s[4].reserve(m * N * B * 8)); CU(ctx->s[5].reserve((size_t)B * 4));
  a.xbar = ctx->s[0].as<double>(); a.l = ctx->s[1].as<double>(); a.L = ctx->s[2].as<double>();
  a.x = ctx->s[3].as<double>(); a.u_new = ctx->s[4].as<double>(); a.status = ctx->s[5].as<int32_t>();
  if (um) { if (int rc = user_launch(ctx, um, rlu::K_ROLLOUT_CLOSED, RL_GRID(a.B, 64), 1, 64, 0, &a, "k_rollout_closed")) return rc; }
  else if (rll::launch_rollout_closed(a, ctx->stream)) FAIL(-5, "(model, cost) pair not compiled in");
  if (int rc = check_launch(ctx, "k_rollout_closed", um ? 0 : 1)) return rc;
  DOWNSYNC(x_new, a.x, n * (N + 1) * B * 8); DOWNSYNC(u_new, a.u_new, m * N * B * 8); DOWNSYNC(status, a.status, (size_t)B * 4);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t ratilqr_integrate_cost_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t B, const double* x,
                                     const double* u, double* cost, int32_t* status) {
  if (!ctx) return -1;
  rll::CompArgs a;
  const rlu::Module* um = nullptr;
  if (int rc = fill_comp(ctx, desc, B, a, false, &um)) return rc;
  const size_t n = a.n, m = a.m, N = a.N;
  UP(ctx->s[0], x, n * (N + 1) * B * 8); UP(ctx->s[1], u, m * N * B * 8);
  CU(ctx->s[2].reserve((size_t)B * 8)); CU(ctx->s[3].reserve((size_t)B * 4));
  a.x = ctx->s[0].as<double>(); a.u = ctx->s[1].as<double>(); a.cost = ctx->s[2].as<double>(); a.status = ctx->s[3].as<int32_t>();
  if (um) { if (int rc = user_launch(ctx, um, rlu::K_INTEGRATE_COST, RL_GRID(a.B, 64), 1, 64, 0, &a, "k_integrate_cost")) return rc; }
  else if (rll::launch_integrate_cost(a, ctx->stream)) FAIL(-5, "(model, cost) pair not compiled in");
  if (int rc = check_launch(ctx, "k_integrate_cost", um ? 0 : 1)) return rc;
  DOWNSYNC(cost, a.cost, (size_t)B * 8); DOWNSYNC(status, a.status, (size_t)B * 4);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t ratilqr_linearize_batch(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t B, const double* x,
                                const double* u, double* q, double* qv, double* Q, double* r, double* R, double* Pm,
                                double* A, double* Bm, int32_t* status) {
  if (!ctx) return -1;
  rll::CompArgs a;
  const rlu::Module* um = nullptr;
  if (int rc = fill_comp(ctx, desc, B, a, true, &um)) return rc;
  const size_t n = a.n, m = a.m, N = a.N;
  const size_t sz[8] = {(N + 1) * B, n * (N + 1) * B, n * n * (N + 1) * B, m * N * B, m * m * N * B, m * n * N * B, n * n * N * B, n * m * N * B};
  UP(ctx->s[0], x, n * (N + 1) * B * 8); UP(ctx->s[1], u, m * N * B * 8);
  for (int i = 0; i < 8; ++i) CU(ctx->s[2 + i].reserve(sz[i] * 8));
  CU(ctx->s[10].reserve((size_t)B * 4));
  CU(cudaMemsetAsync(ctx->s[10].p, 0, (size_t)B * 4, ctx->stream));
  a.x = ctx->s[0].as<double>(); a.u = ctx->s[1].as<double>();
  a.q = ctx->s[2].as<double>(); a.qv = ctx->s[3].as<double>(); a.Q = ctx->s[4].as<double>(); a.r = ctx->s[5].as<double>();
  a.R = ctx->s[6].as<double>(); a.Pm = ctx->s[7].as<double>(); a.A = ctx->s[8].as<double>(); a.Bm = ctx->s[9].as<double>();
  a.status = ctx->s[10].as<int32_t>();
  if (um) { if (int rc = user_launch(ctx, um, rlu::K_LINEARIZE, RL_GRID((size_t)a.B * (a.N + 1), 64), 1, 64, 0, &a, "k_linearize")) return rc; }
  else if (rll::launch_linearize(a, ctx->stream)) FAIL(-5, "(model, cost) pair not compiled in");
  if (int rc = check_launch(ctx, "k_linearize", um ? 0 : 1)) return rc;
  double* outs[8] = {q, qv, Q, r, R, Pm, A, Bm};
  for (int i = 0; i < 8; ++i) DOWNSYNC(outs[i], ctx->s[2 + i].p, sz[i] * 8);
  DOWNSYNC(status, a.status, (size_t)B * 4);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t ratilqr_riccati_batch(ratilqr_ctx* ctx, int32_t n_, int32_t m_, int32_t N_, int32_t B, int32_t optimise,
                              const double* q, const double* qv, const double* Q, const double* r, const double* R,
                              const double* Pm, const double* A, const double* Bm, const double* W, const double* theta,
                              double mu_min, double delta_0, double* mu, double* delta, double* L, double* dl,
                              double* s, double* sv, double* S, int32_t* status, int32_t* restarts) {
  return ratilqr_riccati_batch_tv(ctx, n_, m_, N_, B, optimise, q, qv, Q, r, R, Pm, A, Bm, W, 0, theta, mu_min, delta_0, mu,
                                  delta, L, dl, s, sv, S, status, restarts);
}

int32_t ratilqr_riccati_batch_tv(ratilqr_ctx* ctx, int32_t n_, int32_t m_, int32_t N_, int32_t B, int32_t optimise,
                                 const double* q, const double* qv, const double* Q, const double* r, const double* R,
                                 const double* Pm, const double* A, const double* Bm, const double* W,
                                 int32_t W_time_varying, const double* theta, double mu_min, double delta_0, double* mu,
                                 double* delta, double* L, double* dl, double* s, double* sv, double* S, int32_t* status,
                                 int32_t* restarts) {
  if (!ctx) return -1;
  if (B < 1 || N_ < 1 || !q || !qv || !Q || !r || !R || !Pm || !A || !Bm || !W || !theta || !mu || !delta || !L || !s || !sv || !S)
    FAIL(-1, "null argument");
  if (optimise && !dl) FAIL(-1, "dl output required when optimising");
  ctx->staged = false;
  CU(cudaSetDevice(ctx->device));
  const size_t n = n_, m = m_, N = N_;
  rlh::WPrep wp;
  if (!rlh::prep_W(n_, N_, W, W_time_varying, wp)) FAIL(-2, "W(k) is not positive definite");
  const size_t sz[8] = {(N + 1) * B, n * (N + 1) * B, n * n * (N + 1) * B, m * N * B, m * m * N * B, m * n * N * B, n * n * N * B, n * m * N * B};
  const double* ins[8] = {q, qv, Q, r, R, Pm, A, Bm};
  for (int i = 0; i < 8; ++i) UP(ctx->s[i], ins[i], sz[i] * 8);
  UP(ctx->d_W, wp.W.data(), wp.W.size() * 8); UP(ctx->d_Winv, wp.Winv.data(), wp.Winv.size() * 8);
  UP(ctx->d_detW, wp.detW.data(), wp.detW.size() * 8);
  UP(ctx->s[8], theta, (size_t)B * 8); UP(ctx->s[9], mu, (size_t)B * 8); UP(ctx->s[10], delta, (size_t)B * 8);
  if (optimise) { CU(ctx->s[11].reserve(m * n * N * B * 8)); CU(ctx->s[12].reserve(m * N * B * 8)); }
  else { UP(ctx->s[11], L, m * n * N * B * 8); if (dl) UP(ctx->s[12], dl, m * N * B * 8); }
  CU(ctx->s[13].reserve(sz[0] * 8)); CU(ctx->s[14].reserve(sz[1] * 8)); CU(ctx->s[15].reserve(sz[2] * 8));
  CU(ctx->d_status.reserve((size_t)B * 4)); CU(ctx->d_restarts.reserve((size_t)B * 4));
  rll::RiccatiArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n_; a.m = m_; a.N = N_; a.B = B; a.optimise = optimise;
  a.q = ctx->s[0].as<double>(); a.qv = ctx->s[1].as<double>(); a.Q = ctx->s[2].as<double>(); a.r = ctx->s[3].as<double>();
  a.R = ctx->s[4].as<double>(); a.Pm = ctx->s[5].as<double>(); a.A = ctx->s[6].as<double>(); a.Bm = ctx->s[7].as<double>();
  a.W = ctx->d_W.as<double>(); a.Winv = ctx->d_Winv.as<double>(); a.detW = ctx->d_detW.as<double>(); a.W_tv = W_time_varying ? 1 : 0;
  a.theta = ctx->s[8].as<double>(); a.mu_min = mu_min; a.delta_0 = delta_0;
  a.mu = ctx->s[9].as<double>(); a.delta = ctx->s[10].as<double>(); a.L = ctx->s[11].as<double>(); a.dl = ctx->s[12].as<double>();
  a.has_dl = dl != nullptr;
  a.s = ctx->s[13].as<double>(); a.sv = ctx->s[14].as<double>(); a.S = ctx->s[15].as<double>();
  a.status = ctx->d_status.as<int32_t>(); a.restarts = ctx->d_restarts.as<int32_t>();
  if (rll::launch_riccati(a, ctx->stream)) FAIL(-5, "(n, m) not compiled in");
  if (int rc = check_launch(ctx, "k_riccati")) return rc;
  DOWNSYNC(s, a.s, sz[0] * 8); DOWNSYNC(sv, a.sv, sz[1] * 8); DOWNSYNC(S, a.S, sz[2] * 8);
  DOWNSYNC(mu, a.mu, (size_t)B * 8); DOWNSYNC(delta, a.delta, (size_t)B * 8);
  if (optimise) { DOWNSYNC(L, a.L, m * n * N * B * 8); DOWNSYNC(dl, a.dl, m * N * B * 8); }
  DOWNSYNC(status, a.status, (size_t)B * 4); DOWNSYNC(restarts, a.restarts, (size_t)B * 4);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static int mc_rollout_internal(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t P, const double* xbar,
                               const double* l, const double* L, int32_t n_samples, const double* noise,
                               const ratilqr_noise_mixture* true_noise, uint64_t seed, double theta_risk, double* J,
                               double* stats, double* x_out) {
  if (!ctx) return -1;
  const rlu::Module* um = nullptr;
  if (const char* msg = check_desc_ctx(ctx, desc, false, &um)) FAIL(-1, msg);
  if (P < 1 || n_samples < 1 || !xbar || !l || !L) FAIL(-1, "bad arguments");
  if (desc->cost_params_count != 1 && desc->cost_params_count != P) FAIL(-1, "cost_params_count must be 1 or P");
  ctx->staged = false;
  CU(cudaSetDevice(ctx->device));
  const size_t n = desc->n, m = desc->m, N = desc->N, S = (size_t)n_samples * P;
  rlh::WPrep wp;
  if (!rlh::prep_W(desc->n, desc->N, desc->W, desc->W_time_varying, wp)) FAIL(-2, "W(k) is not positive definite");
  rll::McArgs a;
  memset(&a, 0, sizeof(a));
  a.model_id = desc->model_id; a.cost_id = desc->cost_id; a.N = desc->N; a.P = P; a.n_samples = n_samples;
  for (int i = 0; i < 8; ++i) a.mp[i] = i < desc->n_model_params ? desc->model_params[i] : 0.0;
  UP(ctx->d_cp, desc->cost_params, (size_t)desc->n_cost_params * desc->cost_params_count * 8);
  a.cp = ctx->d_cp.as<double>(); a.ncp = desc->n_cost_params; a.cp_count = desc->cost_params_count;
  UP(ctx->s[0], xbar, n * (N + 1) * P * 8); UP(ctx->s[1], l, m * N * P * 8); UP(ctx->s[2], L, m * n * N * P * 8);
  a.xbar = ctx->s[0].as<double>(); a.l = ctx->s[1].as<double>(); a.L = ctx->s[2].as<double>();
  if (noise) { UP(ctx->s[3], noise, n * N * S * 8); a.noise = ctx->s[3].as<double>(); }
  UP(ctx->s[4], wp.cholW.data(), wp.cholW.size() * 8);
  a.cholW = ctx->s[4].as<double>(); a.W_tv = desc->W_time_varying; a.seed = seed;
  if (int rc = upload_mixture(ctx, desc->n, true_noise, a.mix)) return rc;
  CU(ctx->s[5].reserve(S * 8)); a.J = ctx->s[5].as<double>();
  if (x_out) { CU(ctx->s[6].reserve(n * (N + 1) * S * 8)); a.x_out = ctx->s[6].as<double>(); }
  if (um) { if (int rc = user_launch(ctx, um, rlu::K_MC_ROLLOUT, RL_GRID(a.n_samples, 128), (unsigned)a.P, 128, 0, &a, "k_mc_rollout")) return rc; }
  else if (rll::launch_mc_rollout(a, ctx->stream)) FAIL(-5, "(model, cost) pair not compiled in");
  if (int rc = check_launch(ctx, "k_mc_rollout", um ? 0 : 1)) return rc;
  if (stats) {
    CU(ctx->s[7].reserve((size_t)P * 24));
    const size_t scr = rll::mc_stats_scratch_doubles(n_samples, P);  // > 0: chunked reductions (many samples per problem)
    if (scr) CU(ctx->s[8].reserve(scr * 8));
    rll::launch_mc_stats(a.J, n_samples, P, theta_risk, ctx->s[7].as<double>(), scr ? ctx->s[8].as<double>() : nullptr, ctx->stream);
    if (int rc = check_launch(ctx, "k_mc_stats", scr ? (theta_risk > 0.0 ? 5 : 3) : 1)) return rc;
    DOWNSYNC(stats, ctx->s[7].p, (size_t)P * 24);
  }
  DOWNSYNC(J, a.J, S * 8);
  DOWNSYNC(x_out, a.x_out, n * (N + 1) * S * 8);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t ratilqr_mc_rollout(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t P, const double* xbar,
                           const double* l, const double* L, int32_t n_samples, const double* noise, uint64_t seed,
                           double theta_risk, double* J, double* stats, double* x_out) {
  return mc_rollout_internal(ctx, desc, P, xbar, l, L, n_samples, noise, nullptr, seed, theta_risk, J, stats, x_out);
}

int32_t ratilqr_mc_rollout_true_model(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int32_t P, const double* xbar,
                                      const double* l, const double* L, int32_t n_samples,
                                      const ratilqr_noise_mixture* true_noise, uint64_t seed, double theta_risk,
                                      double* J, double* stats, double* x_out) {
  if (!ctx) return -1;
  if (!true_noise) FAIL(-1, "true_noise is null");
  return mc_rollout_internal(ctx, desc, P, xbar, l, L, n_samples, nullptr, true_noise, seed, theta_risk, J, stats, x_out);
}

static int pets_launch_costs(ratilqr_ctx* ctx, const rlu::Module* um, rll::PetsArgs& a) {
  if (um) {
    const unsigned th = a.particles >= 256 ? 256 : ((a.particles + 31) / 32) * 32;
    return user_launch(ctx, um, rlu::K_PETS_COSTS, (unsigned)a.C, 1, th, 0, &a, "k_pets_costs");
  }
  if (rll::launch_pets_costs(a, ctx->stream)) FAIL(-5, "(model, cost) pair not compiled in");
  return 0;
}

static int pets_fill(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_generative_desc* gen,
                     const double* x0, int C, int particles, rll::PetsArgs& a, const rlu::Module** um) {
  if (const char* msg = check_desc_ctx(ctx, desc, false, um)) FAIL(-1, msg);
  if (C < 1 || particles < 1 || !x0) FAIL(-1, "bad arguments");
  ctx->staged = false;
  CU(cudaSetDevice(ctx->device));
  rlh::WPrep wp;
  if (!rlh::prep_W(desc->n, desc->N, desc->W, 0, wp)) FAIL(-2, "W is not positive definite");
  memset(&a, 0, sizeof(a));
  a.model_id = desc->model_id; a.cost_id = desc->cost_id; a.N = desc->N; a.C = C; a.particles = particles;
  for (int i = 0; i < 8; ++i) a.mp[i] = i < desc->n_model_params ? desc->model_params[i] : 0.0;
  a.n_mp = desc->n_model_params;
  a.noise_kind = gen ? gen->noise_kind : 0;
  a.noise_scale = gen ? gen->noise_scale : 1.0;
  a.n_ens = gen && gen->n_ensemble > 1 ? gen->n_ensemble : 1;
  if (gen && gen->ensemble_params && a.n_ens > 1) {
    UP(ctx->s[8], gen->ensemble_params, (size_t)a.n_ens * desc->n_model_params * 8);
    a.ens_params = ctx->s[8].as<double>();
  }
  UP(ctx->d_cp, desc->cost_params, (size_t)desc->n_cost_params * 8);
  a.cp = ctx->d_cp.as<double>();
  UP(ctx->s[9], x0, (size_t)desc->n * 8);
  a.x0 = ctx->s[9].as<double>();
  UP(ctx->s[10], wp.cholW.data(), wp.cholW.size() * 8);
  a.cholW = ctx->s[10].as<double>();
  if (int rc = upload_mixture(ctx, desc->n, (gen && gen->use_true_model) ? gen->true_model : nullptr, a.mix)) return rc;
  return 0;
}

int32_t ratilqr_pets_costs(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_generative_desc* gen,
                           const double* x0, const double* controls, int32_t C, int32_t particles, const double* noise,
                           uint64_t seed, double* cost) {
  if (!ctx) return -1;
  if (!controls || !cost) FAIL(-1, "null argument");
  rll::PetsArgs a;
  const rlu::Module* um = nullptr;
  if (int rc = pets_fill(ctx, desc, gen, x0, C, particles, a, &um)) return rc;
  const size_t n = desc->n, m = desc->m, N = desc->N;
  UP(ctx->s[0], controls, m * N * C * 8);
  a.controls = ctx->s[0].as<double>();
  if (noise) { UP(ctx->s[1], noise, n * N * (size_t)particles * C * 8); a.noise = ctx->s[1].as<double>(); }
  a.seed = seed; a.stream_offset = 0;
  CU(ctx->s[2].reserve((size_t)C * 8));
  a.cost = ctx->s[2].as<double>();
  if (int rc = pets_launch_costs(ctx, um, a)) return rc;
  if (int rc = check_launch(ctx, "k_pets_costs", um ? 0 : 1)) return rc;
  DOWNSYNC(cost, a.cost, (size_t)C * 8);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t ratilqr_pets_refit(ratilqr_ctx* ctx, int32_t m_, int32_t N_, int32_t C, int32_t num_elite, double smoothing,
                           const double* controls, const double* cost, double* mu, double* Sigma, int32_t* elite_idx) {
  if (!ctx) return -1;
  if (!controls || !cost || !mu || !Sigma || num_elite < 1 || num_elite > C) FAIL(-1, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  const size_t m = m_, N = N_;
  UP(ctx->s[0], controls, m * N * C * 8); UP(ctx->s[2], cost, (size_t)C * 8);
  UP(ctx->s[3], mu, m * N * 8); UP(ctx->s[4], Sigma, m * m * N * 8);
  CU(ctx->s[5].reserve((size_t)num_elite * 4));
  rll::launch_pets_refit(m_, N_, C, num_elite, smoothing, ctx->s[0].as<double>(), ctx->s[2].as<double>(),
                         ctx->s[3].as<double>(), ctx->s[4].as<double>(), ctx->s[5].as<int32_t>(), nullptr, ctx->stream);
  if (int rc = check_launch(ctx, "k_pets_refit", 2)) return rc;
  DOWNSYNC(mu, ctx->s[3].p, m * N * 8); DOWNSYNC(Sigma, ctx->s[4].p, m * m * N * 8);
  DOWNSYNC(elite_idx, ctx->s[5].p, (size_t)num_elite * 4);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t ratilqr_pets_solve(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_generative_desc* gen,
                           const double* x0, int32_t C, int32_t particles, int32_t num_elite, int32_t iter_max,
                           double smoothing, const double* z_inject, const double* noise, uint64_t seed, double* mu,
                           double* Sigma) {
  if (!ctx) return -1;
  if (!mu || !Sigma || num_elite < 2 || num_elite > C || iter_max < 0) FAIL(-1, "bad arguments");
  rll::PetsArgs a;
  const rlu::Module* um = nullptr;
  if (int rc = pets_fill(ctx, desc, gen, x0, C, particles, a, &um)) return rc;
  const size_t n = desc->n, m = desc->m, N = desc->N;
  const size_t zsz = m * N * C, nsz = n * N * (size_t)particles * C;
  if (z_inject) UP(ctx->s[1], z_inject, zsz * iter_max * 8);
  if (noise) UP(ctx->s[11], noise, nsz * iter_max * 8);
  CU(ctx->s[0].reserve(zsz * 8)); CU(ctx->s[2].reserve((size_t)C * 8));
  UP(ctx->s[3], mu, m * N * 8); UP(ctx->s[4], Sigma, m * m * N * 8);
  CU(ctx->s[5].reserve((size_t)num_elite * 4)); CU(ctx->s[6].reserve(4));
  CU(cudaMemsetAsync(ctx->s[6].p, 0, 4, ctx->stream));
  a.controls = ctx->s[0].as<double>(); a.cost = ctx->s[2].as<double>(); a.seed = seed;
  for (int it = 0; it < iter_max; ++it) {  // step! pets.jl:193-245, all on the device
    rll::launch_pets_sample(desc->m, desc->N, C, ctx->s[3].as<double>(), ctx->s[4].as<double>(),
                            z_inject ? ctx->s[1].as<double>() + zsz * it : nullptr, seed ^ 0x5bd1e995u,
                            (uint64_t)it * (uint64_t)C, ctx->s[0].as<double>(), ctx->s[6].as<int32_t>(), ctx->stream);
    a.noise = noise ? ctx->s[11].as<double>() + nsz * it : nullptr;
    a.stream_offset = (uint64_t)it * (uint64_t)C * (uint64_t)particles;
    if (int rc = pets_launch_costs(ctx, um, a)) return rc;
    rll::launch_pets_refit(desc->m, desc->N, C, num_elite, smoothing, a.controls, a.cost, ctx->s[3].as<double>(),
                           ctx->s[4].as<double>(), ctx->s[5].as<int32_t>(), nullptr, ctx->stream);
    if (int rc = check_launch(ctx, "pets step", um ? 3 : 4)) return rc;
  }
  int32_t err = 0;
  CU(cudaMemcpyAsync(&err, ctx->s[6].p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DOWNSYNC(mu, ctx->s[3].p, m * N * 8); DOWNSYNC(Sigma, ctx->s[4].p, m * m * N * 8);
  CU(cudaStreamSynchronize(ctx->stream));
  if (err) FAIL(-4, "a Sigma_t is not positive definite (PosDefException in MvNormal, pets.jl:212)");
  return 0;
}

// ---- RAT iLQR for a fleet of problems (cross_entropy_bilevel_optimization.jl:364-415) ------------------------
static int ce_solve_fleet_block(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                                const ratilqr_ce_opts* ce, int32_t P, long long p0, const double* x0, int32_t x0_count,
                                const double* u_init, int32_t u_count, double kl_bound, const double* z_inject,
                                int64_t nz, uint64_t seed, double* mu_init, double* sigma_init, double* theta_opt,
                                double* value, double* theta_min, double* theta_max, double* mu, double* sigma,
                                int64_t* nz_used, int32_t* rounds_out, ratilqr_ileqg_out* final_out,
                                rll::MpcArgs* sink = nullptr, bool sharded = false) {
  if (!ctx) return -1;
  // sharded: ONE problem whose theta population is split over the ranks of ctx->comm (cross_entropy...jl:180-193 is the
  // reference's fan-out): every rank draws the same population, solves its block, all-gathers (value, status) on device
  // buffers and runs the elite selection redundantly, so the CE state stays replicated without a broadcast
  if (sharded && (P != 1 || !ctx->comm || ctx->world < 2 || (ce && ce->num_samples < ctx->world))) FAIL(-1, "sharded CE needs one problem, an attached communicator and num_samples >= world size");
  if (!ce || P < 1 || !mu_init || !sigma_init || ((!theta_opt || !value) && !sink)) FAIL(-1, "bad arguments");
  if (!(kl_bound >= 0)) FAIL(-3, "KL Divergence Bound must be non-negative");  // :368
  if (ce->num_samples < 1 || ce->num_elite < 1 || ce->num_elite > ce->num_samples || ce->iter_max < 1) FAIL(-3, "bad CE options");
  const int S = ce->num_samples;
  // sub-fleet blocks run on fresh host threads whose current device is 0: bind before any allocation / copy
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // per-problem CE state on the device (scratch slots s[0..12] are free during a solve)
  rll::CeFleet c;
  memset(&c, 0, sizeof(c));
  c.P = P; c.S = S; c.num_elite = ce->num_elite; c.iter_max = ce->iter_max; c.use_theta_max = ce->use_theta_max;
  c.lambda = ce->lambda; c.kl = kl_bound; c.nz = nz; c.seed = seed; c.p0 = p0;
  ratilqr_batch_in in;
  in.P = P; in.K = S; in.x0 = x0; in.x0_count = x0_count; in.u_init = u_init; in.u_count = u_count; in.theta = nullptr;
  int rc = 0, rounds = 0;
  const char* esf = getenv("RATILQR_FLEET_SORT");  // 0 disables the iteration-count ordering of slots (A/B runs)
  const bool sort_fleet = !(esf && esf[0] == '0');
  const size_t Pb = (size_t)P * 8;
  std::vector<double> init_tmin(P, HUGE_VAL), zeros(P, 0.0);
  std::vector<int32_t> ones(P, 1), izeros(P, 0);
  std::vector<long long> lzeros(P, 0);
  const int sh_lo = sharded ? (int)((long long)S * ctx->rank / ctx->world) : 0;
  const int sh_cnt = sharded ? (int)((long long)S * (ctx->rank + 1) / ctx->world) - sh_lo : S;
  if (kl_bound > 0) {
    in.K = sh_cnt;
    if ((rc = stage_internal(ctx, desc, opts, &in, 0, true))) return rc;
    in.K = S;
    UP(ctx->s[0], mu_init, Pb); UP(ctx->s[1], sigma_init, Pb);
    UP(ctx->s[2], mu_init, Pb); UP(ctx->s[3], sigma_init, Pb);          // initialize! :133-138: mu, sigma <- *_init
    UP(ctx->s[4], init_tmin.data(), Pb); UP(ctx->s[5], zeros.data(), Pb);  // theta_min = Inf, theta_max = 0
    CU(ctx->s[6].reserve(Pb)); CU(ctx->s[7].reserve(Pb));
    UP(ctx->s[8], lzeros.data(), Pb); UP(ctx->s[9], ones.data(), (size_t)P * 4);  // cursor = 0 ; iter = 1
    UP(ctx->s[10], ones.data(), (size_t)P * 4); UP(ctx->s[11], izeros.data(), (size_t)P * 4);  // active ; err
    CU(ctx->s[12].reserve(8));
    CU(ctx->s[14].reserve((size_t)P * ce->num_elite * 8));  // elites in rank order (CTA-per-problem refit)
    if (z_inject) { UP(ctx->s[13], z_inject, (size_t)P * nz * 8); c.z = ctx->s[13].as<double>(); }
  } else {
    in.K = 1;
    UP(ctx->s[2], mu_init, Pb); UP(ctx->s[3], sigma_init, Pb);
    UP(ctx->s[0], mu_init, Pb); UP(ctx->s[1], sigma_init, Pb);
    UP(ctx->s[4], zeros.data(), Pb); UP(ctx->s[5], zeros.data(), Pb);
    CU(ctx->s[6].reserve(Pb)); CU(ctx->s[7].reserve(Pb));
    UP(ctx->s[8], lzeros.data(), Pb); UP(ctx->s[9], ones.data(), (size_t)P * 4);
    UP(ctx->s[10], ones.data(), (size_t)P * 4); UP(ctx->s[11], izeros.data(), (size_t)P * 4);
    CU(ctx->s[12].reserve(8));
  }
  c.mu_init = ctx->s[0].as<double>(); c.sigma_init = ctx->s[1].as<double>(); c.mu = ctx->s[2].as<double>();
  c.sigma = ctx->s[3].as<double>(); c.theta_min = ctx->s[4].as<double>(); c.theta_max = ctx->s[5].as<double>();
  c.theta_opt = ctx->s[6].as<double>(); c.value_out = ctx->s[7].as<double>(); c.cursor = ctx->s[8].as<long long>();
  c.iter = ctx->s[9].as<int32_t>(); c.active = ctx->s[10].as<int32_t>(); c.err = ctx->s[11].as<int32_t>();
  c.n_active = ctx->s[12].as<int32_t>();
  if (kl_bound > 0) {
    c.theta = ctx->d_theta.as<double>(); c.value = ctx->sp.value; c.status = ctx->sp.status;
    if (sharded) {  // the population lives in full-size arrays; the solve kernel sees this rank's block of theta
      CU(ctx->d_sh[0].reserve((size_t)S * 8)); CU(ctx->d_sh[1].reserve((size_t)S * 8)); CU(ctx->d_sh[2].reserve((size_t)S * 4));
      c.theta = ctx->d_sh[0].as<double>(); c.value = ctx->d_sh[1].as<double>(); c.status = ctx->d_sh[2].as<int32_t>();
      ctx->sp.theta = c.theta + sh_lo;
    }
    ctx->sp.active = c.active;
    if (sort_fleet && !ctx->coop && P >= 64 && (int)ctx->fleet_key.size() == P) {  // previous call on a fleet of this size
      if ((rc = apply_slot_order(ctx, ctx->fleet_key, P, S))) return rc;
    }
    while (true) {  // one round = draw -> batched solve -> per-problem CE logic; one host sync per round
      rll::launch_ce_draw(c, st);
      if ((rc = check_launch(ctx, "k_ce_draw"))) return rc;
      if ((rc = run_internal(ctx, 1, nullptr))) return rc;
      if (sharded && (rc = shard_allgather(ctx, S, sh_lo, sh_cnt, ctx->sp.value, ctx->sp.status, ctx->d_sh[1].as<double>(), ctx->d_sh[2].as<int32_t>()))) return rc;
      CU(cudaMemsetAsync(c.n_active, 0, 4, st));
      rll::launch_ce_update(c, ctx->s[14].as<double>(), st);
      if ((rc = check_launch(ctx, "k_ce_update"))) return rc;
      int32_t n_active = 0;
      CU(cudaMemcpyAsync(&n_active, c.n_active, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      rounds++;
      if (n_active == 0) {  // keep the work profile of the last round for the final solve and for the next MPC step
        if (sort_fleet && !ctx->coop && P >= 64) { if ((rc = resort_slots_by_last_iters(ctx, P, S))) return rc; }
        break;
      }
      if (rounds > 100000) FAIL(-6, "CE redraw loop does not terminate (the reference would spin forever here)");
      if (sort_fleet && !ctx->coop && P >= 64) { if ((rc = resort_slots_by_last_iters(ctx, P, S))) return rc; }
    }
  }
  // (the per-problem work of the last round, ctx->fleet_key, orders the final solve too)
  const bool order_final = kl_bound > 0 && sort_fleet && !ctx->coop && P >= 64 && (int)ctx->fleet_key.size() == P;
  // final solve at theta_opt with the retry rule (:390-414); B = P instances, theta on the device
  in.K = 1;
  const int fwant = (final_out ? ((final_out->x ? 1 : 0) | (final_out->l ? 2 : 0) | (final_out->L ? 4 : 0)) : 0) | (sink ? 2 : 0);
  if ((rc = stage_internal(ctx, desc, opts, &in, final_out && final_out->eps_hist ? final_out->eps_hist_cap : 0, true, fwant))) return rc;
  rll::launch_ce_pick_theta(c, ctx->d_theta.as<double>(), st);
  if ((rc = check_launch(ctx, "k_ce_pick_theta"))) return rc;
  ctx->sp.active = c.active;
  if (order_final) { if ((rc = apply_slot_order(ctx, ctx->fleet_key, P, 1))) return rc; }
  int final_rounds = 0;
  while (true) {
    if ((rc = run_internal(ctx, 1, nullptr))) return rc;
    CU(cudaMemsetAsync(c.n_active, 0, 4, st));
    rll::launch_ce_final_update(c, ctx->d_theta.as<double>(), ctx->sp.value, ctx->sp.status, st);
    if ((rc = check_launch(ctx, "k_ce_final_update"))) return rc;
    int32_t n_active = 0;
    CU(cudaMemcpyAsync(&n_active, c.n_active, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (n_active == 0) break;
    if (++final_rounds > 10000) FAIL(-6, "final solve retry loop does not terminate");
  }
  ctx->sp.active = nullptr;
  if (sink) {  // receding-horizon driver: true-system step and plan shift on the device, nothing comes back to the host
    if ((rc = device_plan(ctx, &sink->plan))) return rc;
    sink->theta_opt = c.theta_opt; sink->value = c.value_out;
    if (rll::launch_mpc_advance(*sink, st)) FAIL(-5, "this model is not compiled in");
    if ((rc = check_launch(ctx, "k_mpc_advance"))) return rc;
  }
  DOWNSYNC(theta_opt, c.theta_opt, Pb); DOWNSYNC(value, c.value_out, Pb);
  DOWNSYNC(mu_init, c.mu_init, Pb); DOWNSYNC(sigma_init, c.sigma_init, Pb);
  DOWNSYNC(mu, c.mu, Pb); DOWNSYNC(sigma, c.sigma, Pb);
  if (kl_bound > 0) { DOWNSYNC(theta_min, c.theta_min, Pb); DOWNSYNC(theta_max, c.theta_max, Pb); }
  std::vector<long long> cur(P);
  std::vector<int32_t> err(P);
  CU(cudaMemcpyAsync(cur.data(), c.cursor, Pb, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(err.data(), c.err, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (kl_bound <= 0) { for (int p = 0; p < P; ++p) { if (theta_min) theta_min[p] = 0.0; if (theta_max) theta_max[p] = 0.0; } }  // :408
  if (nz_used) for (int p = 0; p < P; ++p) nz_used[p] = cur[p];
  if (rounds_out) *rounds_out = rounds;
  for (int p = 0; p < P; ++p) if (err[p] == 1) FAIL(-5, "a problem exhausted its injected normal stream (or needed > 1e6 draws)");
  if (final_out) return fetch_internal(ctx, final_out);
  return 0;
}

// Problems are independent, but each one's CE loop is a chain of sequential rounds, and a round of P * num_samples
// instances that does not fill the GPU several times over ends in a long under-occupied tail (one thread per instance:
// the slowest lanes finish alone).  So a large fleet is cut into contiguous blocks that run the SAME per-problem
// algorithm concurrently, each on its own stream with its own workspace and host thread: while one block drains its
// round, the others keep the SMs busy.  Results are identical to the one-block run (per-problem state, Philox streams
// indexed by the global problem number).  RATILQR_FLEET_SPLIT=<k> overrides the block count (1 = off).
static int fleet_blocks(const ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, int P, int S) {
  if (const char* e = getenv("RATILQR_FLEET_SPLIT")) { int k = atoi(e); return k < 1 ? 1 : (k > 8 ? 8 : k); }
  if (desc && desc->n > 6) return 1;                   // warp-cooperative kernel: already latency-oriented
  if ((long long)P * S < 56832) return 1;              // less than one resident wave (148 SMs x 384) in total
  (void)ctx;
  return 3;  // A/B on the C5 fleet (profiles/r01_fleet_split_order_ab.jsonl): 1: 281 ms, 2: 279, 3: 268, 4: 277, 6+: worse
}

int32_t ratilqr_ce_solve_fleet(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                               const ratilqr_ce_opts* ce, int32_t P, const double* x0, int32_t x0_count,
                               const double* u_init, int32_t u_count, double kl_bound, const double* z_inject,
                               int64_t nz, uint64_t seed, double* mu_init, double* sigma_init, double* theta_opt,
                               double* value, double* theta_min, double* theta_max, double* mu, double* sigma,
                               int64_t* nz_used, int32_t* rounds_out, ratilqr_ileqg_out* final_out) {
  if (!ctx) return -1;
  const int K = (ce && desc && P > 1) ? std::min(fleet_blocks(ctx, desc, P, ce->num_samples), (int)P) : 1;
  if (K <= 1)
    return ce_solve_fleet_block(ctx, desc, opts, ce, P, 0, x0, x0_count, u_init, u_count, kl_bound, z_inject, nz, seed, mu_init,
                                sigma_init, theta_opt, value, theta_min, theta_max, mu, sigma, nz_used, rounds_out, final_out);
  if (desc->cost_params_count != 1 && desc->cost_params_count != P) FAIL(-1, "cost_params_count must be 1 or P");
  while ((int)ctx->children.size() < K - 1) {
    ratilqr_ctx* ch = nullptr;
    if (int rc = ratilqr_create(&ch, ctx->device)) FAIL(rc, "could not create a sub-fleet context");
    ch->parent = ctx;
    ctx->children.push_back(ch);
  }
  const size_t n = desc->n, m = desc->m, N = desc->N;
  std::vector<int> rcs(K, 0), rounds(K, 0);
  std::vector<std::thread> th;
  auto run_block = [&](int b) {
    const long long lo = (long long)P * b / K, hi = (long long)P * (b + 1) / K;
    const int Pb = (int)(hi - lo);
    ratilqr_ctx* c = b == 0 ? ctx : ctx->children[b - 1];
    ratilqr_problem_desc d = *desc;
    if (desc->cost_params_count == P) { d.cost_params = desc->cost_params + (size_t)lo * desc->n_cost_params; d.cost_params_count = Pb; }
    ratilqr_ileqg_out fo;
    if (final_out) {
      fo = *final_out;
      const size_t cap = (size_t)final_out->eps_hist_cap;
      if (fo.x) fo.x += n * (N + 1) * lo;
      if (fo.l) fo.l += m * N * lo;
      if (fo.L) fo.L += m * n * N * lo;
      if (fo.value) fo.value += lo;
      if (fo.status) fo.status += lo;
      if (fo.iters) fo.iters += lo;
      if (fo.trials) fo.trials += lo;
      if (fo.restarts) fo.restarts += lo;
      if (fo.mu) fo.mu += lo;
      if (fo.d_current) fo.d_current += lo;
      if (fo.eps_hist) fo.eps_hist += 2 * cap * lo;
    }
    auto off = [&](double* q) { return q ? q + lo : nullptr; };
    rcs[b] = ce_solve_fleet_block(c, &d, opts, ce, Pb, lo, x0_count == P ? x0 + n * lo : x0, x0_count == P ? Pb : x0_count,
                                  u_count == P ? u_init + m * N * lo : u_init, u_count == P ? Pb : u_count, kl_bound,
                                  z_inject ? z_inject + (size_t)lo * nz : nullptr, nz, seed, mu_init + lo, sigma_init + lo,
                                  theta_opt + lo, value + lo, off(theta_min), off(theta_max), off(mu), off(sigma),
                                  nz_used ? nz_used + lo : nullptr, &rounds[b], final_out ? &fo : nullptr);
  };
  if (!ce || !mu_init || !sigma_init || !theta_opt || !value) FAIL(-1, "bad arguments");
  if ((x0_count != 1 && x0_count != P) || (u_count != 1 && u_count != P)) FAIL(-1, "x0_count/u_count must be 1 or P");
  int spawned = 1;
  try {
    for (int b = 1; b < K; ++b) { th.emplace_back(run_block, b); ++spawned; }
  } catch (...) {  // no more host threads: the remaining blocks run on this one, after block 0
  }
  run_block(0);
  for (int b = spawned; b < K; ++b) run_block(b);
  for (auto& t : th) t.join();
  int rmax = 0;
  for (int b = 0; b < K; ++b) {
    rmax = std::max(rmax, rounds[b]);
    if (b > 0) ctx->launches += ctx->children[b - 1]->launches, ctx->children[b - 1]->launches = 0;
  }
  if (rounds_out) *rounds_out = rmax;
  for (int b = 0; b < K; ++b)
    if (rcs[b]) { if (b > 0) ctx->err = ctx->children[b - 1]->err; return rcs[b]; }
  return 0;
}

// ---- receding-horizon RAT iLQR for a fleet, `steps` MPC steps without leaving the device --------------
// One block of problems: plan (CE loop + final solve, ce_solve_fleet_block with device-resident x / warm start), then
// k_mpc_advance applies l_0 to the true system, draws the disturbance, shifts the plan.  mu_init / sigma_init persist
// across the steps like the fields of the reference's solver struct (cross_entropy...jl:66-68, 297-301).
static int mpc_block(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                     const ratilqr_ce_opts* ce, int P, long long p0, int Ptot, const double* x0, const double* u_init,
                     int u_count, double kl_bound, const double* z_inject, int64_t nz, uint64_t seed,
                     const ratilqr_mpc_opts* mo, double* mu_init, double* sigma_init, double* x_traj, double* u_traj,
                     double* theta_traj, double* value_traj, float* step_ms, int32_t* rounds_total) {
  CU(cudaSetDevice(ctx->device));
  const size_t n = desc->n, m = desc->m, N = desc->N, steps = mo->steps;
  cudaStream_t st = ctx->stream;
  rlh::WPrep wp;
  if (!rlh::prep_W(desc->n, desc->N, desc->W, desc->W_time_varying, wp)) FAIL(-2, "W(k) is not positive definite");
  // device-resident state: x (n*P) in d_x0, warm start (m*N*P) in d_u
  UP(ctx->d_x0, x0, n * P * 8);
  {
    std::vector<double> u0(m * N * P);
    for (int p = 0; p < P; ++p) memcpy(u0.data() + (size_t)p * m * N, u_init + (u_count > 1 ? (size_t)p * m * N : 0), m * N * 8);
    UP(ctx->d_u, u0.data(), u0.size() * 8);  // pageable source: staged before the call returns
  }
  CU(ctx->d_mpc[0].reserve(n * (steps + 1) * P * 8)); CU(ctx->d_mpc[1].reserve(m * steps * P * 8));
  CU(ctx->d_mpc[2].reserve(steps * P * 8)); CU(ctx->d_mpc[3].reserve(steps * P * 8));
  if (mo->noise) UP(ctx->d_mpc[4], mo->noise, n * steps * P * 8);
  UP(ctx->d_mpc[5], wp.cholW.data(), n * n * 8);
  CU(ctx->d_mpc[6].reserve(8));
  CU(cudaMemsetAsync(ctx->d_mpc[6].p, 0, 4, st));
  CU(cudaMemcpy2DAsync(ctx->d_mpc[0].p, n * (steps + 1) * 8, ctx->d_x0.p, n * 8, n * 8, P, cudaMemcpyDeviceToDevice, st));  // x_traj[:, 0, p] = x0
  rll::MpcArgs a;
  memset(&a, 0, sizeof(a));
  a.model_id = desc->model_id; a.N = desc->N; a.P = P; a.steps = mo->steps; a.p0 = p0;
  for (int i = 0; i < 8; ++i) a.mp[i] = i < desc->n_model_params ? desc->model_params[i] : 0.0;
  a.x = ctx->d_x0.as<double>(); a.u_init = ctx->d_u.as<double>();
  a.noise = mo->noise ? ctx->d_mpc[4].as<double>() : nullptr;
  a.cholW = ctx->d_mpc[5].as<double>(); a.seed = mo->noise_seed;
  if (int rc = upload_mixture(ctx, desc->n, mo->true_noise, a.mix)) return rc;
  a.x_traj = ctx->d_mpc[0].as<double>(); a.u_traj = ctx->d_mpc[1].as<double>();
  a.theta_traj = ctx->d_mpc[2].as<double>(); a.value_traj = ctx->d_mpc[3].as<double>(); a.err = ctx->d_mpc[6].as<int32_t>();
  int rounds_sum = 0;
  for (int t = 0; t < mo->steps; ++t) {
    const auto t0 = std::chrono::steady_clock::now();
    a.t = t;
    ctx->inputs_on_device = true;
    int rounds = 0;
    const double* zt = z_inject ? z_inject + ((size_t)t * Ptot + (size_t)p0) * nz : nullptr;
    int rc = ce_solve_fleet_block(ctx, desc, opts, ce, P, p0, nullptr, P, nullptr, P, kl_bound, zt, nz, seed + (uint64_t)t, mu_init,
                                  sigma_init, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &rounds, nullptr, &a);
    ctx->inputs_on_device = false;
    if (rc) return rc;
    rounds_sum += rounds;
    if (step_ms) step_ms[t] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  int32_t err = 0;
  CU(cudaMemcpyAsync(&err, ctx->d_mpc[6].p, 4, cudaMemcpyDeviceToHost, st));
  DOWNSYNC(x_traj, ctx->d_mpc[0].p, n * (steps + 1) * P * 8); DOWNSYNC(u_traj, ctx->d_mpc[1].p, m * steps * P * 8);
  DOWNSYNC(theta_traj, ctx->d_mpc[2].p, steps * P * 8); DOWNSYNC(value_traj, ctx->d_mpc[3].p, steps * P * 8);
  CU(cudaStreamSynchronize(st));
  if (rounds_total) *rounds_total = rounds_sum;
  if (err) FAIL(-7, "the true system left the domain of its model (DomainError) during the receding-horizon run");
  return 0;
}

int32_t ratilqr_mpc_fleet_run(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                              const ratilqr_ce_opts* ce, const ratilqr_mpc_opts* mo, int32_t P, const double* x0,
                              const double* u_init, int32_t u_count, double kl_bound, const double* z_inject, int64_t nz,
                              uint64_t seed, double* mu_init, double* sigma_init, double* x_traj, double* u_traj,
                              double* theta_traj, double* value_traj, float* step_ms, int32_t* rounds_total) {
  if (!ctx) return -1;
  if (!desc || !ce || !mo || P < 1 || mo->steps < 1 || !x0 || !u_init || !mu_init || !sigma_init) FAIL(-1, "bad arguments");
  if (u_count != 1 && u_count != P) FAIL(-1, "u_count must be 1 or P");
  if (desc->cost_params_count != 1 && desc->cost_params_count != P) FAIL(-1, "cost_params_count must be 1 or P");
  if (desc->model_id >= RATILQR_MODEL_USER_BASE) FAIL(-5, "user models: drive ratilqr_ce_solve_fleet from the host (the true-system step is compiled for registered models)");
  const int K = std::min(fleet_blocks(ctx, desc, P, ce->num_samples), (int)P);
  while ((int)ctx->children.size() < K - 1) {
    ratilqr_ctx* ch = nullptr;
    if (int rc = ratilqr_create(&ch, ctx->device)) FAIL(rc, "could not create a sub-fleet context");
    ch->parent = ctx;
    ctx->children.push_back(ch);
  }
  const size_t n = desc->n, m = desc->m, N = desc->N, steps = mo->steps;
  std::vector<int> rcs(K, 0), rounds(K, 0);
  std::vector<std::vector<float>> ms(K, std::vector<float>(steps, 0.f));
  auto run_block = [&](int b) {
    const long long lo = (long long)P * b / K, hi = (long long)P * (b + 1) / K;
    const int Pb = (int)(hi - lo);
    ratilqr_ctx* c = b == 0 ? ctx : ctx->children[b - 1];
    ratilqr_problem_desc d = *desc;
    if (desc->cost_params_count == P) { d.cost_params = desc->cost_params + (size_t)lo * desc->n_cost_params; d.cost_params_count = Pb; }
    ratilqr_mpc_opts mb = *mo;
    if (mo->noise) mb.noise = mo->noise + n * steps * lo;
    auto off = [&](double* q, size_t per) { return q ? q + per * lo : nullptr; };
    rcs[b] = mpc_block(c, &d, opts, ce, Pb, lo, P, x0 + n * lo, u_count == P ? u_init + m * N * lo : u_init, u_count, kl_bound,
                       z_inject, nz, seed, &mb, mu_init + lo, sigma_init + lo, off(x_traj, n * (steps + 1)), off(u_traj, m * steps),
                       off(theta_traj, steps), off(value_traj, steps), ms[b].data(), &rounds[b]);
  };
  std::vector<std::thread> th;
  int spawned = 1;
  try {
    for (int b = 1; b < K; ++b) { th.emplace_back(run_block, b); ++spawned; }
  } catch (...) {
  }
  run_block(0);
  for (int b = spawned; b < K; ++b) run_block(b);
  for (auto& t : th) t.join();
  int rsum = 0;
  for (int b = 0; b < K; ++b) {
    rsum = std::max(rsum, rounds[b]);
    if (b > 0) ctx->launches += ctx->children[b - 1]->launches, ctx->children[b - 1]->launches = 0;
  }
  if (step_ms) for (size_t t = 0; t < steps; ++t) { float mx = 0.f; for (int b = 0; b < K; ++b) mx = std::max(mx, ms[b][t]); step_ms[t] = mx; }
  if (rounds_total) *rounds_total = rsum;
  for (int b = 0; b < K; ++b)
    if (rcs[b]) { if (b > 0) ctx->err = ctx->children[b - 1]->err; return rcs[b]; }
  return 0;
}

int32_t ratilqr_ce_solve(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                         const ratilqr_ce_opts* ce, const double* x0, const double* u_init, double kl_bound,
                         const double* z_inject, int64_t nz, uint64_t seed, double* mu_init, double* sigma_init,
                         double* theta_opt, double* value, double* theta_min, double* theta_max, double* mu, double* sigma,
                         int64_t* nz_used, int32_t* rounds_out, ratilqr_ileqg_out* final_out) {
  return ratilqr_ce_solve_fleet(ctx, desc, opts, ce, 1, x0, 1, u_init, 1, kl_bound, z_inject, nz, seed, mu_init, sigma_init,
                                theta_opt, value, theta_min, theta_max, mu, sigma, nz_used, rounds_out, final_out);
}

// ---- RAT iLQR++ for a fleet of problems (nelder_mead_bilevel_optimization.jl:276-352) -------------------------------
int32_t ratilqr_nm_solve_fleet(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                               const ratilqr_nm_opts* nm, int32_t P, const double* x0, int32_t x0_count,
                               const double* u_init, int32_t u_count, double kl_bound, double* theta_high_init,
                               double* theta_low_init, double* c_high, double* c_low, int32_t* has_c, double* theta_opt,
                               double* value, int32_t* nm_iters, int32_t* n_evals, ratilqr_ileqg_out* final_out) {
  if (!ctx) return -1;
  if (!nm || P < 1 || !theta_high_init || !theta_low_init || !c_high || !c_low || !has_c || !theta_opt || !value) FAIL(-1, "bad arguments");
  if (!(kl_bound >= 0)) FAIL(-3, "KL Divergence Bound must be non-negative");  // :280
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  rll::NmFleet c;
  memset(&c, 0, sizeof(c));
  c.P = P; c.iter_max = nm->iter_max; c.alpha = nm->alpha; c.beta = nm->beta; c.gamma = nm->gamma; c.eps = nm->eps;
  c.lambda = nm->lambda; c.kl = kl_bound;
  const size_t Pb = (size_t)P * 8;
  std::vector<int32_t> ones(P, 1), izeros(P, 0);
  ratilqr_batch_in in;
  in.P = P; in.K = 6; in.x0 = x0; in.x0_count = x0_count; in.u_init = u_init; in.u_count = u_count; in.theta = nullptr;
  int rc = 0;
  if (kl_bound > 0) { if ((rc = stage_internal(ctx, desc, opts, &in, 0, true))) return rc; }
  UP(ctx->s[0], theta_high_init, Pb); UP(ctx->s[1], theta_low_init, Pb);   // initialize! :164-168: theta_* <- *_init
  UP(ctx->s[2], theta_high_init, Pb); UP(ctx->s[3], theta_low_init, Pb);
  UP(ctx->s[4], c_high, Pb); UP(ctx->s[5], c_low, Pb);
  CU(ctx->s[6].reserve(Pb)); CU(ctx->s[7].reserve(Pb));
  UP(ctx->s[8], has_c, (size_t)P * 8); UP(ctx->s[9], izeros.data(), (size_t)P * 4);
  UP(ctx->s[10], ones.data(), (size_t)P * 4); UP(ctx->s[11], izeros.data(), (size_t)P * 4);
  UP(ctx->s[12], izeros.data(), (size_t)P * 4); CU(ctx->s[13].reserve(8));
  c.th_high = ctx->s[0].as<double>(); c.th_low = ctx->s[1].as<double>(); c.th_high_init = ctx->s[2].as<double>();
  c.th_low_init = ctx->s[3].as<double>(); c.c_high = ctx->s[4].as<double>(); c.c_low = ctx->s[5].as<double>();
  c.theta_opt = ctx->s[6].as<double>(); c.value_out = ctx->s[7].as<double>(); c.has_c = ctx->s[8].as<int32_t>();
  c.iter = ctx->s[9].as<int32_t>(); c.active = ctx->s[10].as<int32_t>(); c.evals = ctx->s[11].as<int32_t>();
  c.phase = ctx->s[12].as<int32_t>(); c.n_active = ctx->s[13].as<int32_t>();
  int rounds = 0;
  // same work placement as the CE fleet: problems heaviest-first by the iterations they needed in the previous round /
  // previous call (resort_slots_by_last_iters); thread-per-instance kernel only
  const char* esf = getenv("RATILQR_FLEET_SORT");
  const bool sort_fleet = !(esf && esf[0] == '0') && P >= 64;
  if (kl_bound > 0) {
    c.theta = ctx->d_theta.as<double>(); c.value = ctx->sp.value; c.status = ctx->sp.status;
    ctx->sp.active = c.active;
    if (sort_fleet && !ctx->coop && (int)ctx->fleet_key.size() == P) { if ((rc = apply_slot_order(ctx, ctx->fleet_key, P, 6))) return rc; }
    while (true) {
      rll::launch_nm_candidates(c, st);
      if ((rc = check_launch(ctx, "k_nm_candidates"))) return rc;
      if ((rc = run_internal(ctx, 1, nullptr))) return rc;
      CU(cudaMemsetAsync(c.n_active, 0, 4, st));
      rll::launch_nm_decide(c, st);
      if ((rc = check_launch(ctx, "k_nm_decide"))) return rc;
      int32_t n_active = 0;
      CU(cudaMemcpyAsync(&n_active, c.n_active, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      if (n_active == 0) break;
      if (++rounds > 100000) FAIL(-6, "Nelder-Mead vertex search does not terminate (the reference would spin forever here)");
      // the first rounds establish the profile; afterwards the per-problem work is stable, re-sort every 4th round
      if (sort_fleet && !ctx->coop && (rounds <= 2 || rounds % 4 == 0)) { if ((rc = resort_slots_by_last_iters(ctx, P, 6))) return rc; }
    }
  }
  in.K = 1;
  const int fwant = final_out ? ((final_out->x ? 1 : 0) | (final_out->l ? 2 : 0) | (final_out->L ? 4 : 0)) : 0;
  if ((rc = stage_internal(ctx, desc, opts, &in, 0, true, fwant))) return rc;
  if (sort_fleet && !ctx->coop && (int)ctx->fleet_key.size() == P) { if ((rc = apply_slot_order(ctx, ctx->fleet_key, P, 1))) return rc; }
  rll::launch_nm_final(c, ctx->d_theta.as<double>(), nullptr, nullptr, 0, st);
  if ((rc = check_launch(ctx, "k_nm_final"))) return rc;
  if ((rc = run_internal(ctx, 1, nullptr))) return rc;
  rll::launch_nm_final(c, ctx->d_theta.as<double>(), ctx->sp.value, ctx->sp.status, 1, st);
  if ((rc = check_launch(ctx, "k_nm_final"))) return rc;
  DOWNSYNC(theta_opt, c.theta_opt, Pb); DOWNSYNC(value, c.value_out, Pb);
  DOWNSYNC(theta_high_init, c.th_high_init, Pb); DOWNSYNC(theta_low_init, c.th_low_init, Pb);
  DOWNSYNC(c_high, c.c_high, Pb); DOWNSYNC(c_low, c.c_low, Pb); DOWNSYNC(has_c, c.has_c, (size_t)P * 8);
  DOWNSYNC(nm_iters, c.iter, (size_t)P * 4); DOWNSYNC(n_evals, c.evals, (size_t)P * 4);
  CU(cudaStreamSynchronize(st));
  if (final_out) return fetch_internal(ctx, final_out);
  return 0;
}

int32_t ratilqr_nm_solve(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                         const ratilqr_nm_opts* nm, const double* x0, const double* u_init, double kl_bound,
                         double* theta_high_init, double* theta_low_init, double* c_high, double* c_low, int32_t* has_c,
                         double* theta_opt, double* value, int32_t* nm_iters, int32_t* n_evals, ratilqr_ileqg_out* final_out) {
  return ratilqr_nm_solve_fleet(ctx, desc, opts, nm, 1, x0, 1, u_init, 1, kl_bound, theta_high_init, theta_low_init, c_high,
                                c_low, has_c, theta_opt, value, nm_iters, n_evals, final_out);
}

// ======== multi-GPU inside the C ABI: NCCL communicators owned by the contexts =========================
int32_t ratilqr_nccl_unique_id(uint8_t* id128) {
  if (!id128) return -1;
  if (nccl::load()) return -20;
  nccl::UniqueId id;
  if (nccl::GetUniqueId(&id) != 0) return -21;
  memcpy(id128, id.internal, 128);
  return 0;
}

int32_t ratilqr_attach_comm(ratilqr_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t world) {
  if (!ctx) return -1;
  if (!id128 || world < 1 || rank < 0 || rank >= world) FAIL(-1, "bad communicator description");
  if (const char* e = nccl::load()) FAIL(-20, e);
  if (ctx->comm) FAIL(-1, "a communicator is already attached to this ctx");
  CU(cudaSetDevice(ctx->device));
  nccl::UniqueId id;
  memcpy(id.internal, id128, 128);
  NC(nccl::CommInitRank(&ctx->comm, world, id, rank));
  ctx->rank = rank; ctx->world = world;
  return 0;
}

int32_t ratilqr_comm_info(const ratilqr_ctx* ctx, int32_t* rank, int32_t* world) {
  if (!ctx) return -1;
  if (rank) *rank = ctx->rank;
  if (world) *world = ctx->comm ? ctx->world : 1;
  return 0;
}

// compute_cost (cross_entropy...jl:173-195) of ONE problem's K-sample theta population, sharded: this rank solves its
// block, one ncclAllGather of (value, status) on device buffers, cost = value + kl/theta for the whole population
int32_t ratilqr_ce_costs_sharded(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                                 const double* x0, const double* u_init, const double* theta, int32_t K, double kl_bound,
                                 double* cost, int32_t* status) {
  if (!ctx) return -1;
  if (!theta || !cost || K < 1) FAIL(-1, "bad arguments");
  if (!ctx->comm || ctx->world < 2) FAIL(-1, "no communicator attached (ratilqr_attach_comm / ratilqr_create_multi)");
  if (K < ctx->world) FAIL(-1, "population smaller than the world size");
  const int W = ctx->world, lo = (int)((long long)K * ctx->rank / W), cnt = (int)((long long)K * (ctx->rank + 1) / W) - lo;
  ratilqr_batch_in in;
  in.P = 1; in.K = cnt; in.x0 = x0; in.x0_count = 1; in.u_init = u_init; in.u_count = 1; in.theta = theta + lo;
  int rc = stage_internal(ctx, desc, opts, &in, 0);
  if (rc) return rc;
  if ((rc = run_internal(ctx, 1, nullptr))) return rc;
  CU(ctx->d_sh[0].reserve((size_t)K * 8)); CU(ctx->d_sh[1].reserve((size_t)K * 8)); CU(ctx->d_sh[2].reserve((size_t)K * 4));
  UP(ctx->d_sh[0], theta, (size_t)K * 8);
  if ((rc = shard_allgather(ctx, K, lo, cnt, ctx->sp.value, ctx->sp.status, ctx->d_sh[1].as<double>(), ctx->d_sh[2].as<int32_t>()))) return rc;
  CU(ctx->d_cost.reserve((size_t)K * 8));
  k_ce_cost<<<(K + 127) / 128, 128, 0, ctx->stream>>>(K, ctx->d_sh[1].as<double>(), ctx->d_sh[2].as<int32_t>(), ctx->d_sh[0].as<double>(), kl_bound, ctx->d_cost.as<double>());
  if ((rc = check_launch(ctx, "k_ce_cost"))) return rc;
  CU(cudaMemcpyAsync(cost, ctx->d_cost.p, (size_t)K * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CU(cudaMemcpyAsync(status, ctx->d_sh[2].p, (size_t)K * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// solve!(::CrossEntropyBilevelOptimizationSolver) of ONE problem with the theta population sharded over the ranks
int32_t ratilqr_ce_solve_sharded(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                                 const ratilqr_ce_opts* ce, const double* x0, const double* u_init, double kl_bound,
                                 const double* z_inject, int64_t nz, uint64_t seed, double* mu_init, double* sigma_init,
                                 double* theta_opt, double* value, double* theta_min, double* theta_max, double* mu,
                                 double* sigma, int64_t* nz_used, int32_t* rounds_out, ratilqr_ileqg_out* final_out) {
  if (!ctx) return -1;
  return ce_solve_fleet_block(ctx, desc, opts, ce, 1, 0, x0, 1, u_init, 1, kl_bound, z_inject, nz, seed, mu_init, sigma_init,
                              theta_opt, value, theta_min, theta_max, mu, sigma, nz_used, rounds_out, final_out, nullptr, true);
}

// compute_cost of PETS (pets.jl:100-126) sharded over the action sequences: this rank rolls out its block (Philox streams
// are indexed by the GLOBAL sequence / particle number, so the result does not depend on the world size), one all-gather
int32_t ratilqr_pets_costs_sharded(ratilqr_ctx* ctx, const ratilqr_problem_desc* desc, const ratilqr_generative_desc* gen,
                                   const double* x0, const double* controls, int32_t C, int32_t particles,
                                   const double* noise, uint64_t seed, double* cost) {
  if (!ctx) return -1;
  if (!controls || !cost) FAIL(-1, "null argument");
  if (!ctx->comm || ctx->world < 2) FAIL(-1, "no communicator attached (ratilqr_attach_comm / ratilqr_create_multi)");
  if (C < ctx->world) FAIL(-1, "fewer action sequences than ranks");
  const int W = ctx->world, lo = (int)((long long)C * ctx->rank / W), cnt = (int)((long long)C * (ctx->rank + 1) / W) - lo;
  rll::PetsArgs a;
  const rlu::Module* um = nullptr;
  if (int rc = pets_fill(ctx, desc, gen, x0, cnt, particles, a, &um)) return rc;
  const size_t n = desc->n, m = desc->m, N = desc->N;
  UP(ctx->s[0], controls + m * N * lo, m * N * cnt * 8);
  a.controls = ctx->s[0].as<double>();
  if (noise) { UP(ctx->s[1], noise + n * N * (size_t)particles * lo, n * N * (size_t)particles * cnt * 8); a.noise = ctx->s[1].as<double>(); }
  a.seed = seed; a.stream_offset = (uint64_t)lo * (uint64_t)particles;
  CU(ctx->s[2].reserve((size_t)cnt * 8)); CU(ctx->d_sh[1].reserve((size_t)C * 8));
  a.cost = ctx->s[2].as<double>();
  if (int rc = pets_launch_costs(ctx, um, a)) return rc;
  if (int rc = check_launch(ctx, "k_pets_costs", um ? 0 : 1)) return rc;
  CU(ctx->d_sh[5].reserve((size_t)cnt * 4));
  CU(cudaMemsetAsync(ctx->d_sh[5].p, 0, (size_t)cnt * 4, ctx->stream));
  if (int rc = shard_allgather(ctx, C, lo, cnt, a.cost, ctx->d_sh[5].as<int32_t>(), ctx->d_sh[1].as<double>(), nullptr)) return rc;
  DOWNSYNC(cost, ctx->d_sh[1].p, (size_t)C * 8);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- one process driving several GPUs: contexts + communicators from ncclCommInitAll, one host thread per device -----
struct ratilqr_multi {
  std::vector<ratilqr_ctx*> ctxs;
  std::string err;
};

int32_t ratilqr_create_multi(ratilqr_multi** out, const int32_t* device_ids, int32_t n_dev) {
  if (!out || !device_ids || n_dev < 1) return -1;
  *out = nullptr;
  if (n_dev > 1 && nccl::load()) return -20;
  ratilqr_multi* mg = new ratilqr_multi();
  for (int i = 0; i < n_dev; ++i) {
    ratilqr_ctx* c = nullptr;
    int rc = ratilqr_create(&c, device_ids[i]);
    if (rc) { for (ratilqr_ctx* q : mg->ctxs) ratilqr_destroy(q); delete mg; return rc; }
    mg->ctxs.push_back(c);
  }
  if (n_dev > 1) {
    std::vector<void*> comms(n_dev, nullptr);
    std::vector<int> devs(device_ids, device_ids + n_dev);
    if (nccl::CommInitAll(comms.data(), n_dev, devs.data()) != 0) { for (ratilqr_ctx* q : mg->ctxs) ratilqr_destroy(q); delete mg; return -21; }
    for (int i = 0; i < n_dev; ++i) { mg->ctxs[i]->comm = comms[i]; mg->ctxs[i]->rank = i; mg->ctxs[i]->world = n_dev; }
  }
  *out = mg;
  return 0;
}
int32_t ratilqr_destroy_multi(ratilqr_multi* mg) {
  if (!mg) return 0;
  for (ratilqr_ctx* c : mg->ctxs) ratilqr_destroy(c);
  delete mg;
  return 0;
}
int32_t ratilqr_multi_size(const ratilqr_multi* mg) { return mg ? (int32_t)mg->ctxs.size() : 0; }
ratilqr_ctx* ratilqr_multi_ctx(ratilqr_multi* mg, int32_t i) { return (mg && i >= 0 && i < (int)mg->ctxs.size()) ? mg->ctxs[i] : nullptr; }
const char* ratilqr_multi_last_error(const ratilqr_multi* mg) { return mg ? mg->err.c_str() : "null handle"; }

}  // extern "C"

// runs fn(rank) on one host thread per device and returns the first failure; every rank's results land in rank 0's outputs
template <class F>
static int multi_run(ratilqr_multi* mg, F&& fn) {
  const int W = (int)mg->ctxs.size();
  std::vector<int> rcs(W, 0);
  std::vector<std::thread> th;
  for (int r = 1; r < W; ++r) th.emplace_back([&, r] { rcs[r] = fn(r); });
  rcs[0] = fn(0);
  for (auto& t : th) t.join();
  for (int r = 0; r < W; ++r) if (rcs[r]) { mg->err = mg->ctxs[r]->err; return rcs[r]; }
  return 0;
}

extern "C" {

int32_t ratilqr_multi_ce_costs(ratilqr_multi* mg, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                               const double* x0, const double* u_init, const double* theta, int32_t K, double kl_bound,
                               double* cost, int32_t* status) {
  if (!mg || !cost || K < 1) return -1;
  const int W = (int)mg->ctxs.size();
  if (W == 1) {
    ratilqr_batch_in in;
    in.P = 1; in.K = K; in.x0 = x0; in.x0_count = 1; in.u_init = u_init; in.u_count = 1; in.theta = theta;
    int rc = ratilqr_ce_costs(mg->ctxs[0], desc, opts, &in, kl_bound, cost, status);
    if (rc) mg->err = mg->ctxs[0]->err;
    return rc;
  }
  std::vector<std::vector<double>> c(W, std::vector<double>(K));
  std::vector<std::vector<int32_t>> s(W, std::vector<int32_t>(K));
  int rc = multi_run(mg, [&](int r) { return ratilqr_ce_costs_sharded(mg->ctxs[r], desc, opts, x0, u_init, theta, K, kl_bound, c[r].data(), s[r].data()); });
  if (rc) return rc;
  for (int r = 1; r < W; ++r)  // the gathered vectors are replicated: a cheap end-to-end check of the collective
    if (memcmp(c[r].data(), c[0].data(), (size_t)K * 8) != 0) { mg->err = "ranks disagree on the gathered cost vector"; return -22; }
  memcpy(cost, c[0].data(), (size_t)K * 8);
  if (status) memcpy(status, s[0].data(), (size_t)K * 4);
  return 0;
}

int32_t ratilqr_multi_ce_solve(ratilqr_multi* mg, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                               const ratilqr_ce_opts* ce, const double* x0, const double* u_init, double kl_bound,
                               const double* z_inject, int64_t nz, uint64_t seed, double* mu_init, double* sigma_init,
                               double* theta_opt, double* value, double* theta_min, double* theta_max, double* mu,
                               double* sigma, int64_t* nz_used, int32_t* rounds_out, ratilqr_ileqg_out* final_out) {
  if (!mg || !mu_init || !sigma_init || !theta_opt || !value) return -1;
  const int W = (int)mg->ctxs.size();
  if (W == 1) {
    int rc = ratilqr_ce_solve(mg->ctxs[0], desc, opts, ce, x0, u_init, kl_bound, z_inject, nz, seed, mu_init, sigma_init, theta_opt,
                              value, theta_min, theta_max, mu, sigma, nz_used, rounds_out, final_out);
    if (rc) mg->err = mg->ctxs[0]->err;
    return rc;
  }
  // every rank carries the whole (replicated) CE state; rank 0 writes the caller's outputs, the others scratch copies
  struct Scratch { double mi, si, to, va, tmin, tmax, mu, sg; int64_t nzu; int32_t rounds; };
  std::vector<Scratch> sc(W);
  for (int r = 0; r < W; ++r) { sc[r].mi = *mu_init; sc[r].si = *sigma_init; }
  int rc = multi_run(mg, [&](int r) {
    Scratch& q = sc[r];
    return ratilqr_ce_solve_sharded(mg->ctxs[r], desc, opts, ce, x0, u_init, kl_bound, z_inject, nz, seed, &q.mi, &q.si, &q.to, &q.va,
                                    &q.tmin, &q.tmax, &q.mu, &q.sg, &q.nzu, &q.rounds, r == 0 ? final_out : nullptr);
  });
  if (rc) return rc;
  for (int r = 1; r < W; ++r)
    if (sc[r].to != sc[0].to || sc[r].mu != sc[0].mu || sc[r].nzu != sc[0].nzu) { mg->err = "ranks disagree on the replicated CE state"; return -22; }
  *mu_init = sc[0].mi; *sigma_init = sc[0].si; *theta_opt = sc[0].to; *value = sc[0].va;
  if (theta_min) *theta_min = sc[0].tmin;
  if (theta_max) *theta_max = sc[0].tmax;
  if (mu) *mu = sc[0].mu;
  if (sigma) *sigma = sc[0].sg;
  if (nz_used) *nz_used = sc[0].nzu;
  if (rounds_out) *rounds_out = sc[0].rounds;
  return 0;
}

int32_t ratilqr_multi_pets_costs(ratilqr_multi* mg, const ratilqr_problem_desc* desc, const ratilqr_generative_desc* gen,
                                 const double* x0, const double* controls, int32_t C, int32_t particles, const double* noise,
                                 uint64_t seed, double* cost) {
  if (!mg || !cost || C < 1) return -1;
  const int W = (int)mg->ctxs.size();
  if (W == 1) {
    int rc = ratilqr_pets_costs(mg->ctxs[0], desc, gen, x0, controls, C, particles, noise, seed, cost);
    if (rc) mg->err = mg->ctxs[0]->err;
    return rc;
  }
  std::vector<std::vector<double>> c(W, std::vector<double>(C));
  int rc = multi_run(mg, [&](int r) { return ratilqr_pets_costs_sharded(mg->ctxs[r], desc, gen, x0, controls, C, particles, noise, seed, c[r].data()); });
  if (rc) return rc;
  memcpy(cost, c[0].data(), (size_t)C * 8);
  return 0;
}

// fleet of independent RAT iLQR problems block-partitioned over the devices: no collective; results land
// in the per-device slices of the caller's arrays; Philox streams are indexed by the global problem number
int32_t ratilqr_multi_ce_solve_fleet(ratilqr_multi* mg, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                                     const ratilqr_ce_opts* ce, int32_t P, const double* x0, const double* u_init,
                                     int32_t u_count, double kl_bound, uint64_t seed, double* mu_init, double* sigma_init,
                                     double* theta_opt, double* value, double* l_out) {
  if (!mg || !desc || !ce || P < 1 || !x0 || !u_init || !mu_init || !sigma_init || !theta_opt || !value) return -1;
  const int W = std::min((int)mg->ctxs.size(), (int)P);
  if (desc->cost_params_count != 1 && desc->cost_params_count != P) { mg->err = "cost_params_count must be 1 or P"; return -1; }
  const size_t n = desc->n, m = desc->m, N = desc->N;
  std::vector<int> rcs(W, 0);
  std::vector<std::thread> th;
  auto run = [&](int r) {
    const long long lo = (long long)P * r / W, hi = (long long)P * (r + 1) / W;
    const int Pb = (int)(hi - lo);
    ratilqr_problem_desc d = *desc;
    if (desc->cost_params_count == P) { d.cost_params = desc->cost_params + (size_t)lo * desc->n_cost_params; d.cost_params_count = Pb; }
    ratilqr_ileqg_out fo;
    memset(&fo, 0, sizeof(fo));
    if (l_out) fo.l = l_out + m * N * lo;
    // Philox streams are indexed by the global problem number (p0 = lo): problem p draws what it would draw on one device
    rcs[r] = ce_solve_fleet_block(mg->ctxs[r], &d, opts, ce, Pb, lo, x0 + n * lo, Pb, u_count == P ? u_init + m * N * lo : u_init,
                                  u_count == P ? Pb : 1, kl_bound, nullptr, 0, seed, mu_init + lo, sigma_init + lo, theta_opt + lo,
                                  value + lo, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, l_out ? &fo : nullptr);
  };
  for (int r = 1; r < W; ++r) th.emplace_back(run, r);
  run(0);
  for (auto& t : th) t.join();
  for (int r = 0; r < W; ++r) if (rcs[r]) { mg->err = mg->ctxs[r]->err; return rcs[r]; }
  return 0;
}

int32_t ratilqr_fp64_peak_probe(ratilqr_ctx* ctx, double* tflops, float* ms) {
  if (!ctx) return -1;
  CU(cudaSetDevice(ctx->device));
  CU(ctx->s[0].reserve(64));
  const int iters = 1 << 16;
  rll::launch_fp64_probe(ctx->s[0].as<double>(), 1 << 10, ctx->stream);  // warm-up
  float best = 1e30f;
  double flops = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    flops = rll::launch_fp64_probe(ctx->s[0].as<double>(), iters, ctx->stream);
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev1));
    float t = 0;
    CU(cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
    if (t < best) best = t;
  }
  if (int rc = check_launch(ctx, "k_fp64_probe", 4)) return rc;
  if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
  if (ms) *ms = best;
  return 0;
}

int32_t ratilqr_fp64_peak_probe_sustained(ratilqr_ctx* ctx, double seconds, double* tflops) {
  if (!ctx) return -1;
  CU(cudaSetDevice(ctx->device));
  CU(ctx->s[0].reserve(64));
  const int iters = 1 << 16;
  double flops = rll::launch_fp64_probe(ctx->s[0].as<double>(), iters, ctx->stream);
  CU(cudaStreamSynchronize(ctx->stream));
  // time the second half of the run only (clocks have settled under the power cap by then)
  const int n_launch = (int)(seconds / 0.0092) + 2, half = n_launch / 2;
  for (int i = 0; i < half; ++i) rll::launch_fp64_probe(ctx->s[0].as<double>(), iters, ctx->stream);
  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int i = half; i < n_launch; ++i) rll::launch_fp64_probe(ctx->s[0].as<double>(), iters, ctx->stream);
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  CU(cudaEventSynchronize(ctx->ev1));
  float t = 0;
  CU(cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
  if (int rc = check_launch(ctx, "k_fp64_probe", n_launch + 1)) return rc;
  if (tflops) *tflops = flops * (n_launch - half) / (t * 1e-3) / 1e12;
  return 0;
}

}  // extern "C"
