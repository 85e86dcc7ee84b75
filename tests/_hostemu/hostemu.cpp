// hostemu.cpp -- TEST INFRASTRUCTURE: compiles the per-instance device arithmetic of
// ratilqr.jl_b200/csrc/rl_core.cuh + rl_components.cuh with g++ and runs it on host memory, one
// loop iteration per "thread".  It exists so that the kernel logic can be checked against the
// oracle on a CPU-only box (`pytest -m "not gpu"`).  It is never linked into, loaded by, or
// shipped with libratilqr_b200.so.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../ratilqr.jl_b200/csrc/rl_components.cuh"
#include "../../ratilqr.jl_b200/csrc/rl_coop.cuh"
#include "../../ratilqr.jl_b200/csrc/rl_spec.cuh"
#include "../../ratilqr.jl_b200/csrc/rl_coop2.cuh"
#include "../../ratilqr.jl_b200/csrc/rl_host.hpp"

#include "../../ratilqr.jl_b200/csrc/rl_user.cuh"

using namespace rl;

// ---- statically compiled copies of the USER-MODEL snippets of tests/user_models/ ---------------------------------
// The GPU tests hand the same text to NVRTC (ratilqr_user_model_register); here g++ compiles it through the same
// adapters (rl_user.cuh: UserDyn / UserCost, forward-mode duals), so the user-model arithmetic has a CPU check too.
// hostemu-only numbering: model 1000 = unicycle snippet, 1001 = drag car, 1002 = unicycle with declared structure;
// cost 100 = goal cost, 101 = obstacle cost, 102 = goal cost with declared (diagonal) structure.
using rl::square;
namespace um_unicycle {
#include "../user_models/unicycle_dynamics.inc"
}
namespace um_dragcar {
#include "../user_models/drag_car_dynamics.inc"
}
namespace um_goal {
#include "../user_models/goal_cost.inc"
}
namespace um_obstacle {
#include "../user_models/obstacle_cost.inc"
}
struct UmUnicycleBody { template <class T> void operator()(const double* p, const T* x, const T* u, T* xn) const { um_unicycle::dynamics<T>(p, x, u, xn); } };
struct UmDragCarBody { template <class T> void operator()(const double* p, const T* x, const T* u, T* xn) const { um_dragcar::dynamics<T>(p, x, u, xn); } };
struct UmGoalFn {
  template <class T> T stage(const double* cp, int k, const T* x, const T* u) const { return um_goal::stage_cost<T>(cp, k, x, u); }
  template <class T> T terminal(const double* cp, const T* x) const { return um_goal::terminal_cost<T>(cp, x); }
};
struct UmObstacleFn {
  template <class T> T stage(const double* cp, int k, const T* x, const T* u) const { return um_obstacle::stage_cost<T>(cp, k, x, u); }
  template <class T> T terminal(const double* cp, const T* x) const { return um_obstacle::terminal_cost<T>(cp, x); }
};
// declared structure (what ratilqr_user_model_desc.a_kind ... p_kind generate): model 1002 / cost 102
struct UmUnicycleKinds {
  static constexpr bool structured = true;
  static constexpr int a_kind(int i, int j) { return i == j ? 1 : ((i < 2 && j >= 2) ? 2 : 0); }
  static constexpr int b_kind(int i, int j) { return ((i == 2 && j == 1) || (i == 3 && j == 0)) ? 2 : 0; }
};
struct UmDiagCostKinds {
  static constexpr int q_kind(int i, int j) { return i == j ? 2 : 0; }
  static constexpr int r_kind(int i, int j) { return i == j ? 2 : 0; }
  static constexpr int p_kind(int, int) { return 0; }
};
typedef UserDyn<4, 2, UmUnicycleBody, UmUnicycleKinds> UmUnicycleS;
typedef UserCost<4, 2, 11, UmGoalFn, UmDiagCostKinds> UmGoalS;
typedef UserDyn<4, 2, UmUnicycleBody> UmUnicycle;
typedef UserDyn<4, 2, UmDragCarBody> UmDragCar;
typedef UserCost<4, 2, 11, UmGoalFn> UmGoal;
typedef UserCost<4, 2, 15, UmObstacleFn> UmObstacle;

static const char* hm_check_desc(const ratilqr_problem_desc* d, bool differentiable) {
  if (d && d->model_id < 1000 && d->cost_id < 100) return rlh::check_desc(d, differentiable);  // registered pair
  if (!d || d->n != 4 || d->m != 2 || d->N < 1 || !d->W || !d->cost_params) return "bad user-model description";
  return nullptr;
}

template <class F> static int dispatch(int model_id, int cost_id, F&& fn) {
  if (model_id >= 1000 || cost_id == 100 || cost_id == 101 || cost_id == 102) {  // (RL_COST_QUAD_DIAG is 0x101: not a user id)
    if (model_id == 1002 && cost_id == 102) { fn(UmUnicycleS(), UmGoalS()); return 0; }
    if (model_id == 1002 && cost_id == 100) { fn(UmUnicycleS(), UmGoal()); return 0; }
    if (model_id == 1000 && cost_id == 102) { fn(UmUnicycle(), UmGoalS()); return 0; }
    using Quad = Cost<RATILQR_COST_QUADRATIC, 4, 2>;
    if (model_id == 1000 && cost_id == RATILQR_COST_QUADRATIC) { fn(UmUnicycle(), Quad()); return 0; }
    if (model_id == 1000 && cost_id == 100) { fn(UmUnicycle(), UmGoal()); return 0; }
    if (model_id == RATILQR_MODEL_UNICYCLE && cost_id == 100) { fn(Dyn<RATILQR_MODEL_UNICYCLE>(), UmGoal()); return 0; }
    if (model_id == 1000 && cost_id == 101) { fn(UmUnicycle(), UmObstacle()); return 0; }
    if (model_id == 1001 && cost_id == 101) { fn(UmDragCar(), UmObstacle()); return 0; }
    if (model_id == 1001 && cost_id == RATILQR_COST_QUADRATIC) { fn(UmDragCar(), Quad()); return 0; }
    return -5;
  }
#define X(MID, CID) if (model_id == MID && cost_id == CID) { fn(Dyn<MID>(), Cost<CID, Dyn<MID>::n, Dyn<MID>::m>()); return 0; }
  RL_FOR_EACH_ILEQG_COMBO(X)
  RL_FOR_EACH_ROLLOUT_ONLY_COMBO(X)
  RL_FOR_EACH_DIAG_COMBO(X)
#undef X
  return -5;
}

template <int n, int m>
static void ric(int N, int B, int optimise, const double* q, const double* qv, const double* Q, const double* r,
                const double* R, const double* Pm, const double* A, const double* Bm, const rlh::WPrep& wp, int W_tv,
                const double* theta, double mu_min, double delta_0, double* mu, double* delta, double* L, double* dl,
                double* s, double* sv, double* S, int32_t* status, int32_t* restarts) {
  for (int b = 0; b < B; ++b) {
    int32_t nr = 0;
    int st = comp_riccati<n, m>(N, optimise, q + (size_t)b * (N + 1), qv + (size_t)b * n * (N + 1), Q + (size_t)b * n * n * (N + 1),
                                r + (size_t)b * m * N, R + (size_t)b * m * m * N, Pm + (size_t)b * m * n * N,
                                A + (size_t)b * n * n * N, Bm + (size_t)b * n * m * N, wp.W.data(), wp.Winv.data(), wp.detW.data(), W_tv,
                                theta[b], mu_min, delta_0, &mu[b], &delta[b], L + (size_t)b * m * n * N,
                                dl ? dl + (size_t)b * m * N : nullptr, s + (size_t)b * (N + 1), sv + (size_t)b * n * (N + 1),
                                S + (size_t)b * n * n * (N + 1), &nr);
    if (status) status[b] = st;
    if (restarts) restarts[b] = nr;
  }
}

static int g_dynamic = 0;  // 1: emulate the opt-in persistent kernel with lane-level refill (RATILQR_DYNAMIC=1)
static int g_coop = 0;  // 1: emulate the warp-cooperative kernel (32 virtual lanes run phase by phase)
static int g_spec = 0;  // 2 / 4 / 8: emulate the speculative latency kernel with this many lanes per instance (rl_spec.cuh)

// the speculative kernel on the host: the G lanes of a group run their round one after another, then the decision is
// replayed once (on the device every lane replays it on shuffled copies of the same results)
template <class D, class CT, int G>
static void run_spec(const SolveParams& P, size_t B) {
  double stage_area[2 * RL_STAGE_NV];
  Stage sg; sg.base = stage_area; sg.stride = 1;
  for (size_t slot = 0; slot < B; ++slot) {
    const size_t inst = P.perm ? (size_t)P.perm[slot] : slot;
    const size_t p = inst / (size_t)P.K;
    if (P.active && !P.active[p]) continue;
    const double* cp = P.cost_params + (P.cp_count > 1 ? p * (size_t)P.ncp : 0);
    const double theta = P.theta[inst];
    const size_t t0 = slot * G;  // first thread of the group
    SpecCols C;
    C.P = &P; C.tile = t0 >> 5; C.lane0 = t0 & 31; C.n = D::n; C.m = D::m; C.N = P.N;
    SpecState S;
    spec_state_init(P, S);
    while (!S.done) {
      SpecLaneRes res[G];
      for (int g = 0; g < G; ++g) res[g] = spec_lane_work<D, CT>(P, C, g, S, cp, theta, p, sg);
      spec_decide<G>(P, S, res, inst, true);
    }
    for (int g = 0; g < G; ++g) spec_write_outputs(P, C, S, inst, g, G);
  }
}

extern "C" {

int32_t hostemu_set_coop(int32_t v) { g_coop = v; return 0; }
int32_t hostemu_set_spec(int32_t v) { g_spec = (v == 2 || v == 4 || v == 8) ? v : 0; return 0; }
int32_t hostemu_set_dynamic(int32_t v) { g_dynamic = v; return 0; }

int32_t hostemu_ileqg_solve_batch(void*, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                                  const ratilqr_batch_in* in, ratilqr_ileqg_out* out) {
  if (hm_check_desc(desc, true)) return -1;
  const int n = desc->n, m = desc->m, N = desc->N;
  const size_t B = (size_t)in->P * in->K;
  rlh::WPrep wp;
  if (!rlh::prep_W(n, N, desc->W, desc->W_time_varying, wp)) return -2;
  const bool spec = g_spec && !g_coop && !g_dynamic && n <= 6 && desc->model_id < 1000 && desc->cost_id < 100;
  const size_t cols = spec ? B * g_spec : B, pol = spec ? 2 : 1;
  const size_t Bp = (cols + 31) / 32 * 32;  // warp-tiled workspace
  const WsLayout wl = ws_layout(n, m, N, (int)pol, model_naux(desc->model_id));  // per-tile records, as the device library lays them out
  std::vector<double> ws(wl.rec * Bp, 0.0), value(B), mu(B), dcur(B), eps;
  double *X = ws.data(), *U = X + wl.oU * 32, *Lg = X + wl.oLg * 32;
  std::vector<int32_t> status(B), iters(B), trials(B), restarts(B), cur(B);
  int cap = out->eps_hist ? out->eps_hist_cap : 0;
  eps.assign(B * cap * 2 + 2, 0.0);
  SolveParams P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.B = (int)B; P.K = in->K;
  for (int i = 0; i < 8; ++i) P.mp[i] = i < desc->n_model_params ? desc->model_params[i] : 0.0;
  P.cost_params = desc->cost_params; P.ncp = desc->n_cost_params; P.cp_count = desc->cost_params_count;
  P.W = wp.W.data(); P.Winv = wp.Winv.data(); P.detW = wp.detW.data(); P.W_tv = desc->W_time_varying;
  P.x0 = in->x0; P.x0_count = in->x0_count; P.u_init = in->u_init; P.u_count = in->u_count; P.theta = in->theta;
  P.mu_min = opts->mu_min; P.delta_0 = opts->delta_0; P.lambda = opts->lambda; P.d = opts->d;
  P.iter_max = opts->iter_max; P.eps_auto = opts->adaptive_eps_init; P.eps_init = opts->eps_init; P.eps_min = opts->eps_min;
  P.X = X; P.U = U; P.Lg = Lg; P.DL = X + wl.oDL * 32; P.AUX = X + wl.oAux * 32; P.rec = wl.rec;
  P.value = value.data(); P.status = status.data(); P.iters = iters.data(); P.trials = trials.data();
  P.restarts = restarts.data(); P.mu_out = mu.data(); P.d_out = dcur.data(); P.cur = cur.data();
  P.eps_hist = cap ? eps.data() : nullptr; P.eps_hist_cap = cap;
  // same slot->instance permutation as the device library (theta ascending within each problem)
  std::vector<int32_t> perm(B);
  for (int p = 0; p < in->P; ++p) {
    int32_t* pp = perm.data() + (size_t)p * in->K;
    for (int i = 0; i < in->K; ++i) pp[i] = (int32_t)((size_t)p * in->K + i);
    std::stable_sort(pp, pp + in->K, [&](int32_t a, int32_t b2) { return in->theta[a] < in->theta[b2]; });
  }
  P.perm = in->K >= 2 ? perm.data() : nullptr;
  const int cost_id = (desc->model_id < 1000 && rlh::quad_is_diag(desc) && desc->model_id != RATILQR_MODEL_QUADROTOR && desc->model_id != RATILQR_MODEL_POWER_LAW) ? RL_COST_QUAD_DIAG : desc->cost_id;
  if (spec) {
    std::vector<double> xo((size_t)n * (N + 1) * B), lo((size_t)m * N * B), Lo((size_t)m * n * N * B, 0.0);
    P.xo = xo.data(); P.lo = lo.data(); P.Lo = Lo.data();
    const int cid = (rlh::quad_is_diag(desc) && desc->model_id != RATILQR_MODEL_POWER_LAW) ? RL_COST_QUAD_DIAG : desc->cost_id;
    int rc2 = dispatch(desc->model_id, cid, [&](auto D, auto CT) {
      using DD = decltype(D);
      using CC = decltype(CT);
      if constexpr (DD::n <= 6) {
        if (g_spec == 8) run_spec<DD, CC, 8>(P, B);
        else if (g_spec == 4) run_spec<DD, CC, 4>(P, B);
        else run_spec<DD, CC, 2>(P, B);
      }
    });
    if (rc2) return rc2;
    for (size_t b = 0; b < B; ++b) {
      if (out->value) out->value[b] = value[b];
      if (out->status) out->status[b] = status[b];
      if (out->iters) out->iters[b] = iters[b];
      if (out->trials) out->trials[b] = trials[b];
      if (out->restarts) out->restarts[b] = restarts[b];
      if (out->mu) out->mu[b] = mu[b];
      if (out->d_current) out->d_current[b] = dcur[b];
    }
    if (out->x) memcpy(out->x, xo.data(), xo.size() * 8);
    if (out->l) memcpy(out->l, lo.data(), lo.size() * 8);
    if (out->L) memcpy(out->L, Lo.data(), Lo.size() * 8);
    if (cap) memcpy(out->eps_hist, eps.data(), B * cap * 16);
    return 0;
  }
  if (g_coop) {
    std::vector<double> xo((size_t)n * (N + 1) * B), lo((size_t)m * N * B), Lo((size_t)m * n * N * B, 0.0);
    P.perm = nullptr; P.xo = xo.data(); P.lo = lo.data(); P.Lo = Lo.data();
    int rc2 = dispatch(desc->model_id, cost_id, [&](auto D, auto CT) {
      using DD = decltype(D);
      std::vector<double> traj(coop_traj_doubles(DD::n, DD::m, N));
      if (g_coop == 2) {  // the two-warp speculative variant: the two "warps" of a round run one after the other
        std::vector<double> tr2(coop2_traj_doubles(DD::n, DD::m, N));
        for (size_t b = 0; b < B; ++b) {
          const size_t p = b / (size_t)P.K;
          if (P.active && !P.active[p]) continue;
          CoopWs<DD::n, DD::m> w2[2];
          Coop2Traj t;
          t.n = DD::n; t.m = DD::m; t.N = N;
          t.X = tr2.data(); t.U = t.X + (size_t)3 * (N + 1) * DD::n; t.Lg = t.U + (size_t)3 * N * DD::m; t.DL = t.Lg + (size_t)2 * N * DD::m * DD::n;
          const double* cp = P.cost_params + (P.cp_count > 1 ? p * (size_t)P.ncp : 0);
          SpecState S;
          spec_state_init(P, S);
          coop_ws_init(0, w2[0]);
          coop_ws_init(0, w2[1]);
          while (!S.done) {
            SpecLaneRes res[2];
            for (int g = 0; g < 2; ++g) res[g] = coop2_warp_work<DD, decltype(CT)>(0, g, P, S, cp, P.theta[b], p, w2[g], t);
            spec_decide<2>(P, S, res, b, true);
          }
          coop2_write_outputs(P, t, S, b, 0, 1);
        }
        return;
      }
      for (size_t b = 0; b < B; ++b) {
        CoopWs<DD::n, DD::m> w;
        CoopTraj tj;
        tj.X = traj.data(); tj.U = tj.X + (size_t)2 * (N + 1) * DD::n; tj.Lg = tj.U + (size_t)2 * N * DD::m; tj.DL = tj.Lg + (size_t)N * DD::m * DD::n;
        int cur = 0;
        if (coop_solve_instance<DD, decltype(CT)>(0, P, b, w, tj, cur)) coop_write_outputs<DD::n, DD::m>(0, P, b, tj, cur);
      }
    });
    if (rc2) return rc2;
    for (size_t b = 0; b < B; ++b) {
      if (out->value) out->value[b] = value[b];
      if (out->status) out->status[b] = status[b];
      if (out->iters) out->iters[b] = iters[b];
      if (out->trials) out->trials[b] = trials[b];
      if (out->restarts) out->restarts[b] = restarts[b];
      if (out->mu) out->mu[b] = mu[b];
      if (out->d_current) out->d_current[b] = dcur[b];
    }
    if (out->x) memcpy(out->x, xo.data(), xo.size() * 8);
    if (out->l) memcpy(out->l, lo.data(), lo.size() * 8);
    if (out->L) memcpy(out->L, Lo.data(), Lo.size() * 8);
    if (cap) memcpy(out->eps_hist, eps.data(), B * cap * 16);
    return 0;
  }
  // g_dynamic: emulate the persistent kernel (lane-level refill).  Threads run one after another here, so slot 0
  // drains most of the queue: a maximal test of workspace-slot reuse across instances.
  unsigned int queue = 0;
  std::vector<double> xo, lo, Lo;
  if (g_dynamic) {
    xo.assign((size_t)n * (N + 1) * B, 0.0); lo.assign((size_t)m * N * B, 0.0); Lo.assign((size_t)m * n * N * B, 0.0);
    P.queue = &queue; P.xo = xo.data(); P.lo = lo.data(); P.Lo = Lo.data();
  }
  int rc = dispatch(desc->model_id, cost_id, [&](auto D, auto CT) {
    double stage_area[2 * RL_STAGE_NV];
    Stage sg; sg.base = stage_area; sg.stride = 1;  // exercises the staged code path on the host
    if (g_dynamic) { for (size_t b = 0; b < std::min<size_t>(B, 3); ++b) solve_dynamic<decltype(D), decltype(CT)>(P, b, sg); }
    else { for (size_t b = 0; b < B; ++b) solve_instance<decltype(D), decltype(CT)>(P, b, sg); }
  });
  if (!rc && g_dynamic) {
    for (size_t b = 0; b < B; ++b) {
      if (out->value) out->value[b] = value[b];
      if (out->status) out->status[b] = status[b];
      if (out->iters) out->iters[b] = iters[b];
      if (out->trials) out->trials[b] = trials[b];
      if (out->restarts) out->restarts[b] = restarts[b];
      if (out->mu) out->mu[b] = mu[b];
      if (out->d_current) out->d_current[b] = dcur[b];
    }
    if (out->x) memcpy(out->x, xo.data(), xo.size() * 8);
    if (out->l) memcpy(out->l, lo.data(), lo.size() * 8);
    if (out->L) memcpy(out->L, Lo.data(), Lo.size() * 8);
    if (cap) memcpy(out->eps_hist, eps.data(), B * cap * 16);
    return 0;
  }
  if (rc) return rc;
  for (size_t b = 0; b < B; ++b) {
    if (out->value) out->value[b] = value[b];
    if (out->status) out->status[b] = status[b];
    if (out->iters) out->iters[b] = iters[b];
    if (out->trials) out->trials[b] = trials[b];
    if (out->restarts) out->restarts[b] = restarts[b];
    if (out->mu) out->mu[b] = mu[b];
    if (out->d_current) out->d_current[b] = dcur[b];
    size_t c = cur[b], inst = P.perm ? (size_t)P.perm[b] : b;  // b is the slot here
    const size_t tb = b >> 5, ln = b & 31;
    if (out->x) for (int e = 0; e < n * (N + 1); ++e) out->x[inst * n * (N + 1) + e] = X[((tb * wl.rec + c * (N + 1) * n) + e) * 32 + ln];
    if (out->l) for (int e = 0; e < m * N; ++e) out->l[inst * m * N + e] = U[((tb * wl.rec + c * N * m) + e) * 32 + ln];
    if (out->L) for (int e = 0; e < m * n * N; ++e) out->L[inst * m * n * N + e] = Lg[(tb * wl.rec + e) * 32 + ln];
  }
  if (cap) memcpy(out->eps_hist, eps.data(), B * cap * 16);
  return 0;
}

int32_t hostemu_ce_costs(void*, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                         const ratilqr_batch_in* in, double kl, double* cost, int32_t* status) {
  size_t B = (size_t)in->P * in->K;
  std::vector<double> value(B);
  std::vector<int32_t> st(B);
  ratilqr_ileqg_out out;
  memset(&out, 0, sizeof(out));
  out.value = value.data(); out.status = st.data();
  int rc = hostemu_ileqg_solve_batch(nullptr, desc, opts, in, &out);
  if (rc) return rc;
  for (size_t b = 0; b < B; ++b) { cost[b] = st[b] == 0 ? value[b] + kl / in->theta[b] : HUGE_VAL; if (status) status[b] = st[b]; }
  return 0;
}

int32_t hostemu_rollout_open_batch(void*, const ratilqr_problem_desc* d, int32_t B, const double* x0, const double* u,
                                   double* x, int32_t* status) {
  return dispatch(d->model_id, d->cost_id, [&](auto D, auto) {
    constexpr int n = decltype(D)::n, m = decltype(D)::m;
    for (int b = 0; b < B; ++b) {
      int st = comp_rollout_open<decltype(D)>(d->model_params, d->N, x0 + (size_t)b * n, u + (size_t)b * m * d->N, x + (size_t)b * n * (d->N + 1));
      if (status) status[b] = st;
    }
  });
}

int32_t hostemu_rollout_closed_batch(void*, const ratilqr_problem_desc* d, int32_t B, const double* xbar, const double* l,
                                     const double* L, double* xn, double* un, int32_t* status) {
  return dispatch(d->model_id, d->cost_id, [&](auto D, auto CT) {
    constexpr int n = decltype(D)::n, m = decltype(D)::m;
    int N = d->N;
    for (int b = 0; b < B; ++b) {
      int st = comp_rollout_closed<decltype(D), decltype(CT)>(d->model_params, d->cost_params, N, xbar + (size_t)b * n * (N + 1),
                                                             l + (size_t)b * m * N, L + (size_t)b * m * n * N, nullptr,
                                                             xn + (size_t)b * n * (N + 1), un + (size_t)b * m * N, nullptr);
      if (status) status[b] = st;
    }
  });
}

int32_t hostemu_integrate_cost_batch(void*, const ratilqr_problem_desc* d, int32_t B, const double* x, const double* u,
                                     double* cost, int32_t* status) {
  return dispatch(d->model_id, d->cost_id, [&](auto D, auto CT) {
    constexpr int n = decltype(D)::n, m = decltype(D)::m;
    int N = d->N;
    for (int b = 0; b < B; ++b) {
      double c = HUGE_VAL;
      int st = comp_integrate_cost<decltype(D), decltype(CT)>(d->cost_params, N, x + (size_t)b * n * (N + 1), u + (size_t)b * m * N, &c);
      cost[b] = st ? HUGE_VAL : c;
      if (status) status[b] = st;
    }
  });
}

int32_t hostemu_linearize_batch(void*, const ratilqr_problem_desc* d, int32_t B, const double* x, const double* u,
                                double* q, double* qv, double* Q, double* r, double* R, double* Pm, double* A, double* Bm,
                                int32_t* status) {
  if (hm_check_desc(d, true)) return -1;
  return dispatch(d->model_id, d->cost_id, [&](auto D, auto CT) {
    constexpr int n = decltype(D)::n, m = decltype(D)::m;
    int N = d->N;
    for (int b = 0; b < B; ++b) {
      int worst = 0;
      for (int k = 0; k <= N; ++k) {
        int st = comp_linearize_stage<decltype(D), decltype(CT)>(
            d->model_params, d->cost_params, N, k, x + (size_t)b * n * (N + 1), u + (size_t)b * m * N, q + (size_t)b * (N + 1),
            qv + (size_t)b * n * (N + 1), Q + (size_t)b * n * n * (N + 1), r + (size_t)b * m * N, R + (size_t)b * m * m * N,
            Pm + (size_t)b * m * n * N, A + (size_t)b * n * n * N, Bm + (size_t)b * n * m * N);
        if (st > worst) worst = st;
      }
      if (status) status[b] = worst;
    }
  });
}

int32_t hostemu_riccati_batch_tv(void*, int32_t n, int32_t m, int32_t N, int32_t B, int32_t optimise, const double* q,
                                 const double* qv, const double* Q, const double* r, const double* R, const double* Pm,
                                 const double* A, const double* Bm, const double* W, int32_t W_tv, const double* theta,
                                 double mu_min, double delta_0, double* mu, double* delta, double* L, double* dl, double* s,
                                 double* sv, double* S, int32_t* status, int32_t* restarts);
int32_t hostemu_riccati_batch(void* c, int32_t n, int32_t m, int32_t N, int32_t B, int32_t optimise, const double* q,
                              const double* qv, const double* Q, const double* r, const double* R, const double* Pm,
                              const double* A, const double* Bm, const double* W, const double* theta, double mu_min,
                              double delta_0, double* mu, double* delta, double* L, double* dl, double* s, double* sv,
                              double* S, int32_t* status, int32_t* restarts) {
  return hostemu_riccati_batch_tv(c, n, m, N, B, optimise, q, qv, Q, r, R, Pm, A, Bm, W, 0, theta, mu_min, delta_0, mu, delta, L,
                                  dl, s, sv, S, status, restarts);
}
int32_t hostemu_riccati_batch_tv(void*, int32_t n, int32_t m, int32_t N, int32_t B, int32_t optimise, const double* q,
                                 const double* qv, const double* Q, const double* r, const double* R, const double* Pm,
                                 const double* A, const double* Bm, const double* W, int32_t W_tv, const double* theta,
                                 double mu_min, double delta_0, double* mu, double* delta, double* L, double* dl, double* s,
                                 double* sv, double* S, int32_t* status, int32_t* restarts) {
  rlh::WPrep wp;
  W_tv = W_tv ? 1 : 0;
  if (!rlh::prep_W(n, N, W, W_tv, wp)) return -2;
#define R_(NN, MM) if (n == NN && m == MM) { ric<NN, MM>(N, B, optimise, q, qv, Q, r, R, Pm, A, Bm, wp, W_tv, theta, mu_min, delta_0, mu, delta, L, dl, s, sv, S, status, restarts); return 0; }
  R_(2, 2) R_(2, 1) R_(4, 2) R_(4, 1) R_(12, 4)
#undef R_
  return -5;
}

static int32_t mc_rollout_impl(const ratilqr_problem_desc* d, int32_t P, const double* xbar, const double* l,
                              const double* L, int32_t n_samples, const double* noise, const ratilqr_noise_mixture* true_noise,
                              uint64_t seed, double* J, double* stats, double* x_out) {
  rlh::WPrep wp;
  if (!rlh::prep_W(d->n, d->N, d->W, d->W_time_varying, wp)) return -2;
  rlh::MixPrep mp;
  MixtureView mx{0, nullptr, nullptr, nullptr};
  if (true_noise) {
    if (rlh::prep_mixture(d->n, true_noise, mp)) return -2;
    mx = MixtureView{mp.k, mp.cumw.data(), mp.mean.data(), mp.chol.data()};
  }
  return dispatch(d->model_id, d->cost_id, [&](auto D, auto CT) {
    constexpr int n = decltype(D)::n, m = decltype(D)::m;
    int N = d->N;
    for (int p = 0; p < P; ++p)
      for (int s = 0; s < n_samples; ++s) {
        size_t gi = (size_t)p * n_samples + s;
        const double* cp = d->cost_params + (d->cost_params_count > 1 ? (size_t)p * d->n_cost_params : 0);
        double c = HUGE_VAL;
        std::vector<double> w((size_t)n * N);
        if (noise) memcpy(w.data(), noise + gi * n * N, sizeof(double) * n * N);
        else if (mx.k > 0) for (int k = 0; k < N; ++k) philox_mixture_noise<n>(seed, gi, (uint32_t)k, mx, &w[(size_t)k * n]);
        else for (int k = 0; k < N; ++k) philox_noise<n>(seed, gi, (uint32_t)k, 0, 1.0, wp.cholW.data() + (d->W_time_varying ? (size_t)k * n * n : 0), &w[(size_t)k * n]);
        int st = comp_rollout_closed<decltype(D), decltype(CT)>(d->model_params, cp, N, xbar + (size_t)p * n * (N + 1), l + (size_t)p * m * N,
                                                               L + (size_t)p * m * n * N, w.data(),
                                                               x_out ? x_out + gi * n * (N + 1) : nullptr, nullptr, &c);
        J[gi] = st ? HUGE_VAL : c;
      }
    (void)stats;
  });
}

int32_t hostemu_mc_rollout(void*, const ratilqr_problem_desc* d, int32_t P, const double* xbar, const double* l,
                           const double* L, int32_t n_samples, const double* noise, uint64_t seed, double, double* J,
                           double* stats, double* x_out) {
  return mc_rollout_impl(d, P, xbar, l, L, n_samples, noise, nullptr, seed, J, stats, x_out);
}
int32_t hostemu_mc_rollout_true_model(void*, const ratilqr_problem_desc* d, int32_t P, const double* xbar, const double* l,
                                      const double* L, int32_t n_samples, const ratilqr_noise_mixture* true_noise,
                                      uint64_t seed, double, double* J, double* stats, double* x_out) {
  return mc_rollout_impl(d, P, xbar, l, L, n_samples, nullptr, true_noise, seed, J, stats, x_out);
}

int32_t hostemu_pets_costs(void*, const ratilqr_problem_desc* d, const ratilqr_generative_desc* gen, const double* x0,
                           const double* controls, int32_t C, int32_t particles, const double* noise, uint64_t seed,
                           double* cost) {
  rlh::WPrep wp;
  if (!rlh::prep_W(d->n, d->N, d->W, 0, wp)) return -2;
  rlh::MixPrep mp;
  MixtureView mx{0, nullptr, nullptr, nullptr};
  if (gen && gen->use_true_model && gen->true_model) {
    if (rlh::prep_mixture(d->n, gen->true_model, mp)) return -2;
    mx = MixtureView{mp.k, mp.cumw.data(), mp.mean.data(), mp.chol.data()};
  }
  return dispatch(d->model_id, d->cost_id, [&](auto D, auto CT) {
    constexpr int n = decltype(D)::n, m = decltype(D)::m;
    int N = d->N;
    int ne = gen && gen->n_ensemble > 1 ? gen->n_ensemble : 1;
    int per = ne > 1 ? std::max(particles / ne, 1) : particles;
    for (int ii = 0; ii < C; ++ii) {
      double acc = 0.0;
      for (int kk = 0; kk < particles; ++kk) {
        const double* mp = d->model_params;
        if (gen && gen->ensemble_params && ne > 1) mp = gen->ensemble_params + (size_t)std::min(kk / per, ne - 1) * d->n_model_params;
        size_t gi = (size_t)ii * particles + kk;
        acc += comp_pets_particle<decltype(D), decltype(CT)>(mp, d->cost_params, N, x0, controls + (size_t)ii * m * N,
                                                            noise ? noise + gi * n * N : nullptr, seed, gi,
                                                            gen ? gen->noise_kind : 0, gen ? gen->noise_scale : 1.0, wp.cholW.data(), &mx);
      }
      cost[ii] = acc / particles;
    }
  });
}

}  // extern "C"
