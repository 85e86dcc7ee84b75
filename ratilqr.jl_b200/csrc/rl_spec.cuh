// rl_spec.cuh -- latency-oriented iLEQG solve for SMALL batches: G lanes per instance, used as speculative workers.
//
// One thread per instance (rl_core.cuh) is throughput-optimal, but its latency is a chain of sequential PASSES: per
// iLEQG iteration one optimising backward pass, then per line-search trial a closed-loop rollout and an evaluating
// backward pass (ileqg.jl:598-613, :494-592).  A batch that does not fill the GPU (a single MPC problem: <= 10 theta;
// configs[1]'s 1 x 1024 theta) leaves almost every lane of the machine idle while each instance walks that chain.
// Here an instance owns G = 2 T lanes of one warp and uses them to walk SEVERAL links of the chain at once:
//
//   * trial speculation: pair j (lanes 2j, 2j+1) rolls out and evaluates the line-search candidate eps * lambda^j.  The
//     T candidates of a round are pure functions of the current (x, l, L, dl), so evaluating them side by side and then
//     replaying the accept / reject rule of line_search! IN ORDER gives exactly the serial result (first acceptable one);
//   * pass speculation: while lane 2j EVALUATES candidate j (policy L of the current iteration, dl = 0), lane 2j+1 already
//     runs the OPTIMISING pass of the NEXT iteration on the same candidate trajectory.  If candidate j is accepted and the
//     solve goes on, that pass is the next iteration's solve_approximate_dp! (same inputs, same mu), so the next round can
//     start with its rollouts at once; otherwise it is discarded.  Both lanes execute ONE instruction stream (the stage
//     code takes optimise-or-evaluate as a per-lane flag), so the pair costs the time of one pass.
//
// An iteration with r rejected trials thus costs ceil((r+1)/T) x (rollout + one pass) instead of
// (one pass) + (r+1) x (rollout + one pass).  Every number is produced by the same per-instance arithmetic as in
// rl_core.cuh (riccati_stage, the model/cost structs), evaluated on the same inputs in the same order: results are
// bit-identical to the one-thread-per-instance kernel (tests/test_spec_kernel.py).
//
// Storage: every lane has its own column of the warp-tiled workspace -- two trajectory buffers X[2], U[2] and two policy
// buffers Lg[2], DL[2].  The CURRENT trajectory is (column cur_col, buffer cur_buf) of the group, the current policy
// (pol_col, pol_buf); a lane writes its candidate / speculative policy into the buffer of its own column that does not
// hold a current object.  Lanes read other lanes' columns only across a __syncwarp().
#pragma once
#include "rl_core.cuh"

namespace rl {

struct SpecLaneRes {
  int st_roll;   // status of the rollout (DOMAIN)
  int rc;        // eval lane: 0 / M-not-PD code / DOMAIN ; opt lane: status of the speculative optimising pass
  double nw;     // eval lane: merit value s_array[1]
  double dmax;   // eval lane: maximum(norm.(l .- u_new))
  double mu, delta;  // opt lane: regularisation after its pass
  int restarts;      // opt lane: cumulative increase_mu_and_delta! count after its pass
};

struct SpecState {  // per instance; every lane of the group holds an identical copy
  int cur_col, cur_buf, pol_col, pol_buf, has_pol;
  int iters, trials, restarts, status, count;
  double mu, delta, d_current, value, eps_init, eps;
  bool init, done;
};

// column pointers of the warp-tiled workspace (element stride RL_TILE); Lg / DL are double-buffered in this mode
struct SpecCols {
  const SolveParams* P;
  size_t tile, lane0;  // warp tile of the group, lane index of the group's column 0 inside the tile
  int n, m, N;
  RL_HD double* X(int col, int buf) const { return P->X + (tile * P->rec + (size_t)buf * (N + 1) * n) * RL_TILE + lane0 + col; }
  RL_HD double* U(int col, int buf) const { return P->U + (tile * P->rec + (size_t)buf * N * m) * RL_TILE + lane0 + col; }
  RL_HD double* L(int col, int buf) const { return P->Lg + (tile * P->rec + (size_t)buf * N * m * n) * RL_TILE + lane0 + col; }
  RL_HD double* DL(int col, int buf) const { return P->DL + (tile * P->rec + (size_t)buf * N * m) * RL_TILE + lane0 + col; }
};

// backward pass over the trajectory (Xb, Ub), fused with approximate_model; is_opt: solve_approximate_dp! incl. the
// mu-restart loop (ileqg.jl:341-406), gains / dl written to (Ldst, DLdst); else solve_approximate_dp (:412-465) of the
// policy Lsrc (zeroL: L = 0) with dl = zeros(m) (:447-451).  Same staging scheme as rl::backward_pass.
// SM: staging decided at run time by sg.base (-1: host emulation) or at compile time (0 / 1: the kernel dispatches once)
template <class D, class CT, int SM = -1>
RL_HD int spec_backward(const SolveParams& P, const double* cp, double theta, bool is_opt, bool zeroL, const double* Xb,
                        const double* Ub, const double* Lsrc, double* Ldst, double* DLdst, double& mu, double& delta,
                        int& restarts, double& value, Stage sg) {
  constexpr int n = D::n, m = D::m;
  using Tr = StageTraits<D, CT>;
  constexpr size_t B = RL_TILE;
  const int N = P.N;
  const bool staged = UseStage<D>::value && (SM < 0 ? sg.base != nullptr : SM == 1);
  const bool needL = !is_opt && !zeroL;
  auto fetch = [&](int k) {
    double* s0 = sg.base + (size_t)(k & 1) * RL_STAGE_NV * sg.stride;
    for (int i = 0; i < n; ++i) rl_stage_put(s0 + (size_t)i * sg.stride, Xb + ((size_t)k * n + i) * B);
    for (int i = 0; i < m; ++i) rl_stage_put(s0 + (size_t)(n + i) * sg.stride, Ub + ((size_t)k * m + i) * B);
    if (needL) for (int i = 0; i < m * n; ++i) rl_stage_put(s0 + (size_t)(n + m + i) * sg.stride, Lsrc + ((size_t)k * m * n + i) * B);
    rl_stage_commit();
  };
  while (true) {
    double S[n * n], sv[n], s;
    if (staged) fetch(N - 1);
    {
      double x[n], Q[n * n];
      ld_vec<n>(Xb + (size_t)N * n * B, B, x);
      if (!CT::terminal(cp, x, true, s, sv, Q)) { if (staged) rl_stage_wait(); return RATILQR_ST_DOMAIN; }
      for (int i = 0; i < n; ++i) for (int j = i; j < n; ++j) {
        double v = (Tr::q_kind(i, j) == 0) ? 0.0 : Q[i + j * n];
        S[i + j * n] = v; S[j + i * n] = v;
      }
    }
    bool restart = false;
    int fail = 0;
    double detprod = 1.0, logacc = 0.0;
    for (int k = N - 1; k >= 0; --k) {
      double x[n], u[m], q, qv[n], Q[n * n], r[m], R[m * m], Pm[m * n], A[n * n], Bm[n * m], L[m * n], dl[m];
      if (staged) {
        rl_stage_wait();
        const double* s0 = sg.base + (size_t)(k & 1) * RL_STAGE_NV * sg.stride;
        for (int i = 0; i < n; ++i) x[i] = s0[(size_t)i * sg.stride];
        for (int i = 0; i < m; ++i) u[i] = s0[(size_t)(n + i) * sg.stride];
        for (int i = 0; i < m * n; ++i) L[i] = needL ? s0[(size_t)(n + m + i) * sg.stride] : 0.0;
        if (k > 0) fetch(k - 1);
      } else {
        ld_vec<n>(Xb + (size_t)k * n * B, B, x);
        ld_vec<m>(Ub + (size_t)k * m * B, B, u);
        for (int i = 0; i < m * n; ++i) L[i] = needL ? Lsrc[((size_t)k * m * n + i) * B] : 0.0;
      }
      for (int i = 0; i < m; ++i) dl[i] = 0.0;
      if (!CT::stage(cp, k, x, u, true, q, qv, Q, r, R, Pm)) { fail = RATILQR_ST_DOMAIN; break; }
      // branch-free fast paths of rsqrt / sincos inside the stage, library routines on the recorded slow-path flag (as in
      // rl::backward_pass; device only: the host emulation runs the plain routines)
#if defined(__CUDA_ARCH__)
      constexpr bool FRS = RL_DEFER_PD && RL_FAST_RSQRT;
#else
      constexpr bool FRS = false;
#endif
      bool lin_slow = false;
      if (FRS) jac_nb<D>(P.mp, x, u, A, Bm, lin_slow); else D::jac(P.mp, x, u, A, Bm);
      const size_t wo = P.W_tv ? (size_t)k * n * n : 0;
      int rc = riccati_stage<Tr, false, true, true, FRS>(theta, mu, P.W + wo, P.Winv + wo, P.detW[P.W_tv ? k : 0], S, sv, s, q, qv, Q, r,
                                                         R, Pm, A, Bm, L, dl, RL_FUSED ? &detprod : nullptr, is_opt, lin_slow);
      if (FRS && rc == 3) {
        if (lin_slow) D::jac(P.mp, x, u, A, Bm);
        rc = riccati_stage<Tr, false, true, true>(theta, mu, P.W + wo, P.Winv + wo, P.detW[P.W_tv ? k : 0], S, sv, s, q, qv, Q, r,
                                                  R, Pm, A, Bm, L, dl, RL_FUSED ? &detprod : nullptr, is_opt);
      }
      if (rc == 1) { fail = is_opt ? RATILQR_ST_M_NOT_PD_OPT : RATILQR_ST_M_NOT_PD_INIT; break; }
      if (RL_FUSED && !(detprod > 1e-250 && detprod < 1e250)) { logacc += log(detprod); detprod = 1.0; }
      if (rc == 2) {  // optimising lanes only: increase_mu_and_delta! and restart the sweep (:372-378)
        delta = fmax(P.delta_0, delta * P.delta_0);
        mu = fmax(P.mu_min, mu * delta);
        restarts++;
        if (!(mu < 1e300)) { fail = RATILQR_ST_MU_OVERFLOW; break; }
        restart = true;
        break;
      }
      if (is_opt) {
        st_vec<m * n>(Ldst + (size_t)k * m * n * B, B, L);  // :380
        st_vec<m>(DLdst + (size_t)k * m * B, B, dl);
      }
    }
    if (staged) rl_stage_wait();  // drain (non-trivial after a restart / failure)
    if (fail) return fail;
    if (!restart) {
      if (RL_FUSED && theta != 0.0) s = s - (1 / (2 * theta)) * (logacc + log(detprod));
      value = s;
      return 0;
    }
  }
}

// closed-loop rollout around (Xc, Uc) with l + eps*dl and gains L of the policy (Lp, DLp) into (Xn, Un)
// (ileqg.jl:509, :62-87); init: open-loop rollout of the initial controls (:225-228).  Mirrors rl::rollout_candidate.
template <class D, int SM = -1>
RL_HD int spec_rollout(const SolveParams& P, const double* Xc, const double* Uc, const double* Lp, const double* DLp,
                       double* Xn, double* Un, double eps, bool init, double& dmax, Stage sg) {
  constexpr int n = D::n, m = D::m;
  constexpr size_t B = RL_TILE;
  const int N = P.N;
  const bool staged = UseStage<D>::value && (SM < 0 ? sg.base != nullptr : SM == 1);
  auto fetch = [&](int k) {
    double* s0 = sg.base + (size_t)(k & 1) * RL_STAGE_NV * sg.stride;
    for (int i = 0; i < n; ++i) rl_stage_put(s0 + (size_t)i * sg.stride, Xc + ((size_t)k * n + i) * B);
    for (int i = 0; i < m; ++i) rl_stage_put(s0 + (size_t)(n + i) * sg.stride, Uc + ((size_t)k * m + i) * B);
    for (int i = 0; i < m; ++i) rl_stage_put(s0 + (size_t)(n + m + i) * sg.stride, DLp + ((size_t)k * m + i) * B);
    for (int i = 0; i < m * n; ++i) rl_stage_put(s0 + (size_t)(n + 2 * m + i) * sg.stride, Lp + ((size_t)k * m * n + i) * B);
    rl_stage_commit();
  };
  if (staged) {
    fetch(0);
    if (N > 1) fetch(1); else rl_stage_commit();
  }
  double x[n];
  ld_vec<n>(Xc, B, x);
  st_vec<n>(Xn, B, x);
  double best = -rl_inf();
  bool has_nan = false;
  for (int k = 0; k < N; ++k) {
    double xb[n], l[m], dl[m], L[m * n], u[m], xn[n];
    if (staged) {
      rl_stage_wait1();
      const double* s0 = sg.base + (size_t)(k & 1) * RL_STAGE_NV * sg.stride;
      for (int i = 0; i < n; ++i) xb[i] = s0[(size_t)i * sg.stride];
      for (int i = 0; i < m; ++i) l[i] = s0[(size_t)(n + i) * sg.stride];
      for (int i = 0; i < m; ++i) dl[i] = s0[(size_t)(n + m + i) * sg.stride];
      for (int i = 0; i < m * n; ++i) L[i] = s0[(size_t)(n + 2 * m + i) * sg.stride];
      rl_stage_fence();
      if (k + 2 < N) fetch(k + 2); else rl_stage_commit();
    } else {
      ld_vec<n>(Xc + (size_t)k * n * B, B, xb);
      ld_vec<m>(Uc + (size_t)k * m * B, B, l);
      ld_vec<m>(DLp + (size_t)k * m * B, B, dl);
      ld_vec<m * n>(Lp + (size_t)k * m * n * B, B, L);
    }
    double dx[n];
    for (int i = 0; i < n; ++i) dx[i] = x[i] - xb[i];
    double acc = 0.0;
    for (int j = 0; j < m; ++j) {
      double uj;
      if (RL_FUSED) {
        uj = l[j] + eps * dl[j];
        for (int i = 0; i < n; ++i) uj = rl_fma(L[j + i * m], dx[i], uj);
      } else {
        double a = L[j] * dx[0];
        for (int i = 1; i < n; ++i) a = rl_fma(L[j + i * m], dx[i], a);
        uj = (l[j] + eps * dl[j]) + a;
      }
      u[j] = init ? l[j] : uj;
      double dd = l[j] - u[j];
      acc = (j == 0) ? dd * dd : rl_fma(dd, dd, acc);
    }
    if (acc != acc) has_nan = true;
    if (acc > best) best = acc;
    bool f_ok;
#if defined(__CUDA_ARCH__)
    if (RL_FAST_RSQRT) { bool f_slow = false; f_ok = f_nb<D>(P.mp, x, u, xn, f_slow); if (f_slow) f_ok = D::f(P.mp, x, u, xn); }
    else
#endif
      f_ok = D::f(P.mp, x, u, xn);
    if (!f_ok) { if (staged) rl_stage_wait(); return RATILQR_ST_DOMAIN; }
    st_vec<m>(Un + (size_t)k * m * B, B, u);
    st_vec<n>(Xn + (size_t)(k + 1) * n * B, B, xn);
    for (int i = 0; i < n; ++i) x[i] = xn[i];
  }
  if (staged) rl_stage_wait();
  dmax = has_nan ? (double)NAN : sqrt(best);
  return 0;
}

RL_HD void spec_state_init(const SolveParams& P, SpecState& S) {  // initialize! :216-219
  S.cur_col = 0; S.cur_buf = 0; S.pol_col = 0; S.pol_buf = 0; S.has_pol = 0;
  S.iters = 0; S.trials = 0; S.restarts = 0; S.status = 0; S.count = 0;
  S.mu = 0.0; S.delta = P.delta_0; S.d_current = rl_inf(); S.value = rl_inf();
  S.eps_init = P.eps_init; S.eps = 0.0;
  S.init = true; S.done = false;
}

// the buffer of column g that holds neither the current trajectory / the current policy
RL_HD int spec_free_traj_buf(const SpecState& S, int g) { return (g == S.cur_col) ? (S.cur_buf ^ 1) : 0; }
RL_HD int spec_free_pol_buf(const SpecState& S, int g) { return (S.has_pol && g == S.pol_col) ? (S.pol_buf ^ 1) : 0; }

// one round of lane g of the group: rollout of candidate j = g/2, then its evaluating (g even) / optimising (g odd) pass
template <class D, class CT, int SM = -1>
RL_HD SpecLaneRes spec_lane_work(const SolveParams& P, const SpecCols& C, int g, const SpecState& S, const double* cp,
                                 double theta, size_t p, Stage sg) {
  constexpr int n = D::n, m = D::m;
  const int N = P.N;
  SpecLaneRes r;
  r.st_roll = 0; r.rc = 0; r.nw = rl_inf(); r.dmax = rl_inf(); r.mu = S.mu; r.delta = S.delta; r.restarts = S.restarts;
  const bool is_opt = (g & 1) != 0;
  const int fb = S.init ? 0 : spec_free_traj_buf(S, g), pb = S.init ? 0 : spec_free_pol_buf(S, g);
  double dmax = rl_inf();
  if (S.init) {  // l_array = copy(u_array) (:228), x_0: into buffer 1 of the lane's own column, rolled out into buffer 0
    const double* x0 = P.x0 + (P.x0_count > 1 ? p * n : 0);
    const double* ui = P.u_init + (P.u_count > 1 ? p * (size_t)m * N : 0);
    double* Xc = C.X(g, 1);
    double* Uc = C.U(g, 1);
    for (int i = 0; i < n; ++i) Xc[(size_t)i * RL_TILE] = x0[i];
    for (int k = 0; k < N; ++k)
      for (int j = 0; j < m; ++j) Uc[((size_t)k * m + j) * RL_TILE] = ui[(size_t)k * m + j];
    r.st_roll = spec_rollout<D, SM>(P, Xc, Uc, C.L(g, 0), C.DL(g, 0), C.X(g, 0), C.U(g, 0), 0.0, true, dmax, sg);
  } else {
    double eps = S.eps;
    for (int i = 0; i < (g >> 1); ++i) eps *= P.lambda;  // the eps the serial loop would reach at its (g/2)-th further trial
    r.st_roll = spec_rollout<D, SM>(P, C.X(S.cur_col, S.cur_buf), C.U(S.cur_col, S.cur_buf), C.L(S.pol_col, S.pol_buf),
                                C.DL(S.pol_col, S.pol_buf), C.X(g, fb), C.U(g, fb), eps, false, dmax, sg);
  }
  r.dmax = dmax;
  if (r.st_roll) return r;
  double val = rl_inf();
  r.rc = spec_backward<D, CT, SM>(P, cp, theta, is_opt, S.init, C.X(g, fb), C.U(g, fb), C.L(S.pol_col, S.pol_buf), C.L(g, pb),
                              C.DL(g, pb), r.mu, r.delta, r.restarts, val, sg);
  r.nw = val;
  return r;
}

// replay of initialize! / line_search! / the stop rule of solve! (ileqg.jl:214-236, :504-592, :640-654) on the T results of
// a round, in trial order; res[2j] is the evaluating lane of candidate j, res[2j+1] its optimising lane.
// `writer`: exactly one lane of the group records the eps history.
template <int G>
RL_HD void spec_decide(const SolveParams& P, SpecState& S, const SpecLaneRes* res, size_t inst, bool writer) {
  constexpr int T = G / 2;
  // commit the speculative optimising pass of candidate j as this iteration's solve_approximate_dp! (step! :598-613)
  auto start_iteration = [&](int j) {
    const SpecLaneRes& o = res[2 * j + 1];
    S.iters++;
    S.mu = o.mu; S.delta = o.delta; S.restarts = o.restarts;
    if (o.rc) { S.status = o.rc; S.done = true; return; }
    const int pb = spec_free_pol_buf(S, 2 * j + 1);
    S.pol_col = 2 * j + 1; S.pol_buf = pb; S.has_pol = 1;
    S.eps = S.eps_init; S.count = 0;
  };
  if (S.init) {
    const SpecLaneRes& e = res[0];
    if (e.st_roll) { S.status = e.st_roll; S.done = true; return; }
    if (e.rc == RATILQR_ST_DOMAIN) { S.status = e.rc; S.done = true; return; }
    if (e.rc) { S.status = RATILQR_ST_M_NOT_PD_INIT; S.done = true; return; }
    S.value = e.nw; S.cur_col = 0; S.cur_buf = 0; S.init = false;
    start_iteration(0);
    return;
  }
  for (int j = 0; j < T; ++j) {
    const SpecLaneRes& e = res[2 * j];
    S.count++;  // :504
    if (S.eps == 0.0 || S.count > 4000) { S.status = RATILQR_ST_LINESEARCH_HANG; S.done = true; return; }
    if (e.st_roll) { S.status = e.st_roll; S.done = true; return; }
    if (e.rc == RATILQR_ST_DOMAIN) { S.status = e.rc; S.done = true; return; }
    if (e.rc) { S.eps *= P.lambda; continue; }  // :529-535
    if (writer && P.eps_hist && S.trials < P.eps_hist_cap) {
      double* h = P.eps_hist + (inst * P.eps_hist_cap + S.trials) * 2;
      h[0] = S.eps; h[1] = e.nw - S.value;
    }
    S.trials++;
    bool accepted = isapprox_default(e.nw, S.value) || e.nw < S.value;  // :538
    if (!accepted) {
      S.eps *= P.lambda;
      if (S.eps < P.eps_min) accepted = true;  // :558-575
    }
    if (!accepted) continue;
    const int fb = spec_free_traj_buf(S, 2 * j);
    S.d_current = e.dmax; S.value = e.nw; S.cur_col = 2 * j; S.cur_buf = fb;
    if (P.eps_auto) {  // :582-591
      if (S.count == 1) S.eps_init = fmin(P.eps_init, S.eps / P.lambda);
      else { double ee = S.eps; while (ee < P.eps_min) ee = ee / P.lambda; S.eps_init = ee; }
    }
    if (P.d > S.d_current && S.mu <= P.mu_min) { S.done = true; return; }  // :642
    if (S.iters == P.iter_max) { S.done = true; return; }                  // :648
    start_iteration(j);
    return;
  }
  // no candidate of this round accepted: the next round evaluates eps * lambda^T ...
}

// results and, if asked for, x_array / l_array / L_array in host layout; the copy is spread over the G lanes
RL_HD void spec_write_outputs(const SolveParams& P, const SpecCols& C, const SpecState& S, size_t inst, int g, int G) {
  const int n = C.n, m = C.m, N = C.N;
  if (g == 0) {
    P.value[inst] = S.status ? rl_inf() : S.value;
    P.status[inst] = S.status;
    P.iters[inst] = S.iters;
    P.trials[inst] = S.trials;
    P.restarts[inst] = S.restarts;
    P.mu_out[inst] = S.mu;
    P.d_out[inst] = S.d_current;
  }
  // a solve that failed in initialize! reports the rolled-out initial trajectory like the one-thread kernel (buffer 0 of col 0)
  if (P.xo) { const double* X = C.X(S.cur_col, S.cur_buf); for (int e = g; e < (N + 1) * n; e += G) P.xo[inst * (size_t)(N + 1) * n + e] = X[(size_t)e * RL_TILE]; }
  if (P.lo) { const double* U = C.U(S.cur_col, S.cur_buf); for (int e = g; e < N * m; e += G) P.lo[inst * (size_t)N * m + e] = U[(size_t)e * RL_TILE]; }
  if (P.Lo) {
    const double* L = C.L(S.pol_col, S.pol_buf);
    for (int e = g; e < N * m * n; e += G) P.Lo[inst * (size_t)N * m * n + e] = S.has_pol ? L[(size_t)e * RL_TILE] : 0.0;
  }
}

}  // namespace rl
