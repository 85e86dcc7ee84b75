// rl_coop2.cuh -- the warp-cooperative iLEQG solve (rl_coop.cuh: matrices and trajectories in shared memory, the outputs
// of every small dense operation spread over a warp) with PASS SPECULATION across two warps: the latency path of the
// large models (n > 6: the quadrotor), where a single Nelder-Mead candidate is a chain of ~30 sequential passes.
//
// A CTA of two warps owns one instance.  In every round warp 0 rolls out the line-search candidate and EVALUATES it
// (policy of the current iteration), while warp 1 rolls out the same candidate into its own buffer and runs the
// OPTIMISING pass of the next iteration on it (rl_spec.cuh explains why that is the next iteration's
// solve_approximate_dp! whenever the candidate is accepted).  The two warps run different instruction streams and meet
// only at the end of the round (__syncthreads), where every thread replays the accept / reject rule on both results
// (rl::spec_decide<2>, the very function the per-lane speculative kernel uses).  An accepted first trial -- the common
// case -- costs rollout + one pass instead of pass + rollout + pass.  Arithmetic per output element is that of
// rl_coop.cuh, so results are bit-identical to the one-warp kernel (tests/test_spec_kernel.py).
//
// Shared memory per CTA: two CoopWs (one per warp) and X[3], U[3] (current, free, warp 1's candidate), Lg[2], DL[2]
// (current and speculative policy): 75 KB for the quadrotor at N = 40.
#pragma once
#include "rl_coop.cuh"
#include "rl_spec.cuh"

namespace rl {

RL_HD size_t coop2_traj_doubles(int n, int m, int N) {
  return (size_t)3 * (N + 1) * n + (size_t)3 * N * m + (size_t)2 * N * m * n + (size_t)2 * N * m;
}

struct Coop2Traj {
  double *X, *U, *Lg, *DL;
  int n, m, N;
  // column 0 (warp 0) owns trajectory buffers 0 / 1, column 1 (warp 1) buffer 2; policies live in column 1's two buffers
  RL_HD double* Xb(int col, int buf) const { return X + (size_t)(col == 0 ? buf : 2) * (N + 1) * n; }
  RL_HD double* Ub(int col, int buf) const { return U + (size_t)(col == 0 ? buf : 2) * N * m; }
  RL_HD double* Lb(int buf) const { return Lg + (size_t)buf * N * m * n; }
  RL_HD double* DLb(int buf) const { return DL + (size_t)buf * N * m; }
};

// one round of warp g (0: evaluate, 1: optimise the next iteration) -- the warp-level twin of rl::spec_lane_work
template <class D, class CT>
RL_HD SpecLaneRes coop2_warp_work(int lane, int g, const SolveParams& P, const SpecState& S, const double* cp, double theta,
                                  size_t p, CoopWs<D::n, D::m>& w, const Coop2Traj& t) {
  constexpr int n = D::n, m = D::m;
  const int N = P.N;
  SpecLaneRes r;
  r.st_roll = 0; r.rc = 0; r.nw = rl_inf(); r.dmax = rl_inf(); r.mu = S.mu; r.delta = S.delta; r.restarts = S.restarts;
  const int fb = S.init ? 0 : spec_free_traj_buf(S, g), pb = S.init ? 0 : spec_free_pol_buf(S, g);
  CoopIO io;
  io.Xd = t.Xb(g, fb); io.Ud = t.Ub(g, fb);
  io.Lr = t.Lb(S.pol_buf); io.DLr = t.DLb(S.pol_buf);
  io.Lw = t.Lb(pb); io.DLw = t.DLb(pb);
  double dmax = rl_inf();
  if (S.init) {  // x_0 and l_array = copy(u_array) (:228) go to the warp's scratch: warp 0 -> its buffer 1, warp 1 -> policy-free
    // space is not available, so both warps roll out straight from the inputs: the open-loop rollout reads only (x_0, u)
    const double* x0 = P.x0 + (P.x0_count > 1 ? p * n : 0);
    const double* ui = P.u_init + (P.u_count > 1 ? p * (size_t)m * N : 0);
    // stage the inputs in the destination buffers' twin so that coop_rollout's (Xs, Us) view is a plain trajectory
    double* Xs = g == 0 ? t.Xb(0, 1) : t.Xb(1, 0);
    double* Us = g == 0 ? t.Ub(0, 1) : t.Ub(1, 0);
    if (g == 0) {
      phase(lane, [&](int l) {
        for (int i = l; i < n; i += 32) Xs[i] = x0[i];
        for (int e = l; e < N * m; e += 32) Us[e] = ui[e];
      });
      io.Xs = Xs; io.Us = Us;
    } else {  // warp 1 has a single buffer: its rollout reads the inputs in place (global memory, host layout == stage-major)
      io.Xs = x0; io.Us = ui;
      (void)Xs; (void)Us;
    }
    r.st_roll = coop_rollout<D>(lane, P, io, 0.0, true, dmax, w);
  } else {
    double eps = S.eps;  // G = 2: one candidate per round
    io.Xs = t.Xb(S.cur_col, S.cur_buf); io.Us = t.Ub(S.cur_col, S.cur_buf);
    r.st_roll = coop_rollout<D>(lane, P, io, eps, false, dmax, w);
  }
  r.dmax = dmax;
  if (r.st_roll) return r;
  CoopIO bio = io;
  bio.Xs = io.Xd; bio.Us = io.Ud;  // the pass sweeps the candidate just rolled out
  double val = rl_inf();
  if (g == 1) r.rc = coop_backward_pass<D, CT, true>(lane, P, cp, theta, bio, false, r.mu, r.delta, r.restarts, val, w);
  else r.rc = coop_backward_pass<D, CT, false>(lane, P, cp, theta, bio, S.init, r.mu, r.delta, r.restarts, val, w);
  r.nw = val;
  return r;
}

// x_array / l_array / L_array in host layout (instance slowest) + the per-instance results; `tid` of `nthreads`
RL_HD void coop2_write_outputs(const SolveParams& P, const Coop2Traj& t, const SpecState& S, size_t inst, int tid, int nthreads) {
  const int n = t.n, m = t.m, N = t.N;
  if (tid == 0) {
    P.value[inst] = S.status ? rl_inf() : S.value;
    P.status[inst] = S.status;
    P.iters[inst] = S.iters;
    P.trials[inst] = S.trials;
    P.restarts[inst] = S.restarts;
    P.mu_out[inst] = S.mu;
    P.d_out[inst] = S.d_current;
  }
  const double* X = t.Xb(S.cur_col, S.cur_buf);
  const double* U = t.Ub(S.cur_col, S.cur_buf);
  const double* L = t.Lb(S.pol_buf);
  if (P.xo) for (int e = tid; e < (N + 1) * n; e += nthreads) P.xo[inst * (size_t)(N + 1) * n + e] = X[e];
  if (P.lo) for (int e = tid; e < N * m; e += nthreads) P.lo[inst * (size_t)N * m + e] = U[e];
  if (P.Lo) for (int e = tid; e < N * m * n; e += nthreads) P.Lo[inst * (size_t)N * m * n + e] = S.has_pol ? L[e] : 0.0;
}

}  // namespace rl
