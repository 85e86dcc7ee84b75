// rl_coop.cuh -- warp-cooperative iLEQG solve: ONE WARP per instance.
//
// The thread-per-instance kernel (rl_core.cuh) is throughput-optimal for fleets but its latency is that of
// one serial thread, and at n = 12 its per-thread matrices spill to strided local memory.  Small batches
// (a single MPC problem: <= 10 theta; Nelder-Mead: <= 6 candidates) and the quadrotor need intra-instance
// parallelism instead: here the per-stage matrices AND the instance's trajectories live in shared memory and
// the OUTPUT elements of every small dense operation are spread over the 32 lanes.  Each output is still
// accumulated by one lane in the canonical order, so results are bit-identical to the serial formulation.
//
// Code is written as a sequence of `phase`s: on the device every lane runs the phase body once and the warp
// synchronises; the g++ test build (tests/_hostemu) runs the 32 virtual lanes of each phase one after another.
// All control flow is warp-uniform (decisions are taken on values every lane reads from shared memory).
#pragma once
#include "rl_core.cuh"

namespace rl {

template <class F>
RL_HD void phase(int lane, F&& f) {
#if defined(__CUDA_ARCH__)
  f(lane);
  __syncwarp();
#else
  (void)lane;
  for (int l = 0; l < 32; ++l) f(l);
#endif
}

// compile-time switches handed to generic lambdas
struct CoopTagFast { static constexpr bool value = true; };
struct CoopTagLib { static constexpr bool value = false; };

// shared-memory workspace of one warp (doubles)
template <int n, int m>
struct CoopWs {
  double S[n * n], sv[n], M[n * n], invd[n], Z[n * n], z[n], DS[n * n], Dsv[n];
  double A[n * n], B[n * m], Q[n * n], R[m * m], Pm[m * n], qv[n], r[m];
  double T[n * n], U[n * m], g[m], G[m * n], H[m * m], CH[m * m], invh[m], L[m * n], dl[m], HL[m * n], Hdl[m];
  double x[n], u[m], xn[n];
  double q, flag, nrm;  // stage cost value; domain-error flag; ||l - u||^2 of the rollout step
  double sc[8];         // per-stage trigonometry shared by the lanes (CoopDyn<quadrotor>)
  double V[m * n], Hdlg[m];  // H L + G and H dl + g (dense stage)
  // (i, j) of the t-th entry of the upper triangle of an n x n / m x m matrix, column by column (coop_ws_init)
  unsigned char tn_i[n * (n + 1) / 2], tn_j[n * (n + 1) / 2], tm_i[m * (m + 1) / 2], tm_j[m * (m + 1) / 2];
};

// once per kernel: the index tables of the workspace
template <int n, int m>
RL_HD void coop_ws_init(int lane, CoopWs<n, m>& w) {
  phase(lane, [&](int l) {
    if (l != 0) return;
    int t = 0;
    for (int c = 0; c < n; ++c) for (int r = 0; r <= c; ++r) { w.tn_i[t] = (unsigned char)r; w.tn_j[t] = (unsigned char)c; ++t; }
    t = 0;
    for (int c = 0; c < m; ++c) for (int r = 0; r <= c; ++r) { w.tm_i[t] = (unsigned char)r; w.tm_j[t] = (unsigned char)c; ++t; }
  });
}

// per-instance trajectories, contiguous per instance: X[2][(N+1)*n], U[2][N*m], Lg[N*m*n], DL[N*m]
struct CoopTraj {
  double *X, *U, *Lg, *DL;
};

// what one pass / rollout of an instance reads and writes (all contiguous per instance, stage-major):
//   Xs, Us  trajectory swept by a backward pass / the nominal (xbar, l) a rollout starts from
//   Xd, Ud  destination of a rollout
//   Lr, DLr policy read: gains of an evaluating pass; gains and dl of a rollout
//   Lw, DLw policy written by an optimising pass
struct CoopIO {
  const double *Xs, *Us;
  double *Xd, *Ud;
  const double *Lr, *DLr;
  double *Lw, *DLw;
};
// the single-warp kernel's view: trajectory double buffer `buf` / `buf ^ 1`, one policy buffer
RL_HD CoopIO coop_io(const CoopTraj& tj, int buf, int n, int m, int N) {
  CoopIO io;
  io.Xs = tj.X + (size_t)buf * (N + 1) * n; io.Us = tj.U + (size_t)buf * N * m;
  io.Xd = tj.X + (size_t)(buf ^ 1) * (N + 1) * n; io.Ud = tj.U + (size_t)(buf ^ 1) * N * m;
  io.Lr = tj.Lg; io.DLr = tj.DL; io.Lw = tj.Lg; io.DLw = tj.DL;
  return io;
}

// Per-model hooks of the cooperative kernel.  trig(l, w): executed by lane l INSIDE the phase that precedes the model
// evaluation (lanes 8.. are free there) -- values every lane of the next phase shares; f(): the dynamics step, run by
// one lane.  Default: nothing to share.
template <class D> struct CoopDyn {
  template <int n, int m> RL_HD static void trig(int, CoopWs<n, m>&, const double*) {}
  template <int n, int m> RL_HD static bool f(const double* mp, CoopWs<n, m>& w) { return D::f(mp, w.x, w.u, w.xn); }
};
// the quadrotor: sin / cos of the three Euler angles, one sincos per lane (lanes 8..10), instead of twelve calls in the
// lane that steps the dynamics and twenty-four in each of the sixteen lanes that seed a dual-number direction
template <> struct CoopDyn<Dyn<RATILQR_MODEL_QUADROTOR>> {
  template <int n, int m> RL_HD static void trig(int l, CoopWs<n, m>& w, const double* x) {  // x: the state (w.x or its source)
    // one lane per angle, sin and cos through one branch-free range reduction (the same bits as sin() / cos())
    if (l >= 8 && l < 11) { const int a = l - 8; rl_sincos_any(x[3 + a], &w.sc[2 * a], &w.sc[2 * a + 1]); }
  }
  template <int n, int m> RL_HD static bool f(const double* mp, CoopWs<n, m>& w) { quadrotor_body_sc<double>(mp, w.x, w.u, w.xn, w.sc); return true; }
};

// models whose Jacobian comes from duals evaluate one seeded direction per lane
// `extra(l)` runs in the same phase (the stage cost, on a lane the Jacobian leaves free: lane 16)
template <class D> struct CoopJac {
  template <int n, int m, class Extra>
  RL_HD static void run(int lane, const double* mp, CoopWs<n, m>& w, Extra extra) {
    phase(lane, [&](int l) { extra(l); if (l == 1) D::jac(mp, w.x, w.u, w.A, w.B); });
  }
};
template <class Body, int n, int m, class Extra>
RL_HD void coop_dual_jac(int lane, Body body, const double* mp, CoopWs<n, m>& w, Extra extra) {
  static_assert(n + m <= 16, "lane 16 runs the extra job of the phase");
  phase(lane, [&](int l) {
    extra(l);
    if (l >= n + m) return;
    Dual<1> xd[n], ud[m], xo[n];
    for (int i = 0; i < n; ++i) { xd[i].v = w.x[i]; xd[i].d[0] = (i == l) ? 1.0 : 0.0; }
    for (int j = 0; j < m; ++j) { ud[j].v = w.u[j]; ud[j].d[0] = (n + j == l) ? 1.0 : 0.0; }
    body(mp, xd, ud, xo);
    for (int i = 0; i < n; ++i) { if (l < n) w.A[i + l * n] = xo[i].d[0]; else w.B[i + (l - n) * n] = xo[i].d[0]; }
  });
}
template <> struct CoopJac<Dyn<RATILQR_MODEL_CARTPOLE>> {
  template <int n, int m, class Extra> RL_HD static void run(int lane, const double* mp, CoopWs<n, m>& w, Extra extra) { coop_dual_jac(lane, CartpoleBody(), mp, w, extra); }
};
struct QuadrotorBodySc {  // the body with the shared trigonometry of CoopDyn<quadrotor>::trig
  const double* sc;
  template <class T> RL_HD void operator()(const double* p, const T* x, const T* u, T* xn) const { quadrotor_body_sc<T>(p, x, u, xn, sc); }
};
template <> struct CoopJac<Dyn<RATILQR_MODEL_QUADROTOR>> {
  template <int n, int m, class Extra> RL_HD static void run(int lane, const double* mp, CoopWs<n, m>& w, Extra extra) {
    QuadrotorBodySc body; body.sc = w.sc;
    coop_dual_jac(lane, body, mp, w, extra);
  }
};

// The lane's NE dot products of length LEN advance TOGETHER (k outer, fully unrolled): NE independent FMA chains per
// lane and compile-time shared-memory offsets, instead of one rolled loop after the other.  get(q, k, a, b) yields the
// k-th operand pair of output q; acc[q] is either started by the first product (start[q]) or accumulates onto its
// initial value -- per output the same sequence of operations as rl::coldot / rl::coldot_acc / rl::dot_acc.
template <int LEN, int NE, class Get>
RL_HD void lane_dots(Get get, const bool* start, double* acc) {
#pragma unroll
  for (int k = 0; k < LEN; ++k) {
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      double a, b;
      get(q, k, a, b);
      acc[q] = (k == 0 && start[q]) ? a * b : rl_fma(a, b, acc[q]);
    }
  }
}

// One Riccati stage for a DENSE model (quadrotor, cart-pole: no compile-time structure of A, B), theta != 0, fused
// accumulation: the arithmetic of coop_riccati_stage below, output element by output element, restructured for
// instruction-level parallelism -- bit-identical results (tests/test_spec_kernel.py::test_coop_dense_stage...).
template <class Tr, bool OPT, bool HAS_DL>
RL_HD int coop_riccati_stage_dense(int lane, CoopWs<Tr::n, Tr::m>& w, double theta, double mu, const double* RL_RESTRICT Winv,
                                   double detW, double& s, double* detprod) {
  constexpr int n = Tr::n, m = Tr::m;
  double detM = 1.0;
#if defined(__CUDA_ARCH__) && !defined(RL_COOP_CHOL_SMEM)
  // Cholesky of M = inv(W) - theta S+ in REGISTERS: lane i < n holds the upper-triangle entries that touch index i
  // (row[k] = M(min(i,k), max(i,k))), i.e. entry (k, i), k <= i, of the running Schur complement lives in lane i.  Step j:
  // the pivot row's entry (j, k) comes from lane k by shuffles (issued before the rsqrt, so they travel under its
  // latency); every lane scales and applies the rank-1 update to its own entries.  No shared-memory round
  // trip and no warp barrier per pivot; per entry the same operations in the same order as the shared-memory version
  // below (right-looking, updates in increasing j), hence the same bits.
  {
    const unsigned full = 0xffffffffu;
    const int i = lane < n ? lane : 0;
    double row[n], crow[n], invs[n], diag[n];
    auto load = [&]() {
#pragma unroll
      for (int k = 0; k < n; ++k) {
        const int a = i < k ? i : k, b = i < k ? k : i;
        row[k] = rl_fma(-theta, w.S[a + b * n], Winv[a + b * n]);
        crow[k] = 0.0;
        // every lane also tracks the whole running DIAGONAL (the same fma lane k applies to its entry (k, k)): the next
        // pivot is then a local value and no shuffle sits on the pivot -> rsqrt -> scale -> update chain of the factorisation
        diag[k] = rl_fma(-theta, w.S[k + k * n], Winv[k + k * n]);
      }
    };
    load();
    // positive-definiteness and the library rsqrt's slow-path test are recorded and acted on after the loop (a failed
    // pivot only poisons values nobody uses): the factorisation is one basic block.  A pivot that needs the slow path
    // (denormal / Inf) re-runs the loop through the library routine.
    bool bad = false, slow = false;
    auto factor = [&](auto fast_tag) {
      constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const double d = diag[j];
        bad = bad || !(d > 0.0);
        double pk[n];
#pragma unroll
        for (int k = j + 1; k < n; ++k) pk[k] = __shfl_sync(full, row[j], k);  // M(j, k) lives in lane k
        detM = (j == 0) ? d : detM * d;
        const double inv = FAST ? rl_rsqrt_nb(d, slow) : rl_rsqrt(d);
        invs[j] = inv;
        const double ci = row[j] * inv;  // C(i, j) for i > j
        crow[j] = ci;
#pragma unroll
        for (int k = j + 1; k < n; ++k) {
          const double ck = pk[k] * inv;
          diag[k] = rl_fma(-ck, ck, diag[k]);
          if (k <= i) row[k] = rl_fma(-ci, ck, row[k]);
        }
      }
    };
    factor(CoopTagFast());
    if (!bad && slow) {  // (warp-uniform: every lane sees the same pivots)
      load();
      factor(CoopTagLib());
    }
    if (bad) return 1;
    phase(lane, [&](int l) {  // the factor goes where the substitutions below read it
      if (l < n) {
#pragma unroll
        for (int j = 0; j < n; ++j) if (j < l) w.M[l + j * n] = crow[j];
      }
      if (l == 0) {
#pragma unroll
        for (int j = 0; j < n; ++j) w.invd[j] = invs[j];
      }
    });
  }
#else
  phase(lane, [&](int l) {
#pragma unroll
    for (int q = 0; q < (n * n + 31) / 32; ++q) { const int e = l + 32 * q; if (e < n * n) w.M[e] = rl_fma(-theta, w.S[e], Winv[e]); }
  });
  // Cholesky in place, RIGHT-LOOKING: the upper triangle holds the running Schur complement, the lower triangle receives
  // the factor.  Step j: every lane reads the pivot, scales the column entries it needs itself, and applies the rank-1
  // update to its share of the trailing (k, i) entries (j < k <= i) -- one phase per step instead of a dependent chain of j
  // FMAs per entry.  An entry still receives its updates in increasing j: the rounding sequence of the left-looking form.
  {
    int bad = 0;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const double d = w.M[j + j * n];
      if (!(d > 0.0)) { bad = 1; break; }
      detM = (j == 0) ? d : detM * d;
      const double inv = rl_rsqrt(d);
      constexpr int dummy = 0;
      (void)dummy;
      const int r = n - j - 1;                 // order of the trailing block
      const int TOT = r * (r + 1) / 2;         // its upper-triangle entries
      phase(lane, [&](int l) {
        if (l == 0) w.invd[j] = inv;
        if (l < r) w.M[(j + 1 + l) + j * n] = w.M[j + (j + 1 + l) * n] * inv;  // C[i, j], i = j + 1 + l
#pragma unroll
        for (int q = 0; q < (n * (n - 1) / 2 + 31) / 32; ++q) {
          const int t = l + 32 * q;
          if (t < TOT) {
            const int k = j + 1 + w.tn_i[t], i = j + 1 + w.tn_j[t];
            const double ci = w.M[j + i * n] * inv, ck = w.M[j + k * n] * inv;
            w.M[k + i * n] = rl_fma(-ci, ck, w.M[k + i * n]);
          }
        }
      });
    }
    if (bad) return 1;
  }
#endif
  // forward substitutions, column c of [S+ | s_vec+] per lane, kept in registers; right-looking: once Z[i] is known every
  // later row takes its update at once (the updates of a row still arrive in increasing i: same rounding sequence)
  phase(lane, [&](int l) {
    if (l > n) return;
    const int c = l;
    double col[n];
#pragma unroll
    for (int i = 0; i < n; ++i) col[i] = (c < n) ? w.S[i + c * n] : w.sv[i];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const double zi = col[i] * w.invd[i];
      col[i] = zi;
#pragma unroll
      for (int r = i + 1; r < n; ++r) col[r] = rl_fma(-w.M[r + i * n], zi, col[r]);
    }
#pragma unroll
    for (int i = 0; i < n; ++i) { if (c < n) w.Z[i + c * n] = col[i]; else w.z[i] = col[i]; }
  });
  // D S+ = S+ + theta Z'Z (every entry computed directly: products commute, so (i, j) and (j, i) get the same bits) and
  // D s_vec+ = s_vec+ + theta Z'z
  phase(lane, [&](int l) {
    constexpr int TOT = n * n + n, NE = (TOT + 31) / 32;
    const double* xa[NE]; const double* xb[NE]; double acc[NE]; bool start[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q, ee = e < TOT ? e : 0;
      const int i = ee < n * n ? ee % n : ee - n * n, j = ee < n * n ? ee / n : 0;
      xa[q] = w.Z + i * n; xb[q] = ee < n * n ? w.Z + j * n : w.z; acc[q] = 0.0; start[q] = true;
    }
    lane_dots<n, NE>([&](int q, int k, double& a, double& b) { a = xa[q][k]; b = xb[q][k]; }, start, acc);
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q;
      if (e < n * n) w.DS[e] = rl_fma(theta, acc[q], w.S[(e % n) + (e / n) * n]);
      else if (e < TOT) w.Dsv[e - n * n] = rl_fma(theta, acc[q], w.sv[e - n * n]);
    }
  });
  double quad = w.z[0] * w.z[0];
  for (int k = 1; k < n; ++k) quad = rl_fma(w.z[k], w.z[k], quad);
  *detprod *= detW * detM;
  const double extra = (theta / 2) * quad;
  phase(lane, [&](int l) {  // T = (D S+) A, U = (D S+) B
    constexpr int TOT = n * n + n * m, NE = (TOT + 31) / 32;
    const double* xa[NE]; const double* xb[NE]; double acc[NE]; bool start[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q, ee = e < TOT ? e : 0;
      const int f = ee < n * n ? ee : ee - n * n;
      xa[q] = w.DS + f % n; xb[q] = (ee < n * n ? w.A : w.B) + (f / n) * n; acc[q] = 0.0; start[q] = true;
    }
    lane_dots<n, NE>([&](int q, int k, double& a, double& b) { a = xa[q][k * n]; b = xb[q][k]; }, start, acc);
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q;
      if (e < n * n) w.T[e] = acc[q]; else if (e < TOT) w.U[e - n * n] = acc[q];
    }
  });
  phase(lane, [&](int l) {  // g, G, H (upper triangle mirrored)
    constexpr int NH = m * (m + 1) / 2, TOT = m + m * n + NH, NE = (TOT + 31) / 32;
    const double* xa[NE]; const double* xb[NE]; double acc[NE]; bool start[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q, ee = e < TOT ? e : 0;
      if (ee < m) { xa[q] = w.Dsv; xb[q] = w.B + ee * n; acc[q] = w.r[ee]; start[q] = false; }
      else if (ee < m + m * n) {
        const int f = ee - m, i = f % m, j = f / m;
        xa[q] = w.T + j * n; xb[q] = w.B + i * n; start[q] = Tr::p_kind(i, j) == 0; acc[q] = start[q] ? 0.0 : w.Pm[f];
      } else {
        const int t = ee - m - m * n, i = w.tm_i[t], j = w.tm_j[t];
        xa[q] = w.U + j * n; xb[q] = w.B + i * n; start[q] = Tr::r_kind(i, j) == 0; acc[q] = start[q] ? 0.0 : w.R[i + j * m];
      }
    }
    lane_dots<n, NE>([&](int q, int k, double& a, double& b) { a = xa[q][k]; b = xb[q][k]; }, start, acc);
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q;
      if (e < m) w.g[e] = acc[q];
      else if (e < m + m * n) w.G[e - m] = acc[q];
      else if (e < TOT) {
        const int t = e - m - m * n, i = w.tm_i[t], j = w.tm_j[t];
        const double h = (i == j) ? acc[q] + mu : acc[q];
        w.H[i + j * m] = h;
        w.H[j + i * m] = h;
      }
    }
  });
  if (OPT) {
    // Cholesky of H (m x m, tiny): every lane factors it redundantly in registers -- no phases, the same operations
    // (device: the pivots' tests are recorded and acted on after the loop, branch-free rsqrt with the library routine on
    // its slow-path flag -- as in the factorisation of M above)
    double ch[m * m], ih[m];
    {
      bool bad = false, slow = false;
      auto factor = [&](auto fast_tag) {
        constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
        for (int j = 0; j < m; ++j) {
          double d = w.H[j + j * m];
#pragma unroll
          for (int k = 0; k < j; ++k) d = rl_fma(-ch[j + k * m], ch[j + k * m], d);
          bad = bad || !(d > 0.0);
          const double inv = FAST ? rl_rsqrt_nb(d, slow) : rl_rsqrt(d);
          ih[j] = inv;
#pragma unroll
          for (int i = j + 1; i < m; ++i) {
            double a = w.H[j + i * m];
#pragma unroll
            for (int k = 0; k < j; ++k) a = rl_fma(-ch[i + k * m], ch[j + k * m], a);
            ch[i + j * m] = a * inv;
          }
        }
      };
#if defined(__CUDA_ARCH__)
      factor(CoopTagFast());
      if (!bad && slow) factor(CoopTagLib());
#else
      factor(CoopTagLib());
#endif
      if (bad) return 2;
    }
    phase(lane, [&](int l) {  // L = -H\G (column c per lane), dl = -H\g (column n)
      if (l > n) return;
      const int c = l;
      double y[m];
#pragma unroll
      for (int i = 0; i < m; ++i) {
        double a = (c < n) ? w.G[i + c * m] : w.g[i];
#pragma unroll
        for (int k = 0; k < i; ++k) a = rl_fma(-ch[i + k * m], y[k], a);
        y[i] = a * ih[i];
      }
#pragma unroll
      for (int i = m - 1; i >= 0; --i) {
        double a = y[i];
#pragma unroll
        for (int k = i + 1; k < m; ++k) a = rl_fma(-ch[k + i * m], y[k], a);
        y[i] = a * ih[i];
      }
#pragma unroll
      for (int i = 0; i < m; ++i) { if (c < n) w.L[i + c * m] = -y[i]; else w.dl[i] = -y[i]; }
    });
  }
  phase(lane, [&](int l) {  // V = H L + G, H dl, H dl + g
    constexpr int TOT = m * n + (HAS_DL ? m : 0), NE = (TOT + 31) / 32;
    const double* xa[NE]; const double* xb[NE]; double acc[NE]; bool start[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q, ee = e < TOT ? e : 0;
      const int i = ee < m * n ? ee % m : ee - m * n;
      xa[q] = w.H + i; xb[q] = ee < m * n ? w.L + (ee / m) * m : w.dl;
      start[q] = !(ee < m * n); acc[q] = start[q] ? 0.0 : w.G[ee];  // V accumulates onto G (rl::riccati_stage, fused order)
    }
    lane_dots<m, NE>([&](int q, int k, double& a, double& b) { a = xa[q][k * m]; b = xb[q][k]; }, start, acc);
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q;
      if (e < m * n) w.V[e] = acc[q];
      else if (e < TOT) { w.Hdl[e - m * n] = acc[q]; w.Hdlg[e - m * n] = acc[q] + w.g[e - m * n]; }
    }
  });
  double sval = w.q + s;
  if (HAS_DL) {
    double a = w.dl[0] * w.Hdl[0];
    for (int k = 1; k < m; ++k) a = rl_fma(w.dl[k], w.Hdl[k], a);
    sval = dot_acc<m>(rl_fma(0.5, a, sval), w.dl, 1, w.g, 1);
  }
  s = sval + extra;
  phase(lane, [&](int l) {  // s_vec and S (upper triangle, mirrored)
    constexpr int NS = n * (n + 1) / 2, TOT = n + NS, NE = (TOT + 31) / 32;
    const double* xa[NE]; const double* xb[NE]; double acc[NE]; bool start[NE]; int oi[NE], oj[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q, ee = e < TOT ? e : 0;
      if (ee < n) { oi[q] = ee; oj[q] = -1; xa[q] = w.Dsv; xb[q] = w.A + ee * n; acc[q] = w.qv[ee]; start[q] = false; }
      else {
        const int t = ee - n, i = w.tn_i[t], j = w.tn_j[t];
        oi[q] = i; oj[q] = j;
        xa[q] = w.T + j * n; xb[q] = w.A + i * n; start[q] = Tr::q_kind(i, j) == 0; acc[q] = start[q] ? 0.0 : w.Q[i + j * n];
      }
    }
    lane_dots<n, NE>([&](int q, int k, double& a, double& b) { a = xa[q][k]; b = xb[q][k]; }, start, acc);
    bool cont[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) cont[q] = false;
    // L'(H dl + g) resp. L'g for s_vec ; L'(H L + G) for S
    lane_dots<m, NE>([&](int q, int k, double& a, double& b) {
      a = w.L[k + oi[q] * m];
      b = oj[q] < 0 ? (HAS_DL ? w.Hdlg[k] : w.g[k]) : w.V[k + oj[q] * m];
    }, cont, acc);
    // G'dl for s_vec (only with dl) ; G'L for S
    lane_dots<m, NE>([&](int q, int k, double& a, double& b) {
      a = w.G[k + oi[q] * m];
      b = oj[q] < 0 ? (HAS_DL ? w.dl[k] : 0.0) : w.L[k + oj[q] * m];
      if (oj[q] < 0 && !HAS_DL) a = 0.0;  // fma(0, 0, acc) == acc: the evaluating pass has no G'dl term
    }, cont, acc);
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = l + 32 * q;
      // straight into S / s_vec: nothing in this phase reads them any more (S+ and s_vec+ were last used for D S+, D s_vec+)
      if (e < n) w.sv[e] = acc[q];
      else if (e < TOT) { w.S[oi[q] + oj[q] * n] = acc[q]; w.S[oj[q] + oi[q] * n] = acc[q]; }
    }
  });
  return 0;
}

#ifndef RL_COOP_DENSE
#define RL_COOP_DENSE 1
#endif

// One Riccati stage, cooperative.  w.S / w.sv / s hold (S+, s_vec+, s+) on entry, the stage's values on exit.
// returns 0 / 1 (M not PD) / 2 (H not PD); identical arithmetic per output element to rl::riccati_stage.
template <class Tr, bool OPT, bool HAS_DL>
RL_HD int coop_riccati_stage(int lane, CoopWs<Tr::n, Tr::m>& w, double theta, double mu, const double* RL_RESTRICT W,
                             const double* RL_RESTRICT Winv, double detW, double& s, double* detprod = nullptr) {
  constexpr int n = Tr::n, m = Tr::m;
  if constexpr (RL_COOP_DENSE && RL_FUSED && !Tr::structured) {
    if (theta != 0.0 && detprod) return coop_riccati_stage_dense<Tr, OPT, HAS_DL>(lane, w, theta, mu, Winv, detW, s, detprod);
  }
  double extra = 0.0;
  if (theta == 0.0) {
    phase(lane, [&](int l) {
      for (int e = l; e < n * n; e += 32) w.DS[e] = w.S[e];
      for (int e = l; e < n; e += 32) w.Dsv[e] = w.sv[e];
    });
    double tr = 0.0;  // every lane redundantly: 1/2 tr(W S+)
    for (int i = 0; i < n; ++i) {
      double t = W[i] * w.S[i * n];
      for (int k = 1; k < n; ++k) t = rl_fma(W[i + k * n], w.S[k + i * n], t);
      tr = (i == 0) ? t : tr + t;
    }
    extra = 0.5 * tr;
  } else {
    phase(lane, [&](int l) { for (int e = l; e < n * n; e += 32) w.M[e] = RL_FUSED ? rl_fma(-theta, w.S[e], Winv[e]) : Winv[e] - theta * w.S[e]; });
    // Cholesky in place, one column per step: lane i owns row i; the pivot is recomputed by every lane
    double detM = 1.0;
    int bad = 0;
    for (int j = 0; j < n; ++j) {
      double d = w.M[j + j * n];
      for (int k = 0; k < j; ++k) d = rl_fma(-w.M[j + k * n], w.M[j + k * n], d);
      if (!(d > 0.0)) { bad = 1; break; }
      detM = (j == 0) ? d : detM * d;
      const double inv = rl_rsqrt(d);
      phase(lane, [&](int l) {
        if (l == 0) w.invd[j] = inv;
        for (int i = j + 1 + l; i < n; i += 32) {
          double a = w.M[j + i * n];
          for (int k = 0; k < j; ++k) a = rl_fma(-w.M[i + k * n], w.M[j + k * n], a);
          w.M[i + j * n] = a * inv;
        }
      });
    }
    if (bad) return 1;
    // forward substitutions: lane c owns column c of Z = C^-1 S+, one more lane owns z = C^-1 s_vec+
    phase(lane, [&](int l) {
      for (int c = l; c <= n; c += 32) {
        for (int i = 0; i < n; ++i) {
          double a = (c < n) ? w.S[i + c * n] : w.sv[i];
          for (int k = 0; k < i; ++k) a = rl_fma(-w.M[i + k * n], (c < n) ? w.Z[k + c * n] : w.z[k], a);
          if (c < n) w.Z[i + c * n] = a * w.invd[i]; else w.z[i] = a * w.invd[i];
        }
      }
    });
    phase(lane, [&](int l) {
      for (int e = l; e < n * n; e += 32) {  // D S+ = S+ + theta Z'Z, upper triangle mirrored
        int i = e % n, j = e / n;
        if (j < i) continue;
        double v = w.Z[i * n] * w.Z[j * n];
        for (int k = 1; k < n; ++k) v = rl_fma(w.Z[k + i * n], w.Z[k + j * n], v);
        v = rl_fma(theta, v, w.S[i + j * n]);
        w.DS[i + j * n] = v;
        w.DS[j + i * n] = v;
      }
      for (int i = l; i < n; i += 32) {
        double v = w.Z[i * n] * w.z[0];
        for (int k = 1; k < n; ++k) v = rl_fma(w.Z[k + i * n], w.z[k], v);
        w.Dsv[i] = rl_fma(theta, v, w.sv[i]);
      }
    });
    double quad = w.z[0] * w.z[0];
    for (int k = 1; k < n; ++k) quad = rl_fma(w.z[k], w.z[k], quad);
    if (detprod) { *detprod *= detW * detM; extra = (theta / 2) * quad; }
    else extra = (theta / 2) * quad - (1 / (2 * theta)) * log(detW * detM);
  }
  phase(lane, [&](int l) {  // T = (D S+) A, U = (D S+) B
    for (int e = l; e < n * n + n * m; e += 32) {
      if (e < n * n) { int i = e % n, j = e / n; w.T[e] = coldot<Tr, KindA, n>(w.A, j, w.DS + i, n); }
      else { int f = e - n * n, i = f % n, j = f / n; w.U[f] = coldot<Tr, KindB, n>(w.B, j, w.DS + i, n); }
    }
  });
  phase(lane, [&](int l) {  // g, G, H (upper mirrored)
    for (int e = l; e < m + m * n + m * m; e += 32) {
      if (e < m) w.g[e] = RL_FUSED ? coldot_acc<Tr, KindB, n>(w.r[e], w.B, e, w.Dsv, 1) : w.r[e] + coldot<Tr, KindB, n>(w.B, e, w.Dsv, 1);
      else if (e < m + m * n) {
        int f = e - m, i = f % m, j = f / m;
        if (RL_FUSED && Tr::p_kind(i, j) != 0) { w.G[f] = coldot_acc<Tr, KindB, n>(w.Pm[f], w.B, i, w.T + j * n, 1); continue; }
        double a = coldot<Tr, KindB, n>(w.B, i, w.T + j * n, 1);
        w.G[f] = (Tr::p_kind(i, j) == 0) ? a : w.Pm[f] + a;
      } else {
        int f = e - m - m * n, i = f % m, j = f / m;
        if (j < i) continue;
        double h;
        if (RL_FUSED && Tr::r_kind(i, j) != 0) h = coldot_acc<Tr, KindB, n>(w.R[i + j * m], w.B, i, w.U + j * n, 1);
        else { double a = coldot<Tr, KindB, n>(w.B, i, w.U + j * n, 1); h = (Tr::r_kind(i, j) == 0) ? a : w.R[i + j * m] + a; }
        if (i == j) h = h + mu;
        w.H[i + j * m] = h;
        w.H[j + i * m] = h;
      }
    }
  });
  if (OPT) {
    int bad = 0;
    for (int j = 0; j < m; ++j) {  // Cholesky of H, same column scheme
      double d = w.H[j + j * m];
      for (int k = 0; k < j; ++k) d = rl_fma(-w.CH[j + k * m], w.CH[j + k * m], d);
      if (!(d > 0.0)) { bad = 1; break; }
      const double inv = rl_rsqrt(d);
      phase(lane, [&](int l) {
        if (l == 0) w.invh[j] = inv;
        for (int i = j + 1 + l; i < m; i += 32) {
          double a = w.H[j + i * m];
          for (int k = 0; k < j; ++k) a = rl_fma(-w.CH[i + k * m], w.CH[j + k * m], a);
          w.CH[i + j * m] = a * inv;
        }
      });
    }
    if (bad) return 2;
    phase(lane, [&](int l) {  // L = -H\G (column c per lane), dl = -H\g (column n)
      for (int c = l; c <= n; c += 32) {
        double y[m];
        for (int i = 0; i < m; ++i) {
          double a = (c < n) ? w.G[i + c * m] : w.g[i];
          for (int k = 0; k < i; ++k) a = rl_fma(-w.CH[i + k * m], y[k], a);
          y[i] = a * w.invh[i];
        }
        for (int i = m - 1; i >= 0; --i) {
          double a = y[i];
          for (int k = i + 1; k < m; ++k) a = rl_fma(-w.CH[k + i * m], y[k], a);
          y[i] = a * w.invh[i];
        }
        for (int i = 0; i < m; ++i) { if (c < n) w.L[i + c * m] = -y[i]; else w.dl[i] = -y[i]; }
      }
    });
  }
  phase(lane, [&](int l) {  // H L, H dl
    for (int e = l; e < m * n + (HAS_DL ? m : 0); e += 32) {
      if (e < m * n) {
        int i = e % m, j = e / m;
        if (RL_FUSED) { w.HL[e] = dot_acc<m>(w.G[e], w.H + i, m, w.L + j * m, 1); continue; }  // V = H L + G (fused order)
        double a = w.H[i] * w.L[j * m];
        for (int k = 1; k < m; ++k) a = rl_fma(w.H[i + k * m], w.L[k + j * m], a);
        w.HL[e] = a;
      } else {
        int i = e - m * n;
        double a = w.H[i] * w.dl[0];
        for (int k = 1; k < m; ++k) a = rl_fma(w.H[i + k * m], w.dl[k], a);
        w.Hdl[i] = a;
      }
    }
  });
  double sval = w.q + s;
  if (HAS_DL) {
    double a = w.dl[0] * w.Hdl[0]; for (int k = 1; k < m; ++k) a = rl_fma(w.dl[k], w.Hdl[k], a);
    if (RL_FUSED) sval = dot_acc<m>(rl_fma(0.5, a, sval), w.dl, 1, w.g, 1);
    else { double b = w.dl[0] * w.g[0]; for (int k = 1; k < m; ++k) b = rl_fma(w.dl[k], w.g[k], b); sval = (sval + 0.5 * a) + b; }
  }
  s = sval + extra;
  phase(lane, [&](int l) {  // s_vec and S (upper); nothing in this phase reads S+ / s_vec+ any more
    for (int e = l; e < n + n * n; e += 32) {
      if (e < n) {
        int i = e;
        if (RL_FUSED) {
          double acc = coldot_acc<Tr, KindA, n>(w.qv[i], w.A, i, w.Dsv, 1);
          if (HAS_DL) { double t[m]; for (int k = 0; k < m; ++k) t[k] = w.Hdl[k] + w.g[k]; acc = dot_acc<m>(acc, w.L + i * m, 1, t, 1); }
          else acc = dot_acc<m>(acc, w.L + i * m, 1, w.g, 1);
          if (HAS_DL) acc = dot_acc<m>(acc, w.G + i * m, 1, w.dl, 1);
          w.sv[i] = acc;
          continue;
        }
        double acc = w.qv[i] + coldot<Tr, KindA, n>(w.A, i, w.Dsv, 1);
        if (HAS_DL) { double b = w.L[i * m] * w.Hdl[0]; for (int k = 1; k < m; ++k) b = rl_fma(w.L[k + i * m], w.Hdl[k], b); acc = acc + b; }
        double c = w.L[i * m] * w.g[0]; for (int k = 1; k < m; ++k) c = rl_fma(w.L[k + i * m], w.g[k], c);
        acc = acc + c;
        if (HAS_DL) { double d = w.G[i * m] * w.dl[0]; for (int k = 1; k < m; ++k) d = rl_fma(w.G[k + i * m], w.dl[k], d); acc = acc + d; }
        w.sv[i] = acc;
      } else {
        int f = e - n, i = f % n, j = f / n;
        if (j < i) continue;
        if (RL_FUSED) {
          double acc = (Tr::q_kind(i, j) == 0) ? coldot<Tr, KindA, n>(w.A, i, w.T + j * n, 1) : coldot_acc<Tr, KindA, n>(w.Q[i + j * n], w.A, i, w.T + j * n, 1);
          acc = dot_acc<m>(acc, w.L + i * m, 1, w.HL + j * m, 1);  // L'(H L + G): w.HL holds V in the fused order
          acc = dot_acc<m>(acc, w.G + i * m, 1, w.L + j * m, 1);
          w.S[i + j * n] = acc;
          w.S[j + i * n] = acc;
          continue;
        }
        double a = coldot<Tr, KindA, n>(w.A, i, w.T + j * n, 1);
        double acc = (Tr::q_kind(i, j) == 0) ? a : w.Q[i + j * n] + a;
        double b = w.L[i * m] * w.HL[j * m]; for (int k = 1; k < m; ++k) b = rl_fma(w.L[k + i * m], w.HL[k + j * m], b);
        acc = acc + b;
        double c = w.L[i * m] * w.G[j * m]; for (int k = 1; k < m; ++k) c = rl_fma(w.L[k + i * m], w.G[k + j * m], c);
        acc = acc + c;
        double d = w.G[i * m] * w.L[j * m]; for (int k = 1; k < m; ++k) d = rl_fma(w.G[k + i * m], w.L[k + j * m], d);
        acc = acc + d;
        w.S[i + j * n] = acc;
        w.S[j + i * n] = acc;
      }
    }
  });
  return 0;
}

// backward pass (approximate_model fused), cooperative; same contract as rl::backward_pass
template <class D, class CT, bool OPT>
RL_HD int coop_backward_pass(int lane, const SolveParams& P, const double* cp, double theta, const CoopIO& io, bool zeroL,
                             double& mu, double& delta, int& restarts, double& value, CoopWs<D::n, D::m>& w) {
  constexpr int n = D::n, m = D::m;
  using Tr = StageTraits<D, CT>;
  const int N = P.N;
  const double* Xb = io.Xs;
  const double* Ub = io.Us;
  while (true) {
    double s = 0.0;
    phase(lane, [&](int l) {
      if (l != 0) return;
      for (int i = 0; i < n; ++i) w.x[i] = Xb[(size_t)N * n + i];
      double q;
      w.flag = CT::terminal(cp, w.x, true, q, w.sv, w.Q) ? 0.0 : 1.0;  // :352-354
      w.q = q;
    });
    if (w.flag != 0.0) return RATILQR_ST_DOMAIN;
    s = w.q;
    phase(lane, [&](int l) {
      for (int e = l; e < n * n; e += 32) {
        int i = e % n, j = e / n;
        int a = i < j ? i : j, b2 = i < j ? j : i;  // Symmetric(): upper triangle mirrored
        w.S[e] = (Tr::q_kind(a, b2) == 0) ? 0.0 : w.Q[a + b2 * n];
      }
    });
    bool restart = false;
    double detprod = 1.0, logacc = 0.0;
    for (int k = N - 1; k >= 0; --k) {
      const double* Lrk = io.Lr + (size_t)k * m * n;
      double* Lwk = io.Lw + (size_t)k * m * n;
      // two phases: [operands into the workspace | the model's trigonometry straight from the trajectory] and
      // [Jacobian (lanes 0 .. n+m-1, or lane 1) | stage cost (lane 16)] -- pure functions, so running the Jacobian beside a
      // cost evaluation that ends in a domain error is harmless
      phase(lane, [&](int l) {
        CoopDyn<D>::trig(l, w, Xb + (size_t)k * n);
        for (int e = l; e < n + m + m * n; e += 32) {
          if (e < n) w.x[e] = Xb[(size_t)k * n + e];
          else if (e < n + m) w.u[e - n] = Ub[(size_t)k * m + (e - n)];
          else if (!OPT) w.L[e - n - m] = zeroL ? 0.0 : Lrk[e - n - m];
        }
      });
      CoopJac<D>::run(lane, P.mp, w, [&](int l) {
        if (l != 16) return;
        double q;
        w.flag = CT::stage(cp, k, w.x, w.u, true, q, w.qv, w.Q, w.r, w.R, w.Pm) ? 0.0 : 1.0;
        w.q = q;
      });
      if (w.flag != 0.0) return RATILQR_ST_DOMAIN;
      const size_t wo = P.W_tv ? (size_t)k * n * n : 0;
      int rc = coop_riccati_stage<Tr, OPT, OPT>(lane, w, theta, mu, P.W + wo, P.Winv + wo, P.detW[P.W_tv ? k : 0], s, RL_FUSED ? &detprod : nullptr);
      if (rc == 1) return OPT ? RATILQR_ST_M_NOT_PD_OPT : RATILQR_ST_M_NOT_PD_INIT;
      if (RL_FUSED && !(detprod > 1e-250 && detprod < 1e250)) { logacc += log(detprod); detprod = 1.0; }
      if (OPT) {
        if (rc == 2) {
          delta = fmax(P.delta_0, delta * P.delta_0);
          mu = fmax(P.mu_min, mu * delta);
          restarts++;
          if (!(mu < 1e300)) return RATILQR_ST_MU_OVERFLOW;
          restart = true;
          break;
        }
        phase(lane, [&](int l) {
          for (int e = l; e < m * n + m; e += 32) {
            if (e < m * n) Lwk[e] = w.L[e]; else io.DLw[(size_t)k * m + (e - m * n)] = w.dl[e - m * n];
          }
        });
      }
    }
    if (!restart) {
      if (RL_FUSED && theta != 0.0) s = s - (1 / (2 * theta)) * (logacc + log(detprod));
      value = s;
      return 0;
    }
  }
}

// closed-loop / open-loop (init) rollout into buffer cur^1, cooperative
template <class D>
RL_HD int coop_rollout(int lane, const SolveParams& P, const CoopIO& io, double eps, bool init, double& dmax,
                       CoopWs<D::n, D::m>& w) {
  constexpr int n = D::n, m = D::m;
  const int N = P.N;
  const double* Xc = io.Xs;
  const double* Uc = io.Us;
  double* Xn = io.Xd;
  double* Un = io.Ud;
  phase(lane, [&](int l) { for (int i = l; i < n; i += 32) { w.x[i] = Xc[i]; Xn[i] = Xc[i]; } });
  double best = -rl_inf();
  bool has_nan = false;
  for (int k = 0; k < N; ++k) {
    const double* Lk = io.Lr + (size_t)k * m * n;
    phase(lane, [&](int l) {
      CoopDyn<D>::trig(l, w, w.x);
      for (int j = l; j < m; j += 32) {
        const double lj = Uc[(size_t)k * m + j];
        double uj = lj;
        if (!init) {
          if (RL_FUSED) {
            uj = lj + eps * io.DLr[(size_t)k * m + j];
            for (int i = 0; i < n; ++i) uj = rl_fma(Lk[j + i * m], w.x[i] - Xc[(size_t)k * n + i], uj);
          } else {
            double a = Lk[j] * (w.x[0] - Xc[(size_t)k * n]);
            for (int i = 1; i < n; ++i) a = rl_fma(Lk[j + i * m], w.x[i] - Xc[(size_t)k * n + i], a);
            uj = (lj + eps * io.DLr[(size_t)k * m + j]) + a;
          }
        }
        w.u[j] = uj;
        w.g[j] = lj - uj;  // scratch: l - u for the norm
        Un[(size_t)k * m + j] = uj;
      }
    });
    phase(lane, [&](int l) {
      if (l != 0) return;
      double acc = w.g[0] * w.g[0];
      for (int j = 1; j < m; ++j) acc = rl_fma(w.g[j], w.g[j], acc);
      w.nrm = sqrt(acc);
      w.flag = CoopDyn<D>::f(P.mp, w) ? 0.0 : 1.0;
    });
    const double nr = w.nrm;
    if (nr != nr) has_nan = true;
    if (nr > best) best = nr;
    if (w.flag != 0.0) return RATILQR_ST_DOMAIN;
    phase(lane, [&](int l) { for (int i = l; i < n; i += 32) { w.x[i] = w.xn[i]; Xn[(size_t)(k + 1) * n + i] = w.xn[i]; } });
  }
  dmax = has_nan ? (double)NAN : best;
  return 0;
}

// the solve state machine of rl::solve_instance, executed by a whole warp for instance `inst`
template <class D, class CT>
RL_HD bool coop_solve_instance(int lane, const SolveParams& P, size_t inst, CoopWs<D::n, D::m>& w, const CoopTraj& tj, int& cur_out) {
  constexpr int n = D::n, m = D::m;
  const int N = P.N;
  const size_t p = inst / (size_t)P.K;
  if (P.active && !P.active[p]) return false;
  const double* cp = P.cost_params + (P.cp_count > 1 ? p * (size_t)P.ncp : 0);
  const double theta = P.theta[inst];
  int cur = 1, iters = 0, trials = 0, restarts = 0, status = 0, count = 0;
  double mu = 0.0, delta = P.delta_0, d_current = rl_inf(), value = rl_inf();
  double eps_init = P.eps_init, eps = 0.0;
  bool init = true, need_opt = false;
  coop_ws_init(lane, w);
  phase(lane, [&](int l) { for (int e = l; e < N * m * n; e += 32) tj.Lg[e] = 0.0; });  // initialize!: L = 0 (:230-232)
  {
    const double* x0 = P.x0 + (P.x0_count > 1 ? p * n : 0);
    const double* ui = P.u_init + (P.u_count > 1 ? p * (size_t)m * N : 0);
    double* Xc = tj.X + (size_t)cur * (N + 1) * n;
    double* Uc = tj.U + (size_t)cur * N * m;
    phase(lane, [&](int l) {
      for (int i = l; i < n; i += 32) Xc[i] = x0[i];
      for (int e = l; e < N * m; e += 32) Uc[e] = ui[e];
    });
  }
  while (true) {
    if (need_opt) {
      double dummy;
      status = coop_backward_pass<D, CT, true>(lane, P, cp, theta, coop_io(tj, cur, n, m, N), false, mu, delta, restarts, dummy, w);
      if (status) break;
      need_opt = false;
    }
    if (!init) {
      count++;
      if (eps == 0.0 || count > 4000) { status = RATILQR_ST_LINESEARCH_HANG; break; }
    }
    double dmax, nw;
    status = coop_rollout<D>(lane, P, coop_io(tj, cur, n, m, N), eps, init, dmax, w);
    if (status) break;
    int rc = coop_backward_pass<D, CT, false>(lane, P, cp, theta, coop_io(tj, cur ^ 1, n, m, N), init, mu, delta, restarts, nw, w);
    if (rc == RATILQR_ST_DOMAIN) { status = rc; break; }
    if (init) {
      if (rc) { status = RATILQR_ST_M_NOT_PD_INIT; break; }
      value = nw; cur ^= 1; init = false;
      iters++; need_opt = true; eps = eps_init; count = 0;
      continue;
    }
    if (rc) { eps *= P.lambda; continue; }
    if (P.eps_hist && trials < P.eps_hist_cap && lane == 0) {
      double* h = P.eps_hist + (inst * P.eps_hist_cap + trials) * 2;
      h[0] = eps; h[1] = nw - value;
    }
    trials++;
    bool accepted = isapprox_default(nw, value) || nw < value;
    if (!accepted) {
      eps *= P.lambda;
      if (eps < P.eps_min) accepted = true;
    }
    if (!accepted) continue;
    d_current = dmax; value = nw; cur ^= 1;
    if (P.eps_auto) {
      if (count == 1) eps_init = fmin(P.eps_init, eps / P.lambda);
      else { while (eps < P.eps_min) eps = eps / P.lambda; eps_init = eps; }
    }
    if (P.d > d_current && mu <= P.mu_min) break;
    if (iters == P.iter_max) break;
    iters++; need_opt = true; eps = eps_init; count = 0;
  }
  if (status) value = rl_inf();
  if (lane == 0) {
    P.value[inst] = value;
    P.status[inst] = status;
    P.iters[inst] = iters;
    P.trials[inst] = trials;
    P.restarts[inst] = restarts;
    P.mu_out[inst] = mu;
    P.d_out[inst] = d_current;
  }
  cur_out = cur;
  return true;
}

// final x_array, l_array, L_array in the host layout of the C ABI (instance slowest): plain contiguous copies
template <int n, int m>
RL_HD void coop_write_outputs(int lane, const SolveParams& P, size_t inst, const CoopTraj& tj, int cur) {
  const int N = P.N;
  phase(lane, [&](int l) {
    if (P.xo) for (int e = l; e < (N + 1) * n; e += 32) P.xo[inst * (size_t)(N + 1) * n + e] = tj.X[(size_t)cur * (N + 1) * n + e];
    if (P.lo) for (int e = l; e < N * m; e += 32) P.lo[inst * (size_t)N * m + e] = tj.U[(size_t)cur * N * m + e];
    if (P.Lo) for (int e = l; e < N * m * n; e += 32) P.Lo[inst * (size_t)N * m * n + e] = tj.Lg[e];
  });
}

}  // namespace rl
