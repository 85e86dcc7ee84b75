// rl_components.cuh -- per-instance bodies of the component entry points (unit-parity API):
// rollouts, integrate_cost, approximate_model, the two Riccati passes on caller-supplied
// approximations, Monte Carlo rollouts and PETS rollouts.  Arrays here are in the HOST layout of
// the C ABI (column-major, instance slowest); one thread handles one instance / sample.
#pragma once
#include "rl_core.cuh"

namespace rl {

// simulate_dynamics, open loop (ileqg.jl:18-38)
template <class D>
RL_HD int comp_rollout_open(const double* mp, int N, const double* x0, const double* u, double* x) {
  constexpr int n = D::n, m = D::m;
  double xc[n], xn[n], uc[m];
  for (int i = 0; i < n; ++i) { xc[i] = x0[i]; x[i] = x0[i]; }
  for (int k = 0; k < N; ++k) {
    for (int j = 0; j < m; ++j) uc[j] = u[(size_t)k * m + j];
    if (!D::f(mp, xc, uc, xn)) return RATILQR_ST_DOMAIN;
    for (int i = 0; i < n; ++i) { xc[i] = xn[i]; x[(size_t)(k + 1) * n + i] = xn[i]; }
  }
  return 0;
}

// simulate_dynamics, closed loop (ileqg.jl:62-87); optional additive noise w (n*N) (:94-109);
// optionally accumulates integrate_cost (:115-124) on the fly (cost != nullptr)
template <class D, class CT>
RL_HD int comp_rollout_closed(const double* mp, const double* cp, int N, const double* xbar, const double* l,
                              const double* L, const double* w, double* x_new, double* u_new, double* cost) {
  constexpr int n = D::n, m = D::m;
  double x[n], xn[n], u[m];
  for (int i = 0; i < n; ++i) { x[i] = xbar[i]; if (x_new) x_new[i] = x[i]; }
  double J = 0.0;
  for (int k = 0; k < N; ++k) {
    double dx[n];
    for (int i = 0; i < n; ++i) dx[i] = x[i] - xbar[(size_t)k * n + i];
    const double* Lk = L + (size_t)k * m * n;
    for (int j = 0; j < m; ++j) {
      if (RL_FUSED) {  // same accumulation as rl::rollout_candidate
        double uj = l[(size_t)k * m + j];
        for (int i = 0; i < n; ++i) uj = rl_fma(Lk[j + i * m], dx[i], uj);
        u[j] = uj;
      } else {
        double a = Lk[j] * dx[0];
        for (int i = 1; i < n; ++i) a = rl_fma(Lk[j + i * m], dx[i], a);
        u[j] = l[(size_t)k * m + j] + a;
      }
      if (u_new) u_new[(size_t)k * m + j] = u[j];
    }
    if (cost) {
      double q;
      if (!CT::stage(cp, k, x, u, false, q, nullptr, nullptr, nullptr, nullptr, nullptr)) return RATILQR_ST_DOMAIN;
      J += q;
    }
    if (!D::f(mp, x, u, xn)) return RATILQR_ST_DOMAIN;
    for (int i = 0; i < n; ++i) {
      x[i] = w ? xn[i] + w[(size_t)k * n + i] : xn[i];
      if (x_new) x_new[(size_t)(k + 1) * n + i] = x[i];
    }
  }
  if (cost) {
    double q;
    if (!CT::terminal(cp, x, false, q, nullptr, nullptr)) return RATILQR_ST_DOMAIN;
    *cost = J + q;
  }
  return 0;
}

// integrate_cost (ileqg.jl:115-124)
template <class D, class CT>
RL_HD int comp_integrate_cost(const double* cp, int N, const double* x, const double* u, double* cost) {
  constexpr int n = D::n, m = D::m;
  double J = 0.0;
  for (int k = 0; k < N; ++k) {
    double xc[n], uc[m], q;
    for (int i = 0; i < n; ++i) xc[i] = x[(size_t)k * n + i];
    for (int j = 0; j < m; ++j) uc[j] = u[(size_t)k * m + j];
    if (!CT::stage(cp, k, xc, uc, false, q, nullptr, nullptr, nullptr, nullptr, nullptr)) return RATILQR_ST_DOMAIN;
    J += q;
  }
  double xc[n], q;
  for (int i = 0; i < n; ++i) xc[i] = x[(size_t)N * n + i];
  if (!CT::terminal(cp, xc, false, q, nullptr, nullptr)) return RATILQR_ST_DOMAIN;
  *cost = J + q;
  return 0;
}

// approximate_model (ileqg.jl:258-322), one stage k (k == N: terminal)
template <class D, class CT>
RL_HD int comp_linearize_stage(const double* mp, const double* cp, int N, int k, const double* x, const double* u,
                               double* q, double* qv, double* Q, double* r, double* R, double* Pm, double* A,
                               double* Bm) {
  constexpr int n = D::n, m = D::m;
  double xc[n], uc[m];
  for (int i = 0; i < n; ++i) xc[i] = x[(size_t)k * n + i];
  if (k == N) {
    double qq, qvl[n], Ql[n * n];
    for (int i = 0; i < n * n; ++i) Ql[i] = 0.0;
    if (!CT::terminal(cp, xc, true, qq, qvl, Ql)) return RATILQR_ST_DOMAIN;
    q[N] = qq;
    for (int i = 0; i < n; ++i) qv[(size_t)N * n + i] = qvl[i];
    for (int i = 0; i < n; ++i) for (int j = i; j < n; ++j) {  // Symmetric(): mirror the upper triangle (:273)
      Q[(size_t)N * n * n + i + j * n] = Ql[i + j * n];
      Q[(size_t)N * n * n + j + i * n] = Ql[i + j * n];
    }
    return 0;
  }
  for (int j = 0; j < m; ++j) uc[j] = u[(size_t)k * m + j];
  double qq, qvl[n], Ql[n * n], rl_[m], Rl[m * m], Pl[m * n], Al[n * n], Bl[n * m];
  for (int i = 0; i < n * n; ++i) Ql[i] = 0.0;  // structured costs only write their non-zero entries
  for (int i = 0; i < m * m; ++i) Rl[i] = 0.0;
  for (int i = 0; i < m * n; ++i) Pl[i] = 0.0;
  if (!CT::stage(cp, k, xc, uc, true, qq, qvl, Ql, rl_, Rl, Pl)) return RATILQR_ST_DOMAIN;
  D::jac(mp, xc, uc, Al, Bl);
  q[k] = qq;
  for (int i = 0; i < n; ++i) qv[(size_t)k * n + i] = qvl[i];
  for (int i = 0; i < n; ++i) for (int j = i; j < n; ++j) {
    Q[(size_t)k * n * n + i + j * n] = Ql[i + j * n];
    Q[(size_t)k * n * n + j + i * n] = Ql[i + j * n];
  }
  for (int j = 0; j < m; ++j) r[(size_t)k * m + j] = rl_[j];
  for (int i = 0; i < m; ++i) for (int j = i; j < m; ++j) {
    R[(size_t)k * m * m + i + j * m] = Rl[i + j * m];
    R[(size_t)k * m * m + j + i * m] = Rl[i + j * m];
  }
  for (int i = 0; i < m * n; ++i) Pm[(size_t)k * m * n + i] = Pl[i];
  for (int i = 0; i < n * n; ++i) A[(size_t)k * n * n + i] = Al[i];
  for (int i = 0; i < n * m; ++i) Bm[(size_t)k * n * m + i] = Bl[i];
  return 0;
}

// solve_approximate_dp! / solve_approximate_dp on caller-supplied approximations (ileqg.jl:341-465)
template <int n, int m>
RL_HD int comp_riccati(int N, int optimise, const double* q, const double* qv, const double* Q, const double* r,
                       const double* R, const double* Pm, const double* A, const double* Bm, const double* W,
                       const double* Winv, const double* detWs, int W_tv, double theta, double mu_min, double delta_0, double* mu,
                       double* delta, double* L, double* dl, double* s, double* sv, double* S, int32_t* restarts) {
  // W_tv: W, Winv are n*n*N and detWs has N entries (W(k) of stage k, ileqg.jl:364,438); else one shared matrix
  int nrestart = 0;
  while (true) {
    double Sc[n * n], svc[n], sc;
    sc = q[N];
    for (int i = 0; i < n; ++i) svc[i] = qv[(size_t)N * n + i];
    for (int i = 0; i < n; ++i) for (int j = i; j < n; ++j) {
      Sc[i + j * n] = Q[(size_t)N * n * n + i + j * n];
      Sc[j + i * n] = Sc[i + j * n];
    }
    s[N] = sc;
    for (int i = 0; i < n; ++i) sv[(size_t)N * n + i] = svc[i];
    for (int i = 0; i < n * n; ++i) S[(size_t)N * n * n + i] = Sc[i];
    bool restart = false;
    for (int k = N - 1; k >= 0; --k) {
      double qvl[n], Ql[n * n], rl_[m], Rl[m * m], Pl[m * n], Al[n * n], Bl[n * m], Ll[m * n], dll[m];
      for (int i = 0; i < n; ++i) qvl[i] = qv[(size_t)k * n + i];
      for (int i = 0; i < n * n; ++i) { Ql[i] = Q[(size_t)k * n * n + i]; Al[i] = A[(size_t)k * n * n + i]; }
      for (int i = 0; i < m; ++i) rl_[i] = r[(size_t)k * m + i];
      for (int i = 0; i < m * m; ++i) Rl[i] = R[(size_t)k * m * m + i];
      for (int i = 0; i < m * n; ++i) { Pl[i] = Pm[(size_t)k * m * n + i]; Bl[i] = Bm[(size_t)k * n * m + i]; }
      int rc;
      const double* Wk = W + (W_tv ? (size_t)k * n * n : 0);
      const double* Wik = Winv + (W_tv ? (size_t)k * n * n : 0);
      const double detW = detWs[W_tv ? k : 0];
      if (optimise) {
        rc = riccati_stage<DenseTraits<n, m>, true, true>(theta, *mu, Wk, Wik, detW, Sc, svc, sc, q[k], qvl, Ql, rl_, Rl, Pl, Al, Bl, Ll, dll);
      } else {
        for (int i = 0; i < m * n; ++i) Ll[i] = L[(size_t)k * m * n + i];
        if (dl) {
          for (int i = 0; i < m; ++i) dll[i] = dl[(size_t)k * m + i];
          rc = riccati_stage<DenseTraits<n, m>, false, true>(theta, *mu, Wk, Wik, detW, Sc, svc, sc, q[k], qvl, Ql, rl_, Rl, Pl, Al, Bl, Ll, dll);
        } else {
          rc = riccati_stage<DenseTraits<n, m>, false, false>(theta, *mu, Wk, Wik, detW, Sc, svc, sc, q[k], qvl, Ql, rl_, Rl, Pl, Al, Bl, Ll, dll);
        }
      }
      if (rc == 1) { *restarts = nrestart; return optimise ? RATILQR_ST_M_NOT_PD_OPT : RATILQR_ST_M_NOT_PD_INIT; }
      if (rc == 2) {
        *delta = fmax(delta_0, *delta * delta_0);
        *mu = fmax(mu_min, *mu * *delta);
        nrestart++;
        if (!(*mu < 1e300)) { *restarts = nrestart; return RATILQR_ST_MU_OVERFLOW; }
        restart = true;
        break;
      }
      if (optimise) {
        for (int i = 0; i < m * n; ++i) L[(size_t)k * m * n + i] = Ll[i];
        for (int i = 0; i < m; ++i) dl[(size_t)k * m + i] = dll[i];
      }
      s[k] = sc;
      for (int i = 0; i < n; ++i) sv[(size_t)k * n + i] = svc[i];
      for (int i = 0; i < n * n; ++i) S[(size_t)k * n * n + i] = Sc[i];
    }
    if (!restart) break;
  }
  *restarts = nrestart;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG + Box-Muller: noise is a pure function of
// (seed, stream index, step), hence independent of the launch shape and of the GPU count.
// ---------------------------------------------------------------------------------------------
RL_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two independent standard normals from one Philox block (53-bit uniforms, Box-Muller)
RL_HD void philox_normal2(uint64_t seed, uint64_t stream, uint32_t step, uint32_t lane, double* z0, double* z1) {
  uint32_t o[4];
  philox4x32_10((uint32_t)stream, (uint32_t)(stream >> 32), step, lane, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  double u1 = ((double)(((uint64_t)o[0] << 21) ^ (uint64_t)(o[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  double u2 = ((double)(((uint64_t)o[2] << 21) ^ (uint64_t)(o[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  double rad = sqrt(-2.0 * log(u1));
  double ang = 6.283185307179586476925286766559 * u2;
  double sn, cs;
  rl_sincos_any(ang, &sn, &cs);  // one range reduction for the pair (same bits as cos(ang), sin(ang))
  *z0 = rad * cs;
  *z1 = rad * sn;
}
RL_HD double philox_uniform(uint64_t seed, uint64_t stream, uint32_t step, uint32_t lane, int which) {
  uint32_t o[4];
  philox4x32_10((uint32_t)stream, (uint32_t)(stream >> 32), step, lane, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  uint32_t hi = which ? o[2] : o[0], lo = which ? o[3] : o[1];
  return (double)(((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11)) * (1.0 / 9007199254740992.0);
}

// w = chol_lower(W) * z (or uniform*scale), z from Philox(seed, stream, step)
template <int n>
RL_HD void philox_noise(uint64_t seed, uint64_t stream, uint32_t step, int kind, double scale, const double* cholW, double* w) {
  double z[n + 1];
  if (kind == 1) {
    for (int i = 0; i < n; ++i) w[i] = scale * philox_uniform(seed, stream, step, (uint32_t)(i >> 1), i & 1);
    return;
  }
  for (int i = 0; i < n; i += 2) philox_normal2(seed, stream, step, (uint32_t)(i >> 1), &z[i], &z[i + 1]);
  for (int i = 0; i < n; ++i) {
    double a = cholW[i] * z[0];
    for (int k = 1; k <= i; ++k) a = rl_fma(cholW[i + k * n], z[k], a);
    w[i] = a;
  }
}

// "true model" noise: Gaussian mixture (optimal_control_problems.jl:105-109).  cumw: cumulative normalised weights,
// mean n*k, chol n*n*k (lower factors), prepared on the host (rlh::prep_mixture).  The component index consumes its
// own Philox counter lane, so the normals of a step are the same draws the Gaussian model would use.
struct MixtureView { int k; const double* cumw; const double* mean; const double* chol; };
template <int n>
RL_HD void philox_mixture_noise(uint64_t seed, uint64_t stream, uint32_t step, const MixtureView& mx, double* w) {
  const double uu = philox_uniform(seed, stream, step, 0xFFFFu, 0);
  int c = 0;
  while (c < mx.k - 1 && uu >= mx.cumw[c]) ++c;
  double z[n + 1];
  for (int i = 0; i < n; i += 2) philox_normal2(seed, stream, step, (uint32_t)(i >> 1), &z[i], &z[i + 1]);
  const double* C = mx.chol + (size_t)c * n * n;
  for (int i = 0; i < n; ++i) {
    double a = mx.mean[(size_t)c * n + i];
    for (int k = 0; k <= i; ++k) a = rl_fma(C[i + k * n], z[k], a);
    w[i] = a;
  }
}

// one PETS particle: compute_cost_serial's inner loops for (sequence ii, particle kk) (pets.jl:141-152)
template <class D, class CT>
RL_HD double comp_pets_particle(const double* mp, const double* cp, int N, const double* x0, const double* useq,
                                const double* w /*n*N or null*/, uint64_t seed, uint64_t stream, int noise_kind,
                                double noise_scale, const double* cholW, const MixtureView* mx = nullptr) {
  constexpr int n = D::n, m = D::m;
  double x[n], xn[n], u[m], wk[n];
  for (int i = 0; i < n; ++i) x[i] = x0[i];
  double c = 0.0;
  for (int tt = 0; tt < N; ++tt) {
    for (int j = 0; j < m; ++j) u[j] = useq[(size_t)tt * m + j];
    double q;
    if (!CT::stage(cp, tt, x, u, false, q, nullptr, nullptr, nullptr, nullptr, nullptr)) return rl_inf();
    c += q;
    if (!D::f(mp, x, u, xn)) return rl_inf();
    if (w) { for (int i = 0; i < n; ++i) wk[i] = w[(size_t)tt * n + i]; }
    else if (mx && mx->k > 0) philox_mixture_noise<n>(seed, stream, (uint32_t)tt, *mx, wk);
    else philox_noise<n>(seed, stream, (uint32_t)tt, noise_kind, noise_scale, cholW, wk);
    for (int i = 0; i < n; ++i) x[i] = xn[i] + wk[i];
  }
  double q;
  if (!CT::terminal(cp, x, false, q, nullptr, nullptr)) return rl_inf();
  return c + q;
}

}  // namespace rl
