// rl_user.cuh -- user-extensible device models: the device half.
//
// The reference differentiates arbitrary Julia closures with ForwardDiff (ileqg.jl:265-273: cx, cxx, cu, cuu, cux
// on the cost; fx, fu on the dynamics).  The device analogue: the user hands the library a CUDA C++ snippet that
// defines the model ONCE as a template over the scalar type; the library compiles it at run time (NVRTC, sm_100a)
// together with these headers and instantiates it with
//   double        -- rollouts, cost integration
//   Dual<n+m>     -- first-order forward mode: A = df/dx, B = df/du
//   Dual2<n+m>    -- second-order forward mode: grad_x c, grad_u c, hess_xx c, hess_uu c, P = d(grad_u c)/dx
// and plugs the result into the same kernels the registered models use (rl_kernels_model.cuh): the persistent
// iLEQG solve, the rollouts, approximate_model, the Monte Carlo and the PETS rollouts.
//
// What the snippet must define (any helper functions / constants are fine; functions are __device__ by default):
//   dynamics:  template <class T> void dynamics(const double* p, const T* x, const T* u, T* xn);
//   cost:      template <class T> T stage_cost(const double* cp, int k, const T* x, const T* u);   // k is 0-based
//              template <class T> T terminal_cost(const double* cp, const T* x);
// with sin, cos, tan, exp, log, sqrt, tanh, atan, fabs, pow(T, double), square(T) and + - * / < > available for
// every T.  A NaN produced by the model (sqrt/log/pow outside their domain) is reported as RATILQR_ST_DOMAIN for
// that instance -- where Julia would have thrown a DomainError (test/ileqg_test.jl:151-174).
//
// This header is only ever compiled by NVRTC (and, for its arithmetic, by the CPU tests through RL_HD).
#pragma once
#include "rl_core.cuh"

namespace rl {

// the overloads below must not hide the double versions from code inside namespace rl
using ::sin; using ::cos; using ::tan; using ::exp; using ::log; using ::sqrt; using ::tanh; using ::atan; using ::fabs; using ::pow;

// ---- first-order duals: standard math names (found by argument-dependent lookup from user code) -------------------
template <int NP> RL_HD Dual<NP> chain1(const Dual<NP>& a, double f0, double f1) {
  Dual<NP> r; r.v = f0;
  for (int i = 0; i < NP; ++i) r.d[i] = f1 * a.d[i];
  return r;
}
template <int NP> RL_HD Dual<NP> operator-(double b, const Dual<NP>& a) { Dual<NP> r; r.v = b - a.v; for (int i = 0; i < NP; ++i) r.d[i] = -a.d[i]; return r; }
template <int NP> RL_HD Dual<NP> operator/(double b, const Dual<NP>& a) { double inv = 1.0 / a.v; return chain1(a, b * inv, -b * inv * inv); }
template <int NP> RL_HD Dual<NP> sin(const Dual<NP>& a) { return chain1(a, ::sin(a.v), ::cos(a.v)); }
template <int NP> RL_HD Dual<NP> cos(const Dual<NP>& a) { return chain1(a, ::cos(a.v), -::sin(a.v)); }
template <int NP> RL_HD Dual<NP> tan(const Dual<NP>& a) { double t = ::tan(a.v); return chain1(a, t, 1.0 + t * t); }
template <int NP> RL_HD Dual<NP> exp(const Dual<NP>& a) { double e = ::exp(a.v); return chain1(a, e, e); }
template <int NP> RL_HD Dual<NP> log(const Dual<NP>& a) { return chain1(a, ::log(a.v), 1.0 / a.v); }
template <int NP> RL_HD Dual<NP> sqrt(const Dual<NP>& a) { double s = ::sqrt(a.v); return chain1(a, s, 0.5 / s); }
template <int NP> RL_HD Dual<NP> tanh(const Dual<NP>& a) { double t = ::tanh(a.v); return chain1(a, t, 1.0 - t * t); }
template <int NP> RL_HD Dual<NP> atan(const Dual<NP>& a) { return chain1(a, ::atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
template <int NP> RL_HD Dual<NP> fabs(const Dual<NP>& a) { return chain1(a, ::fabs(a.v), a.v < 0.0 ? -1.0 : 1.0); }
template <int NP> RL_HD Dual<NP> pow(const Dual<NP>& a, double e) { double p1 = ::pow(a.v, e - 1.0); return chain1(a, ::pow(a.v, e), e * p1); }
template <int NP> RL_HD Dual<NP> square(const Dual<NP>& a) { return chain1(a, a.v * a.v, 2.0 * a.v); }
template <int NP> RL_HD bool operator<(const Dual<NP>& a, const Dual<NP>& b) { return a.v < b.v; }
template <int NP> RL_HD bool operator>(const Dual<NP>& a, const Dual<NP>& b) { return a.v > b.v; }
template <int NP> RL_HD bool operator<(const Dual<NP>& a, double b) { return a.v < b; }
template <int NP> RL_HD bool operator>(const Dual<NP>& a, double b) { return a.v > b; }
RL_HD double square(double a) { return a * a; }

// ---- second-order forward mode: value, gradient (NP) and packed symmetric Hessian (NP(NP+1)/2, i >= j) --------
template <int NP>
struct Dual2 {
  static constexpr int NH = NP * (NP + 1) / 2;
  double v;
  double g[NP];
  double h[NH];
  RL_HD static constexpr int idx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
};
template <int NP> RL_HD Dual2<NP> d2const(double v) {
  Dual2<NP> r; r.v = v;
  for (int i = 0; i < NP; ++i) r.g[i] = 0.0;
  for (int i = 0; i < Dual2<NP>::NH; ++i) r.h[i] = 0.0;
  return r;
}
// r = f(a): r' = f1 a', r'' = f1 a'' + f2 a' a'^T
template <int NP> RL_HD Dual2<NP> chain2(const Dual2<NP>& a, double f0, double f1, double f2) {
  Dual2<NP> r; r.v = f0;
  for (int i = 0; i < NP; ++i) r.g[i] = f1 * a.g[i];
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j <= i; ++j) r.h[i * (i + 1) / 2 + j] = f1 * a.h[i * (i + 1) / 2 + j] + f2 * a.g[i] * a.g[j];
  return r;
}
template <int NP> RL_HD Dual2<NP> operator+(const Dual2<NP>& a, const Dual2<NP>& b) {
  Dual2<NP> r; r.v = a.v + b.v;
  for (int i = 0; i < NP; ++i) r.g[i] = a.g[i] + b.g[i];
  for (int i = 0; i < Dual2<NP>::NH; ++i) r.h[i] = a.h[i] + b.h[i];
  return r;
}
template <int NP> RL_HD Dual2<NP> operator-(const Dual2<NP>& a, const Dual2<NP>& b) {
  Dual2<NP> r; r.v = a.v - b.v;
  for (int i = 0; i < NP; ++i) r.g[i] = a.g[i] - b.g[i];
  for (int i = 0; i < Dual2<NP>::NH; ++i) r.h[i] = a.h[i] - b.h[i];
  return r;
}
template <int NP> RL_HD Dual2<NP> operator-(const Dual2<NP>& a) { return chain2(a, -a.v, -1.0, 0.0); }
template <int NP> RL_HD Dual2<NP> operator*(const Dual2<NP>& a, const Dual2<NP>& b) {
  Dual2<NP> r; r.v = a.v * b.v;
  for (int i = 0; i < NP; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j <= i; ++j) {
      const int e = i * (i + 1) / 2 + j;
      r.h[e] = (a.h[e] * b.v + a.v * b.h[e]) + (a.g[i] * b.g[j] + a.g[j] * b.g[i]);
    }
  return r;
}
template <int NP> RL_HD Dual2<NP> recip(const Dual2<NP>& a) { double inv = 1.0 / a.v; return chain2(a, inv, -inv * inv, 2.0 * inv * inv * inv); }
template <int NP> RL_HD Dual2<NP> operator/(const Dual2<NP>& a, const Dual2<NP>& b) { return a * recip(b); }
template <int NP> RL_HD Dual2<NP> operator+(const Dual2<NP>& a, double b) { Dual2<NP> r = a; r.v = a.v + b; return r; }
template <int NP> RL_HD Dual2<NP> operator+(double b, const Dual2<NP>& a) { Dual2<NP> r = a; r.v = b + a.v; return r; }
template <int NP> RL_HD Dual2<NP> operator-(const Dual2<NP>& a, double b) { Dual2<NP> r = a; r.v = a.v - b; return r; }
template <int NP> RL_HD Dual2<NP> operator-(double b, const Dual2<NP>& a) { return chain2(a, b - a.v, -1.0, 0.0); }
template <int NP> RL_HD Dual2<NP> operator*(const Dual2<NP>& a, double b) { return chain2(a, a.v * b, b, 0.0); }
template <int NP> RL_HD Dual2<NP> operator*(double b, const Dual2<NP>& a) { return chain2(a, a.v * b, b, 0.0); }
template <int NP> RL_HD Dual2<NP> operator/(const Dual2<NP>& a, double b) { double inv = 1.0 / b; return chain2(a, a.v * inv, inv, 0.0); }
template <int NP> RL_HD Dual2<NP> operator/(double b, const Dual2<NP>& a) { return recip(a) * b; }
template <int NP> RL_HD Dual2<NP> sin(const Dual2<NP>& a) { double s = ::sin(a.v), c = ::cos(a.v); return chain2(a, s, c, -s); }
template <int NP> RL_HD Dual2<NP> cos(const Dual2<NP>& a) { double s = ::sin(a.v), c = ::cos(a.v); return chain2(a, c, -s, -c); }
template <int NP> RL_HD Dual2<NP> tan(const Dual2<NP>& a) { double t = ::tan(a.v), d = 1.0 + t * t; return chain2(a, t, d, 2.0 * t * d); }
template <int NP> RL_HD Dual2<NP> exp(const Dual2<NP>& a) { double e = ::exp(a.v); return chain2(a, e, e, e); }
template <int NP> RL_HD Dual2<NP> log(const Dual2<NP>& a) { double inv = 1.0 / a.v; return chain2(a, ::log(a.v), inv, -inv * inv); }
template <int NP> RL_HD Dual2<NP> sqrt(const Dual2<NP>& a) { double s = ::sqrt(a.v); return chain2(a, s, 0.5 / s, -0.25 / (s * a.v)); }
template <int NP> RL_HD Dual2<NP> tanh(const Dual2<NP>& a) { double t = ::tanh(a.v), d = 1.0 - t * t; return chain2(a, t, d, -2.0 * t * d); }
template <int NP> RL_HD Dual2<NP> atan(const Dual2<NP>& a) { double d = 1.0 / (1.0 + a.v * a.v); return chain2(a, ::atan(a.v), d, -2.0 * a.v * d * d); }
template <int NP> RL_HD Dual2<NP> fabs(const Dual2<NP>& a) { return chain2(a, ::fabs(a.v), a.v < 0.0 ? -1.0 : 1.0, 0.0); }
template <int NP> RL_HD Dual2<NP> pow(const Dual2<NP>& a, double e) {
  return chain2(a, ::pow(a.v, e), e * ::pow(a.v, e - 1.0), e * (e - 1.0) * ::pow(a.v, e - 2.0));
}
template <int NP> RL_HD Dual2<NP> square(const Dual2<NP>& a) { return chain2(a, a.v * a.v, 2.0 * a.v, 2.0); }
template <int NP> RL_HD bool operator<(const Dual2<NP>& a, const Dual2<NP>& b) { return a.v < b.v; }
template <int NP> RL_HD bool operator>(const Dual2<NP>& a, const Dual2<NP>& b) { return a.v > b.v; }
template <int NP> RL_HD bool operator<(const Dual2<NP>& a, double b) { return a.v < b; }
template <int NP> RL_HD bool operator>(const Dual2<NP>& a, double b) { return a.v > b; }

RL_HD bool rl_finite(double v) { return v - v == 0.0; }  // false for NaN and +-Inf

// ---- adapters: a user snippet -> the Dyn / Cost interfaces of rl_core.cuh ------------------------------------------
// Body: struct with  template <class T> void operator()(const double* p, const T* x, const T* u, T* xn) const
// Kinds: struct with  static constexpr bool structured;  static constexpr int a_kind(int i, int j), b_kind(int i, int j)
// (DenseKinds, or the user's declared structure generated by rl_user_host.cu)
template <int N_, int M_, class Body, class Kinds = DenseKinds>
struct UserDyn : Kinds {
  static constexpr int n = N_, m = M_;
  RL_HD static bool f(const double* p, const double* x, const double* u, double* xn) {
    Body()(p, x, u, xn);
    bool ok = true;
    for (int i = 0; i < n; ++i) ok = ok && rl_finite(xn[i]);
    return ok;  // NaN / Inf <=> Julia DomainError
  }
  RL_HD static void jac(const double* p, const double* x, const double* u, double* A, double* B) {  // fx, fu  ileqg.jl:272-273
    dual_jacobian<n, m>(Body(), p, x, u, A, B);
  }
};

// Fn: struct with  template <class T> T stage(const double* cp, int k, const T* x, const T* u) const
//                  template <class T> T terminal(const double* cp, const T* x) const
struct DenseCostKinds {
  RL_HD static constexpr int q_kind(int, int) { return 2; }
  RL_HD static constexpr int r_kind(int, int) { return 2; }
  RL_HD static constexpr int p_kind(int, int) { return 2; }
};
template <int n, int m, int NPAR_, class Fn, class CKinds = DenseCostKinds>
struct UserCost : CKinds {
  static constexpr int NPAR = NPAR_;
  RL_HD static bool stage(const double* RL_RESTRICT cp, int k, const double* x, const double* u, bool der,
                          double& q, double* qv, double* Q, double* r, double* R, double* P) {
    if (!der) {
      q = Fn().template stage<double>(cp, k, x, u);
      return rl_finite(q);
    }
    constexpr int NP = n + m;
    typedef Dual2<NP> T;
    T xd[n], ud[m];
    for (int i = 0; i < n; ++i) { xd[i] = d2const<NP>(x[i]); xd[i].g[i] = 1.0; }
    for (int j = 0; j < m; ++j) { ud[j] = d2const<NP>(u[j]); ud[j].g[n + j] = 1.0; }
    T c = Fn().template stage<T>(cp, k, xd, ud);
    q = c.v;                                                                         // c        ileqg.jl:296
    bool ok = rl_finite(c.v);
    for (int i = 0; i < n; ++i) { qv[i] = c.g[i]; ok = ok && rl_finite(c.g[i]); }    // cx       :265
    for (int j = 0; j < m; ++j) { r[j] = c.g[n + j]; ok = ok && rl_finite(c.g[n + j]); }  // cu  :267
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) Q[i + j * n] = c.h[T::idx(i, j)];          // cxx (Symmetric) :266
    for (int j = 0; j < m; ++j) for (int i = 0; i < m; ++i) R[i + j * m] = c.h[T::idx(n + i, n + j)];  // cuu  :268
    for (int i = 0; i < n; ++i) for (int j = 0; j < m; ++j) P[j + i * m] = c.h[T::idx(n + j, i)];      // cux  :269  (m x n)
    for (int e = 0; e < T::NH; ++e) ok = ok && rl_finite(c.h[e]);
    return ok;
  }
  RL_HD static bool terminal(const double* RL_RESTRICT cp, const double* x, bool der, double& q, double* qv, double* Q) {
    if (!der) {
      q = Fn().template terminal<double>(cp, x);
      return rl_finite(q);
    }
    typedef Dual2<n> T;
    T xd[n];
    for (int i = 0; i < n; ++i) { xd[i] = d2const<n>(x[i]); xd[i].g[i] = 1.0; }
    T c = Fn().template terminal<T>(cp, xd);
    q = c.v;
    bool ok = rl_finite(c.v);
    for (int i = 0; i < n; ++i) { qv[i] = c.g[i]; ok = ok && rl_finite(c.g[i]); }
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) Q[i + j * n] = c.h[T::idx(i, j)];
    for (int e = 0; e < T::NH; ++e) ok = ok && rl_finite(c.h[e]);
    return ok;
  }
};

}  // namespace rl
