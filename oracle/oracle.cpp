// oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A CPU restatement of the reference algorithm (StanfordMSL/RATiLQR.jl, pure Julia) for the
// hot path named in BASELINE.json.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product
// (ratilqr.jl_b200/csrc) never links, imports or calls it.
//
// PARITY STATUS: the reference cannot be executed in this image (no Julia toolchain,
// no network).  This restatement is pinned against the reference's own known-answer
// tests (test/ileqg_test.jl, test/*_test.jl: closed-form linearisations, theta=0 LQR
// gains, pass-consistency, schedule values -- see tests/test_oracle_reference_suite.py).
// Everything that depends on Julia's RNG streams (MersenneTwister/ziggurat) is
// "parity unpinned": randomness is injected as tensors instead.
//
// Each function cites the reference file:line it follows.  The code is written with
// run-time sizes and Julia-like exceptions on purpose: it shares no source with the CUDA
// kernels (which are compile-time-sized and exception-free).  Third-party arithmetic
// (ForwardDiff 0.10.12, LAPACK factorizations, Distributions 0.24.2) is restated as:
// forward-mode dual numbers for df/dx, df/du; closed-form gradients/Hessians for the
// registered cost families; Cholesky factorizations for M and H.
//
// Floating-point policy (DESIGN.md "Arithmetic order"): compile with -ffp-contract=off;
// inner products accumulate in index order with explicit fma, so that the GPU kernels,
// which follow the same order, agree to the last few ulps (libm calls excepted).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/ratilqr.h"

namespace {

constexpr double INF = std::numeric_limits<double>::infinity();
inline double FMA(double a, double b, double c) { return __builtin_fma(a, b, c); }

struct DomainError {};  // Julia DomainError (negative base of a real power)
struct NotPosDef {};    // failed `@assert isposdef(M)`  ileqg.jl:366,440

typedef std::vector<double> vec;

// strided inner product, canonical order: a0*b0, then fma in increasing index
inline double dotp(int n, const double* a, int sa, const double* b, int sb) {
  double acc = a[0] * b[0];
  for (int i = 1; i < n; ++i) acc = FMA(a[i * sa], b[i * sb], acc);
  return acc;
}

// ---------------------------------------------------------------------------------------
// forward-mode dual numbers (restates what ForwardDiff.jacobian does at ileqg.jl:265-266)
// ---------------------------------------------------------------------------------------
constexpr int MAXP = 16;
struct Dual {
  double v;
  double d[MAXP];
  int np;
};
inline Dual mk(double v, int np) { Dual r; r.v = v; r.np = np; for (int i = 0; i < np; ++i) r.d[i] = 0.0; return r; }
inline Dual operator+(const Dual& a, const Dual& b) { Dual r = mk(a.v + b.v, a.np); for (int i = 0; i < a.np; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
inline Dual operator-(const Dual& a, const Dual& b) { Dual r = mk(a.v - b.v, a.np); for (int i = 0; i < a.np; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
inline Dual operator-(const Dual& a) { Dual r = mk(-a.v, a.np); for (int i = 0; i < a.np; ++i) r.d[i] = -a.d[i]; return r; }
inline Dual operator*(const Dual& a, const Dual& b) { Dual r = mk(a.v * b.v, a.np); for (int i = 0; i < a.np; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
inline Dual operator/(const Dual& a, const Dual& b) {
  Dual r = mk(a.v / b.v, a.np);
  double inv = 1.0 / b.v;
  for (int i = 0; i < a.np; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
inline Dual operator+(const Dual& a, double b) { Dual r = a; r.v = a.v + b; return r; }
inline Dual operator+(double b, const Dual& a) { Dual r = a; r.v = b + a.v; return r; }
inline Dual operator-(const Dual& a, double b) { Dual r = a; r.v = a.v - b; return r; }
inline Dual operator-(double b, const Dual& a) { Dual r = -a; r.v = b - a.v; return r; }
inline Dual operator*(const Dual& a, double b) { Dual r = mk(a.v * b, a.np); for (int i = 0; i < a.np; ++i) r.d[i] = a.d[i] * b; return r; }
inline Dual operator*(double b, const Dual& a) { return a * b; }
inline Dual operator/(const Dual& a, double b) { Dual r = mk(a.v / b, a.np); for (int i = 0; i < a.np; ++i) r.d[i] = a.d[i] / b; return r; }
inline Dual sin(const Dual& a) { Dual r = mk(std::sin(a.v), a.np); double c = std::cos(a.v); for (int i = 0; i < a.np; ++i) r.d[i] = c * a.d[i]; return r; }
inline Dual cos(const Dual& a) { Dual r = mk(std::cos(a.v), a.np); double s = -std::sin(a.v); for (int i = 0; i < a.np; ++i) r.d[i] = s * a.d[i]; return r; }
inline Dual tan(const Dual& a) { double t = std::tan(a.v); Dual r = mk(t, a.np); double s = 1.0 + t * t; for (int i = 0; i < a.np; ++i) r.d[i] = s * a.d[i]; return r; }
// x^p for a real constant p: DiffRules gives p*x^(p-1); a negative base is a DomainError in Julia
inline double rpow(double x, double p) { if (x < 0.0) throw DomainError(); return std::pow(x, p); }
inline Dual rpow(const Dual& a, double p) {
  Dual r = mk(rpow(a.v, p), a.np);
  double g = p * std::pow(a.v, p - 1.0);
  for (int i = 0; i < a.np; ++i) r.d[i] = g * a.d[i];
  return r;
}
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double tan(double x) { return std::tan(x); }

// ---------------------------------------------------------------------------------------
// registered dynamics (DESIGN.md "Model registry"); templated on the scalar so the same
// expression yields values (double) and Jacobians (Dual)
// ---------------------------------------------------------------------------------------
template <class T>
void dynamics(int id, const double* p, const T* x, const T* u, T* xn) {
  switch (id) {
    case RATILQR_MODEL_SINGLE_INTEGRATOR: {  // test/ileqg_test.jl:12 (dt=1), getting-started.md
      double dt = p[0];
      xn[0] = x[0] + dt * u[0];
      xn[1] = x[1] + dt * u[1];
    } break;
    case RATILQR_MODEL_POWER_LAW: {  // test/ileqg_test.jl:151  f(x,u) = x.^1.3 + u.^1.5
      xn[0] = rpow(x[0], p[0]) + rpow(u[0], p[1]);
      xn[1] = rpow(x[1], p[0]) + rpow(u[1], p[1]);
    } break;
    case RATILQR_MODEL_DOUBLE_INTEGRATOR: {
      double dt = p[0];
      xn[0] = x[0] + dt * x[2];
      xn[1] = x[1] + dt * x[3];
      xn[2] = x[2] + dt * u[0];
      xn[3] = x[3] + dt * u[1];
    } break;
    case RATILQR_MODEL_PENDULUM: {
      double dt = p[0], g = p[1], len = p[2], mass = p[3], damp = p[4];
      double inertia = mass * len * len;
      T alpha = (u[0] - damp * x[1] - (mass * g * len) * sin(x[0])) / inertia;
      xn[0] = x[0] + dt * x[1];
      xn[1] = x[1] + dt * alpha;
    } break;
    case RATILQR_MODEL_CARTPOLE: {  // state (pos, th, vel, thd); th from the downward vertical
      double dt = p[0], mc = p[1], mp = p[2], len = p[3], g = p[4];
      T s = sin(x[1]), c = cos(x[1]);
      T den = mc + mp * (s * s);
      T thd2 = x[3] * x[3];
      T acc = (u[0] + mp * s * (len * thd2 + g * c)) / den;
      T thacc = (-(u[0] * c) - (mp * len) * thd2 * c * s - ((mc + mp) * g) * s) / (len * den);
      xn[0] = x[0] + dt * x[2];
      xn[1] = x[1] + dt * x[3];
      xn[2] = x[2] + dt * acc;
      xn[3] = x[3] + dt * thacc;
    } break;
    case RATILQR_MODEL_UNICYCLE: {  // (px, py, psi, v ; a, omega)
      double dt = p[0];
      T s = sin(x[2]), c = cos(x[2]);
      xn[0] = x[0] + dt * (x[3] * c);
      xn[1] = x[1] + dt * (x[3] * s);
      xn[2] = x[2] + dt * u[1];
      xn[3] = x[3] + dt * u[0];
    } break;
    case RATILQR_MODEL_QUADROTOR: {  // p(3), euler phi/th/psi (3), v world (3), omega body (3); u = thrust, torques
      double dt = p[0], mass = p[1], g = p[2], Ix = p[3], Iy = p[4], Iz = p[5];
      T sph = sin(x[3]), cph = cos(x[3]);
      T sth = sin(x[4]), cth = cos(x[4]);
      T sps = sin(x[5]), cps = cos(x[5]);
      T tth = sth / cth;
      T wp = x[9], wq = x[10], wr = x[11];
      T qr = wq * sph + wr * cph;
      T dphi = wp + qr * tth;
      T dth = wq * cph - wr * sph;
      T dpsi = qr / cth;
      T tm = u[0] / mass;
      T ax = tm * (cph * sth * cps + sph * sps);
      T ay = tm * (cph * sth * sps - sph * cps);
      T az = tm * (cph * cth) - g;
      T dwp = (u[1] + (Iy - Iz) * (wq * wr)) / Ix;
      T dwq = (u[2] + (Iz - Ix) * (wp * wr)) / Iy;
      T dwr = (u[3] + (Ix - Iy) * (wp * wq)) / Iz;
      xn[0] = x[0] + dt * x[6];
      xn[1] = x[1] + dt * x[7];
      xn[2] = x[2] + dt * x[8];
      xn[3] = x[3] + dt * dphi;
      xn[4] = x[4] + dt * dth;
      xn[5] = x[5] + dt * dpsi;
      xn[6] = x[6] + dt * ax;
      xn[7] = x[7] + dt * ay;
      xn[8] = x[8] + dt * az;
      xn[9] = x[9] + dt * dwp;
      xn[10] = x[10] + dt * dwq;
      xn[11] = x[11] + dt * dwr;
    } break;
    default: break;
  }
}

bool model_dims(int id, int* n, int* m, int* np) {
  switch (id) {
    case RATILQR_MODEL_SINGLE_INTEGRATOR: *n = 2; *m = 2; *np = 1; return true;
    case RATILQR_MODEL_POWER_LAW: *n = 2; *m = 2; *np = 2; return true;
    case RATILQR_MODEL_DOUBLE_INTEGRATOR: *n = 4; *m = 2; *np = 1; return true;
    case RATILQR_MODEL_PENDULUM: *n = 2; *m = 1; *np = 5; return true;
    case RATILQR_MODEL_CARTPOLE: *n = 4; *m = 1; *np = 5; return true;
    case RATILQR_MODEL_UNICYCLE: *n = 4; *m = 2; *np = 1; return true;
    case RATILQR_MODEL_QUADROTOR: *n = 12; *m = 4; *np = 6; return true;
  }
  return false;
}

// A problem instance as the Julia struct holds it (optimal_control_problems.jl:67-73)
struct Problem {
  int model_id, cost_id, n, m, N;
  const double* mp;
  const double* cp;  // this problem's cost parameter block
  const double* W;   // n*n or n*n*N
  int W_tv;
  const double* Wk(int k) const { return W + (W_tv ? (size_t)k * n * n : 0); }
};

void f_eval(const Problem& pr, const double* x, const double* u, double* xn) {
  dynamics<double>(pr.model_id, pr.mp, x, u, xn);
}

// fx, fu of ileqg.jl:265-266 (joint seeding of [x;u] gives the same partials)
void f_jac(const Problem& pr, const double* x, const double* u, double* A, double* B) {
  int n = pr.n, m = pr.m, np = n + m;
  Dual xd[12], ud[4], xo[12];
  for (int i = 0; i < n; ++i) { xd[i] = mk(x[i], np); xd[i].d[i] = 1.0; }
  for (int j = 0; j < m; ++j) { ud[j] = mk(u[j], np); ud[j].d[n + j] = 1.0; }
  for (int i = 0; i < n; ++i) xo[i] = mk(0.0, np);
  dynamics<Dual>(pr.model_id, pr.mp, xd, ud, xo);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) A[i + j * n] = xo[i].d[j];
    for (int j = 0; j < m; ++j) B[i + j * n] = xo[i].d[n + j];
  }
}

// ---------------------------------------------------------------------------------------
// registered costs: value + closed-form derivatives (what cx,cu,cxx,cuu,cux,hx,hxx of
// ileqg.jl:267-273 evaluate to).  P = d(grad_u c)/dx is m x n (ileqg.jl:269).
// ---------------------------------------------------------------------------------------
struct QuadView {
  double ws0, ws1, c0, c1, h0;
  const double *xg, *Q, *R, *Pc, *Qf;
  QuadView(const double* p, int n, int m) {
    ws0 = p[0]; ws1 = p[1]; c0 = p[2]; c1 = p[3]; h0 = p[4];
    xg = p + 5; Q = xg + n; R = Q + n * n; Pc = R + m * m; Qf = Pc + n * m;
  }
};

// stage cost; derivative outputs may be null (value only)
double cost_stage(const Problem& pr, int k, const double* x, const double* u,
                  double* qv, double* Q, double* r, double* R, double* P) {
  int n = pr.n, m = pr.m;
  bool der = qv != nullptr;
  if (pr.cost_id == RATILQR_COST_QUADRATIC) {
    QuadView c(pr.cp, n, m);
    double w = c.ws0 + c.ws1 * (double)k;
    double dx[12], Qdx[12], Pcu[12], Ru[4], Ptdx[4];
    for (int i = 0; i < n; ++i) dx[i] = x[i] - c.xg[i];
    for (int i = 0; i < n; ++i) Qdx[i] = dotp(n, c.Q + i, n, dx, 1);
    for (int i = 0; i < n; ++i) Pcu[i] = dotp(m, c.Pc + i, n, u, 1);
    for (int j = 0; j < m; ++j) Ru[j] = dotp(m, c.R + j, m, u, 1);
    for (int j = 0; j < m; ++j) Ptdx[j] = dotp(n, c.Pc + j * n, 1, dx, 1);
    double a = dotp(n, dx, 1, Qdx, 1), b = dotp(m, u, 1, Ru, 1), cc = dotp(n, dx, 1, Pcu, 1);
    double val = (w * ((0.5 * a + 0.5 * b) + cc) + c.c0) + c.c1 * (double)k;
    if (der) {
      for (int i = 0; i < n; ++i) qv[i] = w * (Qdx[i] + Pcu[i]);
      for (int j = 0; j < m; ++j) r[j] = w * (Ru[j] + Ptdx[j]);
      for (int i = 0; i < n * n; ++i) Q[i] = w * c.Q[i];
      for (int i = 0; i < m * m; ++i) R[i] = w * c.R[i];
      for (int j = 0; j < m; ++j) for (int i = 0; i < n; ++i) P[j + i * m] = w * c.Pc[i + j * n];
    }
    return val;
  } else if (pr.cost_id == RATILQR_COST_POWER_LAW) {  // sum(x.^p + u.^p): needs n == m
    double p = pr.cp[0];
    double val = 0.0;
    for (int i = 0; i < n; ++i) {
      double t = rpow(x[i], p) + rpow(u[i], p);
      val = (i == 0) ? t : val + t;
    }
    if (der) {
      for (int i = 0; i < n * n; ++i) Q[i] = 0.0;
      for (int i = 0; i < m * m; ++i) R[i] = 0.0;
      for (int i = 0; i < m * n; ++i) P[i] = 0.0;
      for (int i = 0; i < n; ++i) { qv[i] = p * std::pow(x[i], p - 1.0); Q[i + i * n] = p * ((p - 1.0) * std::pow(x[i], p - 2.0)); }
      for (int j = 0; j < m; ++j) { r[j] = p * std::pow(u[j], p - 1.0); R[j + j * m] = p * ((p - 1.0) * std::pow(u[j], p - 2.0)); }
    }
    return val;
  } else {  // L1_CONTROL: rollout-only
    double val = 0.0;
    for (int j = 0; j < m; ++j) val = (j == 0) ? std::fabs(u[j]) : val + std::fabs(u[j]);
    return val;
  }
}

double cost_terminal(const Problem& pr, const double* x, double* qv, double* Q) {
  int n = pr.n, m = pr.m;
  if (pr.cost_id == RATILQR_COST_QUADRATIC) {
    QuadView c(pr.cp, n, m);
    double dx[12], Qdx[12];
    for (int i = 0; i < n; ++i) dx[i] = x[i] - c.xg[i];
    for (int i = 0; i < n; ++i) Qdx[i] = dotp(n, c.Qf + i, n, dx, 1);
    double val = 0.5 * dotp(n, dx, 1, Qdx, 1) + c.h0;
    if (qv) {
      for (int i = 0; i < n; ++i) qv[i] = Qdx[i];
      for (int i = 0; i < n * n; ++i) Q[i] = c.Qf[i];
    }
    return val;
  }
  double h0 = (pr.cost_id == RATILQR_COST_POWER_LAW) ? pr.cp[1] : pr.cp[0];
  if (qv) {
    for (int i = 0; i < n; ++i) qv[i] = 0.0;
    for (int i = 0; i < n * n; ++i) Q[i] = 0.0;
  }
  return h0;
}

// ---------------------------------------------------------------------------------------
// W pre-processing: inv(W) (ileqg.jl:365), det(W) (for logdet(W*M), :387), chol(W) (noise)
// ---------------------------------------------------------------------------------------
struct WInfo { vec Winv, cholW; double detW; };

bool chol_lower(int n, const double* Msym, double* C, double* invd, double* det) {
  // uses the upper triangle of Msym, like Julia's Symmetric; C lower, column-major n x n
  double dprod = 1.0;
  for (int j = 0; j < n; ++j) {
    double d = Msym[j + j * n];
    for (int k = 0; k < j; ++k) d = FMA(-C[j + k * n], C[j + k * n], d);
    if (!(d > 0.0)) return false;
    dprod = (j == 0) ? d : dprod * d;
    double cjj = std::sqrt(d);
    double inv = 1.0 / cjj;
    C[j + j * n] = cjj;
    invd[j] = inv;
    for (int i = j + 1; i < n; ++i) {
      double a = Msym[j + i * n];
      for (int k = 0; k < j; ++k) a = FMA(-C[i + k * n], C[j + k * n], a);
      C[i + j * n] = a * inv;
    }
    for (int i = 0; i < j; ++i) C[i + j * n] = 0.0;
  }
  if (det) *det = dprod;
  return true;
}

bool prep_W(int n, const double* W, WInfo& o) {
  o.Winv.assign(n * n, 0.0);
  o.cholW.assign(n * n, 0.0);
  vec invd(n);
  if (!chol_lower(n, W, o.cholW.data(), invd.data(), &o.detW)) return false;
  // inv(W) = C^-T C^-1 : solve C Y = I (forward), then Winv = Y' Y
  vec Y(n * n, 0.0);
  const double* C = o.cholW.data();
  for (int c = 0; c < n; ++c)
    for (int i = 0; i < n; ++i) {
      double a = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) a = FMA(-C[i + k * n], Y[k + c * n], a);
      Y[i + c * n] = a * invd[i];
    }
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double e = dotp(n, Y.data() + i * n, 1, Y.data() + j * n, 1);
      o.Winv[i + j * n] = e;
      o.Winv[j + i * n] = e;
    }
  return true;
}

// ---------------------------------------------------------------------------------------
// approximate_model  ileqg.jl:258-322
// ---------------------------------------------------------------------------------------
struct Approx {
  int n, m, N;
  vec q, qv, Q, r, R, P, A, B;  // [N+1], [n(N+1)], [nn(N+1)], [mN], [mmN], [mnN], [nnN], [nmN]
  void resize(int n_, int m_, int N_) {
    n = n_; m = m_; N = N_;
    q.assign(N + 1, 0); qv.assign((size_t)n * (N + 1), 0); Q.assign((size_t)n * n * (N + 1), 0);
    r.assign((size_t)m * N, 0); R.assign((size_t)m * m * N, 0); P.assign((size_t)m * n * N, 0);
    A.assign((size_t)n * n * N, 0); B.assign((size_t)n * m * N, 0);
  }
};

void approximate_model(const Problem& pr, const double* u, const double* x, Approx& ap) {
  int n = pr.n, m = pr.m, N = pr.N;
  ap.resize(n, m, N);
  for (int ii = 0; ii < N; ++ii) {  // ileqg.jl:294-313, k = ii (0-based)
    const double* xk = x + (size_t)ii * n;
    const double* uk = u + (size_t)ii * m;
    ap.q[ii] = cost_stage(pr, ii, xk, uk, &ap.qv[(size_t)ii * n], &ap.Q[(size_t)ii * n * n],
                          &ap.r[(size_t)ii * m], &ap.R[(size_t)ii * m * m], &ap.P[(size_t)ii * m * n]);
    f_jac(pr, xk, uk, &ap.A[(size_t)ii * n * n], &ap.B[(size_t)ii * n * m]);
  }
  ap.q[N] = cost_terminal(pr, x + (size_t)N * n, &ap.qv[(size_t)N * n], &ap.Q[(size_t)N * n * n]);  // :314-316
}

// ---------------------------------------------------------------------------------------
// one stage of the risk-sensitive Riccati recursion (identical in ileqg.jl:360-395, :434-461)
// Canonical arithmetic order documented in DESIGN.md.  Throws NotPosDef for M.
// optimise: computes L, dl (returns false if H not PD -> caller increases mu and restarts)
// evaluate: L given, dl given or nullptr.
// ---------------------------------------------------------------------------------------
bool riccati_stage(int n, int m, bool optimise, double theta, double mu,
                   const double* W, const double* Winv, double detW,
                   const double* Sp, const double* svp, double sp,
                   double q, const double* qv, const double* Q, const double* r, const double* R,
                   const double* P, const double* A, const double* B,
                   double* L, double* dl, double* s, double* sv, double* S) {
  double DS[144], Dsv[12];
  double extra;
  if (theta == 0.0) {  // ileqg.jl:384-385 (D = I)
    for (int i = 0; i < n * n; ++i) DS[i] = Sp[i];
    for (int i = 0; i < n; ++i) Dsv[i] = svp[i];
    double tr = 0.0;
    for (int i = 0; i < n; ++i) {
      double t = dotp(n, W + i, n, Sp + i * n, 1);
      tr = (i == 0) ? t : tr + t;
    }
    extra = 0.5 * tr;
  } else {
    double M[144], C[144], invd[12], Z[144], z[12];
    for (int i = 0; i < n * n; ++i) M[i] = Winv[i] - theta * Sp[i];  // :365
    double detM;
    if (!chol_lower(n, M, C, invd, &detM)) throw NotPosDef();  // :366
    // Z = C^-1 Sp, z = C^-1 svp  =>  Sp M^-1 Sp = Z'Z ;  D*Sp = Sp + theta Z'Z  (:367)
    for (int c = 0; c < n; ++c)
      for (int i = 0; i < n; ++i) {
        double a = Sp[i + c * n];
        for (int k = 0; k < i; ++k) a = FMA(-C[i + k * n], Z[k + c * n], a);
        Z[i + c * n] = a * invd[i];
      }
    for (int i = 0; i < n; ++i) {
      double a = svp[i];
      for (int k = 0; k < i; ++k) a = FMA(-C[i + k * n], z[k], a);
      z[i] = a * invd[i];
    }
    for (int i = 0; i < n; ++i)
      for (int j = i; j < n; ++j) {
        double e = dotp(n, &Z[i * n], 1, &Z[j * n], 1);
        double v = FMA(theta, e, Sp[i + j * n]);
        DS[i + j * n] = v;
        DS[j + i * n] = v;
      }
    for (int i = 0; i < n; ++i) Dsv[i] = FMA(theta, dotp(n, &Z[i * n], 1, z, 1), svp[i]);
    double quad = dotp(n, z, 1, z, 1);
    extra = (theta / 2) * quad - (1 / (2 * theta)) * std::log(detW * detM);  // :387
  }
  double T[144], U[48], g[4], G[48], H[16];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) T[i + j * n] = dotp(n, &DS[i], n, A + j * n, 1);
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < n; ++i) U[i + j * n] = dotp(n, &DS[i], n, B + j * n, 1);
  for (int i = 0; i < m; ++i) g[i] = r[i] + dotp(n, B + i * n, 1, Dsv, 1);  // :368
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) G[i + j * m] = P[i + j * m] + dotp(n, B + i * n, 1, &T[j * n], 1);  // :369
  for (int i = 0; i < m; ++i)
    for (int j = i; j < m; ++j) {  // :370-371 (Symmetric: upper triangle mirrored)
      double h = R[i + j * m] + dotp(n, B + i * n, 1, &U[j * n], 1);
      if (i == j) h = h + mu;
      H[i + j * m] = h;
      H[j + i * m] = h;
    }
  if (optimise) {
    double CH[16], invh[4];
    if (!chol_lower(m, H, CH, invh, nullptr)) return false;  // :372
    // L = -H\G ; dl = -H\g  (:379-382)
    for (int c = 0; c <= n; ++c) {
      double y[4];
      const double* rhs = (c < n) ? &G[c * m] : g;
      for (int i = 0; i < m; ++i) {
        double a = rhs[i];
        for (int k = 0; k < i; ++k) a = FMA(-CH[i + k * m], y[k], a);
        y[i] = a * invh[i];
      }
      for (int i = m - 1; i >= 0; --i) {
        double a = y[i];
        for (int k = i + 1; k < m; ++k) a = FMA(-CH[k + i * m], y[k], a);
        y[i] = a * invh[i];
      }
      for (int i = 0; i < m; ++i) {
        if (c < n) L[i + c * m] = -y[i]; else dl[i] = -y[i];
      }
    }
  }
  double HL[48], Hdl[4];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) HL[i + j * m] = dotp(m, &H[i], m, L + j * m, 1);
  double sval = q + sp;  // :383 / :452
  if (dl) {
    for (int i = 0; i < m; ++i) Hdl[i] = dotp(m, &H[i], m, dl, 1);
    sval = (sval + 0.5 * dotp(m, dl, 1, Hdl, 1)) + dotp(m, dl, 1, g, 1);
  }
  *s = sval + extra;
  for (int i = 0; i < n; ++i) {  // :389 / :458
    double a = qv[i] + dotp(n, A + i * n, 1, Dsv, 1);
    if (dl) a = a + dotp(m, L + i * m, 1, Hdl, 1);
    a = a + dotp(m, L + i * m, 1, g, 1);
    if (dl) a = a + dotp(m, &G[i * m], 1, dl, 1);
    sv[i] = a;
  }
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {  // :390-391 / :459-460
      double a = Q[i + j * n] + dotp(n, A + i * n, 1, &T[j * n], 1);
      a = a + dotp(m, L + i * m, 1, &HL[j * m], 1);
      a = a + dotp(m, L + i * m, 1, &G[j * m], 1);
      a = a + dotp(m, &G[i * m], 1, L + j * m, 1);
      S[i + j * n] = a;
      S[j + i * n] = a;
    }
  return true;
}

struct DP { vec s, sv, S; };

void dp_terminal(const Approx& ap, DP& dp) {  // ileqg.jl:352-354 / :429-431
  int n = ap.n, N = ap.N;
  dp.s.assign(N + 1, 0.0); dp.sv.assign((size_t)n * (N + 1), 0.0); dp.S.assign((size_t)n * n * (N + 1), 0.0);
  dp.s[N] = ap.q[N];
  for (int i = 0; i < n; ++i) dp.sv[(size_t)N * n + i] = ap.qv[(size_t)N * n + i];
  const double* Qn = &ap.Q[(size_t)N * n * n];
  double* Sn = &dp.S[(size_t)N * n * n];
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) { Sn[i + j * n] = Qn[i + j * n]; Sn[j + i * n] = Qn[i + j * n]; }
}

struct WSet {  // W(k) for k = 0..N-1, pre-processed
  std::vector<WInfo> w; int tv; const double* W; int n;
  const WInfo& at(int k) const { return w[tv ? k : 0]; }
  const double* Wk(int k) const { return W + (tv ? (size_t)k * n * n : 0); }
};
bool prep_Wset(int n, int N, const double* W, int tv, WSet& ws) {
  ws.tv = tv; ws.W = W; ws.n = n;
  ws.w.resize(tv ? N : 1);
  for (size_t k = 0; k < ws.w.size(); ++k)
    if (!prep_W(n, W + k * n * n, ws.w[k])) return false;
  return true;
}

// solve_approximate_dp  ileqg.jl:412-465 (evaluate a given policy)
void solve_dp_eval(const Approx& ap, const WSet& ws, const double* L, const double* dl,
                   double theta, double mu, DP& dp) {
  int n = ap.n, m = ap.m, N = ap.N;
  dp_terminal(ap, dp);
  for (int ii = N - 1; ii >= 0; --ii) {
    const WInfo& wi = ws.at(ii);
    riccati_stage(n, m, false, theta, mu, ws.Wk(ii), wi.Winv.data(), wi.detW,
                  &dp.S[(size_t)(ii + 1) * n * n], &dp.sv[(size_t)(ii + 1) * n], dp.s[ii + 1],
                  ap.q[ii], &ap.qv[(size_t)ii * n], &ap.Q[(size_t)ii * n * n], &ap.r[(size_t)ii * m],
                  &ap.R[(size_t)ii * m * m], &ap.P[(size_t)ii * m * n], &ap.A[(size_t)ii * n * n],
                  &ap.B[(size_t)ii * n * m],
                  const_cast<double*>(L + (size_t)ii * m * n),
                  dl ? const_cast<double*>(dl + (size_t)ii * m) : nullptr,
                  &dp.s[ii], &dp.sv[(size_t)ii * n], &dp.S[(size_t)ii * n * n]);
  }
}

struct Solver {  // ILEQGSolver  ileqg.jl:164-189
  double mu_min, mu, delta_0, delta, lambda, d;
  int iter_max; bool eps_auto; double eps_init, eps_min, eps_init_init;
  vec x, l, L;
  double value; int iter; double d_current;
  std::vector<double> eps_hist;  // flattened pairs
  int trials, restarts;
};

void increase_mu_delta(Solver& s) {  // ileqg.jl:471-474
  s.delta = std::max(s.delta_0, s.delta * s.delta_0);
  s.mu = std::max(s.mu_min, s.mu * s.delta);
  s.restarts++;
}

// solve_approximate_dp!  ileqg.jl:341-406
void solve_dp_opt(Solver& sv_, const Approx& ap, const WSet& ws, double theta, DP& dp, vec& dl) {
  int n = ap.n, m = ap.m, N = ap.N;
  dp_terminal(ap, dp);
  dl.assign((size_t)m * N, 0.0);
  bool all_pd = false;
  while (!all_pd) {
    if (!(sv_.mu < 1e300)) throw NotPosDef();  // guard: reference would spin forever with mu = Inf/NaN
    for (int ii = N - 1; ii >= 0; --ii) {
      const WInfo& wi = ws.at(ii);
      bool ok = riccati_stage(n, m, true, theta, sv_.mu, ws.Wk(ii), wi.Winv.data(), wi.detW,
                              &dp.S[(size_t)(ii + 1) * n * n], &dp.sv[(size_t)(ii + 1) * n], dp.s[ii + 1],
                              ap.q[ii], &ap.qv[(size_t)ii * n], &ap.Q[(size_t)ii * n * n], &ap.r[(size_t)ii * m],
                              &ap.R[(size_t)ii * m * m], &ap.P[(size_t)ii * m * n], &ap.A[(size_t)ii * n * n],
                              &ap.B[(size_t)ii * n * m],
                              &sv_.L[(size_t)ii * m * n], &dl[(size_t)ii * m],
                              &dp.s[ii], &dp.sv[(size_t)ii * n], &dp.S[(size_t)ii * n * n]);
      if (!ok) { increase_mu_delta(sv_); break; }  // :372-378
      if (ii == 0) all_pd = true;
    }
  }
}

// simulate_dynamics open loop  ileqg.jl:18-38
void rollout_open(const Problem& pr, const double* x0, const double* u, double* x) {
  int n = pr.n, m = pr.m;
  for (int i = 0; i < n; ++i) x[i] = x0[i];
  for (int ii = 0; ii < pr.N; ++ii) f_eval(pr, x + (size_t)ii * n, u + (size_t)ii * m, x + (size_t)(ii + 1) * n);
}

// simulate_dynamics closed loop  ileqg.jl:62-87 (+ optional additive noise w: :94-109)
void rollout_closed(const Problem& pr, const double* xbar, const double* l, const double* L,
                    double* xn, double* un, const double* noise) {
  int n = pr.n, m = pr.m;
  for (int i = 0; i < n; ++i) xn[i] = xbar[i];
  for (int ii = 0; ii < pr.N; ++ii) {
    double dx[12];
    const double* Lk = L + (size_t)ii * m * n;
    for (int i = 0; i < n; ++i) dx[i] = xn[(size_t)ii * n + i] - xbar[(size_t)ii * n + i];
    for (int j = 0; j < m; ++j) un[(size_t)ii * m + j] = l[(size_t)ii * m + j] + dotp(n, Lk + j, m, dx, 1);
    f_eval(pr, xn + (size_t)ii * n, un + (size_t)ii * m, xn + (size_t)(ii + 1) * n);
    if (noise)
      for (int i = 0; i < n; ++i) xn[(size_t)(ii + 1) * n + i] += noise[(size_t)ii * n + i];
  }
}

// integrate_cost  ileqg.jl:115-124
double integrate_cost(const Problem& pr, const double* x, const double* u) {
  double cost = 0.0;
  for (int ii = 0; ii < pr.N; ++ii)
    cost += cost_stage(pr, ii, x + (size_t)ii * pr.n, u + (size_t)ii * pr.m, nullptr, nullptr, nullptr, nullptr, nullptr);
  cost += cost_terminal(pr, x + (size_t)pr.N * pr.n, nullptr, nullptr);
  return cost;
}

// initialize!  ileqg.jl:214-236
void ileqg_initialize(Solver& s, const Problem& pr, const WSet& ws, const double* x0, const double* u, double theta) {
  int n = pr.n, m = pr.m, N = pr.N;
  s.mu = 0.0; s.delta = s.delta_0; s.d_current = INF; s.iter = 0;
  s.eps_init = s.eps_init_init; s.eps_hist.clear(); s.trials = 0; s.restarts = 0;
  s.x.assign((size_t)n * (N + 1), 0.0);
  rollout_open(pr, x0, u, s.x.data());
  s.l.assign(u, u + (size_t)m * N);
  s.L.assign((size_t)m * n * N, 0.0);
  Approx ap; DP dp;
  approximate_model(pr, s.l.data(), s.x.data(), ap);
  solve_dp_eval(ap, ws, s.L.data(), nullptr, theta, s.mu, dp);
  s.value = dp.s[0];
}

inline bool isapprox_default(double a, double b) {  // Base.isapprox, rtol = sqrt(eps), atol = 0
  if (a == b) return true;
  if (!std::isfinite(a) || !std::isfinite(b)) return false;
  const double rtol = 1.4901161193847656e-8;
  return std::fabs(a - b) <= rtol * std::max(std::fabs(a), std::fabs(b));
}

double max_norm_diff(int m, int N, const double* l, const double* u) {  // maximum(norm.(l .- u))  ileqg.jl:539
  double best = -INF;
  bool nan = false;
  for (int k = 0; k < N; ++k) {
    double acc = 0.0;
    for (int j = 0; j < m; ++j) {
      double dd = l[(size_t)k * m + j] - u[(size_t)k * m + j];
      acc = (j == 0) ? dd * dd : FMA(dd, dd, acc);
    }
    double nr = std::sqrt(acc);
    if (std::isnan(nr)) nan = true;
    if (nr > best) best = nr;
  }
  return nan ? std::numeric_limits<double>::quiet_NaN() : best;
}

struct LineSearchHang {};

// line_search!  ileqg.jl:494-592
void line_search(Solver& s, const Problem& pr, const WSet& ws, const vec& dl, double theta) {
  int n = pr.n, m = pr.m, N = pr.N;
  double cur = s.value;
  double eps = s.eps_init;
  int count = 0;
  vec lnew((size_t)m * N), xn((size_t)n * (N + 1)), un((size_t)m * N);
  Approx ap; DP dp;
  while (true) {
    count++;
    if (eps == 0.0 || count > 4000) throw LineSearchHang();  // reference never exits here (:526-535)
    for (size_t i = 0; i < lnew.size(); ++i) lnew[i] = s.l[i] + eps * dl[i];  // :509
    rollout_closed(pr, s.x.data(), lnew.data(), s.L.data(), xn.data(), un.data(), nullptr);  // :517-519
    approximate_model(pr, un.data(), xn.data(), ap);                                          // :520
    bool ok = true;
    try { solve_dp_eval(ap, ws, s.L.data(), nullptr, theta, s.mu, dp); } catch (NotPosDef&) { ok = false; }  // :522-528
    if (!ok) { eps *= s.lambda; continue; }  // :529-535
    double nw = dp.s[0];
    s.eps_hist.push_back(eps); s.eps_hist.push_back(nw - cur); s.trials++;  // :537
    if (isapprox_default(nw, cur) || nw < cur) {  // :538
      s.d_current = max_norm_diff(m, N, s.l.data(), un.data());
      s.value = nw; s.x = xn; s.l = un;
      break;
    } else {
      eps *= s.lambda;
      if (eps < s.eps_min) {  // :558-575
        s.d_current = max_norm_diff(m, N, s.l.data(), un.data());
        s.value = nw; s.x = xn; s.l = un;
        break;
      }
    }
  }
  if (s.eps_auto) {  // :582-591
    if (count == 1) s.eps_init = std::min(s.eps_init_init, eps / s.lambda);
    else { while (eps < s.eps_min) eps = eps / s.lambda; s.eps_init = eps; }
  }
}

// step!  ileqg.jl:598-613
void ileqg_step(Solver& s, const Problem& pr, const WSet& ws, double theta) {
  s.iter++;
  Approx ap; DP dp; vec dl;
  approximate_model(pr, s.l.data(), s.x.data(), ap);
  solve_dp_opt(s, ap, ws, theta, dp, dl);
  line_search(s, pr, ws, dl, theta);
}

Solver make_solver(const ratilqr_ileqg_opts& o) {
  Solver s;
  s.mu_min = o.mu_min; s.mu = o.mu_min; s.delta_0 = o.delta_0; s.delta = o.delta_0; s.lambda = o.lambda;
  s.d = o.d; s.iter_max = o.iter_max; s.eps_auto = o.adaptive_eps_init != 0; s.eps_init = o.eps_init;
  s.eps_min = o.eps_min; s.eps_init_init = o.eps_init; s.value = INF; s.iter = 0; s.d_current = INF;
  s.trials = 0; s.restarts = 0;
  return s;
}

// solve!  ileqg.jl:635-659 ; returns status, exceptions mapped exactly as the bilevel workers see them
int ileqg_solve(Solver& s, const Problem& pr, const WSet& ws, const double* x0, const double* u, double theta) {
  int stage = 0;
  try {
    ileqg_initialize(s, pr, ws, x0, u, theta);
    stage = 1;
    while (true) {
      ileqg_step(s, pr, ws, theta);
      if (s.d > s.d_current && s.mu <= s.mu_min) break;  // :642
      else if (s.iter == s.iter_max) break;              // :648
    }
  } catch (NotPosDef&) {
    s.value = INF;
    if (stage == 1 && !(s.mu < 1e300)) return RATILQR_ST_MU_OVERFLOW;
    return stage == 0 ? RATILQR_ST_M_NOT_PD_INIT : RATILQR_ST_M_NOT_PD_OPT;
  } catch (DomainError&) {
    s.value = INF; return RATILQR_ST_DOMAIN;
  } catch (LineSearchHang&) {
    s.value = INF; return RATILQR_ST_LINESEARCH_HANG;
  }
  return RATILQR_ST_OK;
}

Problem make_problem(const ratilqr_problem_desc* d, int p) {
  Problem pr;
  pr.model_id = d->model_id; pr.cost_id = d->cost_id; pr.n = d->n; pr.m = d->m; pr.N = d->N;
  pr.mp = d->model_params;
  pr.cp = d->cost_params + (d->cost_params_count > 1 ? (size_t)p * d->n_cost_params : 0);
  pr.W = d->W; pr.W_tv = d->W_time_varying;
  return pr;
}

int g_threads = 0;  // 0 = hardware_concurrency

template <class F>
void parallel_for(int count, F fn) {
  int nt = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > count) nt = count;
  if (nt <= 1) { for (int i = 0; i < count; ++i) fn(i); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([=]() {  // static block partition (the analogue of the round-robin over worker processes,
      int lo = (int)((int64_t)count * t / nt), hi = (int)((int64_t)count * (t + 1) / nt);  // cross_entropy...:180-192)
      for (int i = lo; i < hi; ++i) fn(i);
    });
  for (auto& t : th) t.join();
}

}  // namespace

// =========================================================================================
// C ABI mirroring include/ratilqr.h with the prefix oracle_ (ctx argument ignored)
// =========================================================================================
extern "C" {

int32_t oracle_set_threads(int32_t n) { g_threads = n; return 0; }
int32_t oracle_get_threads(void) { return g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency(); }

int32_t oracle_model_dims(int32_t id, int32_t* n, int32_t* m, int32_t* np) { return model_dims(id, n, m, np) ? 0 : -1; }

int32_t oracle_ileqg_solve_batch(void*, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                                 const ratilqr_batch_in* in, ratilqr_ileqg_out* out) {
  int n = desc->n, m = desc->m, N = desc->N, P = in->P, K = in->K;
  WSet ws;
  if (!prep_Wset(n, N, desc->W, desc->W_time_varying, ws)) return -2;
  parallel_for(P * K, [&](int b) {
    int p = b / K;
    Problem pr = make_problem(desc, p);
    Solver s = make_solver(*opts);
    const double* x0 = in->x0 + (in->x0_count > 1 ? (size_t)p * n : 0);
    const double* u = in->u_init + (in->u_count > 1 ? (size_t)p * m * N : 0);
    int st = ileqg_solve(s, pr, ws, x0, u, in->theta[b]);
    if (out->status) out->status[b] = st;
    if (out->value) out->value[b] = s.value;
    if (out->iters) out->iters[b] = s.iter;
    if (out->trials) out->trials[b] = s.trials;
    if (out->restarts) out->restarts[b] = s.restarts;
    if (out->mu) out->mu[b] = s.mu;
    if (out->d_current) out->d_current[b] = s.d_current;
    if (out->x && s.x.size()) std::memcpy(out->x + (size_t)b * n * (N + 1), s.x.data(), sizeof(double) * n * (N + 1));
    if (out->l && s.l.size()) std::memcpy(out->l + (size_t)b * m * N, s.l.data(), sizeof(double) * m * N);
    if (out->L && s.L.size()) std::memcpy(out->L + (size_t)b * m * n * N, s.L.data(), sizeof(double) * m * n * N);
    if (out->eps_hist) {
      double* h = out->eps_hist + (size_t)b * 2 * out->eps_hist_cap;
      for (int i = 0; i < 2 * out->eps_hist_cap; ++i) h[i] = 0.0;
      size_t cnt = std::min(s.eps_hist.size(), (size_t)2 * out->eps_hist_cap);
      for (size_t i = 0; i < cnt; ++i) h[i] = s.eps_hist[i];
    }
  });
  return 0;
}

// compute_cost / compute_cost_serial  cross_entropy_bilevel_optimization.jl:173-227
int32_t oracle_ce_costs(void*, const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* opts,
                        const ratilqr_batch_in* in, double kl_bound, double* cost, int32_t* status) {
  int B = in->P * in->K;
  vec value(B);
  std::vector<int32_t> st(B);
  ratilqr_ileqg_out out;
  std::memset(&out, 0, sizeof(out));
  out.value = value.data(); out.status = st.data();
  int rc = oracle_ileqg_solve_batch(nullptr, desc, opts, in, &out);
  if (rc) return rc;
  for (int b = 0; b < B; ++b) {
    cost[b] = st[b] == 0 ? value[b] + kl_bound / in->theta[b] : INF;  // :193 / :221-224
    if (status) status[b] = st[b];
  }
  return 0;
}

int32_t oracle_rollout_open_batch(void*, const ratilqr_problem_desc* desc, int32_t B, const double* x0,
                                  const double* u, double* x, int32_t* status) {
  int n = desc->n, m = desc->m, N = desc->N;
  for (int b = 0; b < B; ++b) {
    Problem pr = make_problem(desc, 0);
    int st = 0;
    try { rollout_open(pr, x0 + (size_t)b * n, u + (size_t)b * m * N, x + (size_t)b * n * (N + 1)); }
    catch (DomainError&) { st = RATILQR_ST_DOMAIN; }
    if (status) status[b] = st;
  }
  return 0;
}

int32_t oracle_rollout_closed_batch(void*, const ratilqr_problem_desc* desc, int32_t B, const double* xbar,
                                    const double* l, const double* L, double* x_new, double* u_new, int32_t* status) {
  int n = desc->n, m = desc->m, N = desc->N;
  for (int b = 0; b < B; ++b) {
    Problem pr = make_problem(desc, 0);
    int st = 0;
    try {
      rollout_closed(pr, xbar + (size_t)b * n * (N + 1), l + (size_t)b * m * N, L + (size_t)b * m * n * N,
                     x_new + (size_t)b * n * (N + 1), u_new + (size_t)b * m * N, nullptr);
    } catch (DomainError&) { st = RATILQR_ST_DOMAIN; }
    if (status) status[b] = st;
  }
  return 0;
}

int32_t oracle_integrate_cost_batch(void*, const ratilqr_problem_desc* desc, int32_t B, const double* x,
                                    const double* u, double* cost, int32_t* status) {
  int n = desc->n, m = desc->m, N = desc->N;
  for (int b = 0; b < B; ++b) {
    Problem pr = make_problem(desc, 0);
    int st = 0;
    try { cost[b] = integrate_cost(pr, x + (size_t)b * n * (N + 1), u + (size_t)b * m * N); }
    catch (DomainError&) { st = RATILQR_ST_DOMAIN; cost[b] = INF; }
    if (status) status[b] = st;
  }
  return 0;
}

int32_t oracle_linearize_batch(void*, const ratilqr_problem_desc* desc, int32_t B, const double* x, const double* u,
                               double* q, double* qv, double* Q, double* r, double* R, double* Pm, double* A,
                               double* Bm, int32_t* status) {
  int n = desc->n, m = desc->m, N = desc->N;
  for (int b = 0; b < B; ++b) {
    Problem pr = make_problem(desc, 0);
    Approx ap;
    int st = 0;
    try { approximate_model(pr, u + (size_t)b * m * N, x + (size_t)b * n * (N + 1), ap); }
    catch (DomainError&) { st = RATILQR_ST_DOMAIN; }
    if (status) status[b] = st;
    if (st) continue;
    std::memcpy(q + (size_t)b * (N + 1), ap.q.data(), sizeof(double) * (N + 1));
    std::memcpy(qv + (size_t)b * n * (N + 1), ap.qv.data(), sizeof(double) * n * (N + 1));
    std::memcpy(Q + (size_t)b * n * n * (N + 1), ap.Q.data(), sizeof(double) * n * n * (N + 1));
    std::memcpy(r + (size_t)b * m * N, ap.r.data(), sizeof(double) * m * N);
    std::memcpy(R + (size_t)b * m * m * N, ap.R.data(), sizeof(double) * m * m * N);
    std::memcpy(Pm + (size_t)b * m * n * N, ap.P.data(), sizeof(double) * m * n * N);
    std::memcpy(A + (size_t)b * n * n * N, ap.A.data(), sizeof(double) * n * n * N);
    std::memcpy(Bm + (size_t)b * n * m * N, ap.B.data(), sizeof(double) * n * m * N);
  }
  return 0;
}

int32_t oracle_riccati_batch_tv(void*, int32_t n, int32_t m, int32_t N, int32_t B, int32_t optimise,
                                const double* q, const double* qv, const double* Q, const double* r, const double* R,
                                const double* Pm, const double* A, const double* Bm, const double* W,
                                int32_t W_time_varying, const double* theta, double mu_min, double delta_0, double* mu,
                                double* delta, double* L, double* dl, double* s, double* sv, double* S, int32_t* status,
                                int32_t* restarts);
int32_t oracle_riccati_batch(void* c, int32_t n, int32_t m, int32_t N, int32_t B, int32_t optimise,
                             const double* q, const double* qv, const double* Q, const double* r, const double* R,
                             const double* Pm, const double* A, const double* Bm, const double* W,
                             const double* theta, double mu_min, double delta_0, double* mu, double* delta,
                             double* L, double* dl, double* s, double* sv, double* S, int32_t* status,
                             int32_t* restarts) {
  return oracle_riccati_batch_tv(c, n, m, N, B, optimise, q, qv, Q, r, R, Pm, A, Bm, W, 0, theta, mu_min, delta_0, mu, delta,
                                 L, dl, s, sv, S, status, restarts);
}
int32_t oracle_riccati_batch_tv(void*, int32_t n, int32_t m, int32_t N, int32_t B, int32_t optimise,
                                const double* q, const double* qv, const double* Q, const double* r, const double* R,
                                const double* Pm, const double* A, const double* Bm, const double* W,
                                int32_t W_time_varying, const double* theta, double mu_min, double delta_0, double* mu,
                                double* delta, double* L, double* dl, double* s, double* sv, double* S, int32_t* status,
                                int32_t* restarts) {
  WSet ws;
  if (!prep_Wset(n, N, W, W_time_varying, ws)) return -2;
  for (int b = 0; b < B; ++b) {
    Approx ap;
    ap.resize(n, m, N);
    std::memcpy(ap.q.data(), q + (size_t)b * (N + 1), sizeof(double) * (N + 1));
    std::memcpy(ap.qv.data(), qv + (size_t)b * n * (N + 1), sizeof(double) * n * (N + 1));
    std::memcpy(ap.Q.data(), Q + (size_t)b * n * n * (N + 1), sizeof(double) * n * n * (N + 1));
    std::memcpy(ap.r.data(), r + (size_t)b * m * N, sizeof(double) * m * N);
    std::memcpy(ap.R.data(), R + (size_t)b * m * m * N, sizeof(double) * m * m * N);
    std::memcpy(ap.P.data(), Pm + (size_t)b * m * n * N, sizeof(double) * m * n * N);
    std::memcpy(ap.A.data(), A + (size_t)b * n * n * N, sizeof(double) * n * n * N);
    std::memcpy(ap.B.data(), Bm + (size_t)b * n * m * N, sizeof(double) * n * m * N);
    DP dp;
    int st = 0;
    Solver sol;
    sol.mu_min = mu_min; sol.delta_0 = delta_0; sol.mu = mu[b]; sol.delta = delta[b]; sol.restarts = 0;
    try {
      if (optimise) {
        sol.L.assign((size_t)m * n * N, 0.0);
        vec dlv;
        solve_dp_opt(sol, ap, ws, theta[b], dp, dlv);
        std::memcpy(L + (size_t)b * m * n * N, sol.L.data(), sizeof(double) * m * n * N);
        std::memcpy(dl + (size_t)b * m * N, dlv.data(), sizeof(double) * m * N);
        mu[b] = sol.mu; delta[b] = sol.delta;
      } else {
        solve_dp_eval(ap, ws, L + (size_t)b * m * n * N, dl ? dl + (size_t)b * m * N : nullptr, theta[b], mu[b], dp);
      }
    } catch (NotPosDef&) { st = optimise ? RATILQR_ST_M_NOT_PD_OPT : RATILQR_ST_M_NOT_PD_INIT; }
    if (status) status[b] = st;
    if (restarts) restarts[b] = sol.restarts;
    if (st) continue;
    std::memcpy(s + (size_t)b * (N + 1), dp.s.data(), sizeof(double) * (N + 1));
    std::memcpy(sv + (size_t)b * n * (N + 1), dp.sv.data(), sizeof(double) * n * (N + 1));
    std::memcpy(S + (size_t)b * n * n * (N + 1), dp.S.data(), sizeof(double) * n * n * (N + 1));
  }
  return 0;
}

// noisy closed-loop rollouts + cost (ileqg.jl:94-109 + :115-124); noise is an injected w tensor
int32_t oracle_mc_rollout(void*, const ratilqr_problem_desc* desc, int32_t P, const double* xbar, const double* l,
                          const double* L, int32_t n_samples, const double* noise, uint64_t /*seed*/,
                          double theta_risk, double* J, double* stats, double* x_out) {
  int n = desc->n, m = desc->m, N = desc->N;
  if (!noise) return -3;  // the oracle only supports injected noise
  for (int p = 0; p < P; ++p) {
    Problem pr = make_problem(desc, p);
    const double* xb = xbar + (size_t)p * n * (N + 1);
    const double* lp = l + (size_t)p * m * N;
    const double* Lp = L + (size_t)p * m * n * N;
    parallel_for(n_samples, [&](int sidx) {
      size_t gi = (size_t)p * n_samples + sidx;
      vec xn((size_t)n * (N + 1)), un((size_t)m * N);
      double cost;
      try {
        rollout_closed(pr, xb, lp, Lp, xn.data(), un.data(), noise + gi * n * N);
        cost = integrate_cost(pr, xn.data(), un.data());
      } catch (DomainError&) { cost = INF; }
      J[gi] = cost;
      if (x_out) std::memcpy(x_out + gi * n * (N + 1), xn.data(), sizeof(double) * n * (N + 1));
    });
    if (stats) {
      const double* Jp = J + (size_t)p * n_samples;
      double mean = 0.0;
      for (int i = 0; i < n_samples; ++i) mean += Jp[i];
      mean /= n_samples;
      double var = 0.0;
      for (int i = 0; i < n_samples; ++i) var += (Jp[i] - mean) * (Jp[i] - mean);
      var = n_samples > 1 ? var / (n_samples - 1) : 0.0;
      double risk = mean;
      if (theta_risk > 0.0) {  // 1/theta log mean exp(theta J), log-sum-exp form
        double mx = -INF;
        for (int i = 0; i < n_samples; ++i) mx = std::max(mx, theta_risk * Jp[i]);
        double acc = 0.0;
        for (int i = 0; i < n_samples; ++i) acc += std::exp(theta_risk * Jp[i] - mx);
        risk = (mx + std::log(acc / n_samples)) / theta_risk;
      }
      stats[3 * p + 0] = mean; stats[3 * p + 1] = var; stats[3 * p + 2] = risk;
    }
  }
  return 0;
}

// compute_cost_serial  pets.jl:128-157 with injected noise (additive: x+ = f(x,u) + w)
int32_t oracle_pets_costs(void*, const ratilqr_problem_desc* desc, const ratilqr_generative_desc* gen,
                          const double* x0, const double* controls, int32_t C, int32_t particles,
                          const double* noise, uint64_t /*seed*/, double* cost) {
  int n = desc->n, m = desc->m, N = desc->N;
  if (!noise) return -3;
  int ne = gen && gen->n_ensemble > 1 ? gen->n_ensemble : 1;
  int per = particles / ne;
  if (per < 1) per = 1;
  parallel_for(C, [&](int ii) {
    const double* useq = controls + (size_t)ii * m * N;
    double total = 0.0;
    for (int kk = 0; kk < particles; ++kk) {
      Problem pr = make_problem(desc, 0);
      if (gen && gen->ensemble_params && ne > 1) pr.mp = gen->ensemble_params + (size_t)std::min(kk / per, ne - 1) * desc->n_model_params;
      const double* w = noise + ((size_t)ii * particles + kk) * n * N;
      double x[12], xn[12];
      for (int i = 0; i < n; ++i) x[i] = x0[i];
      double c = 0.0;
      try {
        for (int tt = 0; tt < N; ++tt) {
          c += cost_stage(pr, tt, x, useq + (size_t)tt * m, nullptr, nullptr, nullptr, nullptr, nullptr);  // :147
          f_eval(pr, x, useq + (size_t)tt * m, xn);                                                        // :148
          for (int i = 0; i < n; ++i) x[i] = xn[i] + w[(size_t)tt * n + i];
        }
        c += cost_terminal(pr, x, nullptr, nullptr);  // :151
      } catch (DomainError&) { c = INF; }
      total += c;
    }
    cost[ii] = total / particles;  // :154
  });
  return 0;
}

// get_elite_samples + compute_new_distribution  pets.jl:159-191
int32_t oracle_pets_refit(void*, int32_t m, int32_t N, int32_t C, int32_t num_elite, double smoothing,
                          const double* controls, const double* cost, double* mu, double* Sigma, int32_t* elite_idx) {
  std::vector<int> idx(C);
  for (int i = 0; i < C; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {  // isless: NaN last
    double ca = cost[a], cb = cost[b];
    if (std::isnan(ca)) return false;
    if (std::isnan(cb)) return true;
    return ca < cb;
  });
  if (elite_idx) for (int e = 0; e < num_elite; ++e) elite_idx[e] = idx[e];
  for (int tt = 0; tt < N; ++tt)
    for (int j = 0; j < m; ++j) {
      double sum = 0.0;
      for (int e = 0; e < num_elite; ++e) sum += controls[(size_t)idx[e] * m * N + (size_t)tt * m + j];
      double mean = sum / num_elite;  // :184
      double ss = 0.0;
      for (int e = 0; e < num_elite; ++e) {
        double dd = controls[(size_t)idx[e] * m * N + (size_t)tt * m + j] - mean;
        ss += dd * dd;
      }
      double var = ss / (num_elite - 1);  // Julia var: unbiased  :185
      double* mu_t = mu + (size_t)tt * m;
      double* Sg = Sigma + (size_t)tt * m * m;
      mu_t[j] = (1.0 - smoothing) * mean + smoothing * mu_t[j];  // :187
      for (int i = 0; i < m; ++i) {                              // :188 (column j of Sigma_t)
        double cv = (i == j) ? var : 0.0;
        Sg[i + j * m] = (1.0 - smoothing) * cv + smoothing * Sg[i + j * m];
      }
    }
  return 0;
}

// step!/solve!  pets.jl:193-245,270-281 with injected standard normals z (m*N*C*iter_max) and noise
// (n*N*particles*C*iter_max): u = mu_t + chol_lower(Sigma_t) z   (MvNormal sampling, :212-213)
int32_t oracle_pets_solve(void*, const ratilqr_problem_desc* desc, const ratilqr_generative_desc* gen,
                          const double* x0, int32_t C, int32_t particles, int32_t num_elite, int32_t iter_max,
                          double smoothing, const double* z_inject, const double* noise, uint64_t seed,
                          double* mu, double* Sigma) {
  int n = desc->n, m = desc->m, N = desc->N;
  if (!z_inject || !noise) return -3;
  vec controls((size_t)m * N * C), cost(C), Cs(m * m), invd(m);
  for (int it = 0; it < iter_max; ++it) {
    const double* z = z_inject + (size_t)it * m * N * C;
    for (int tt = 0; tt < N; ++tt) {
      if (!chol_lower(m, Sigma + (size_t)tt * m * m, Cs.data(), invd.data(), nullptr)) return -4;  // PosDefException
      for (int ii = 0; ii < C; ++ii)
        for (int j = 0; j < m; ++j) {
          double a = mu[(size_t)tt * m + j];
          for (int k = 0; k <= j; ++k) a = FMA(Cs[j + k * m], z[(size_t)ii * m * N + (size_t)tt * m + k], a);
          controls[(size_t)ii * m * N + (size_t)tt * m + j] = a;
        }
    }
    int rc = oracle_pets_costs(nullptr, desc, gen, x0, controls.data(), C, particles,
                               noise + (size_t)it * n * N * particles * C, seed, cost.data());
    if (rc) return rc;
    oracle_pets_refit(nullptr, m, N, C, num_elite, smoothing, controls.data(), cost.data(), mu, Sigma, nullptr);
  }
  return 0;
}

// ---- bilevel optimisers (host logic of the reference, restated for whole-solve parity) ----

typedef struct {
  double mu_init, sigma_init; int32_t num_samples, num_elite, iter_max; double lambda; int32_t use_theta_max;
} oracle_ce_opts;  // cross_entropy_bilevel_optimization.jl:100-116

// solve!(ce_solver, ...)  cross_entropy_bilevel_optimization.jl:364-415 for ONE problem, with the
// rng replaced by an injected stream of standard normals z (rand(rng, Normal(mu,sigma)) = mu + sigma*z).
// returns: theta_opt, value(+kl/theta), theta_min, theta_max, mu/sigma (final), mu_init/sigma_init (in-out),
// n_z_used, and x,l,L of the final solve.
int32_t oracle_ce_solve(const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* iopts, oracle_ce_opts* ce,
                        const double* x0, const double* u_init, double kl_bound, const double* z, int64_t nz,
                        double* theta_opt_out, double* value_out, double* theta_min_out, double* theta_max_out,
                        double* mu_out, double* sigma_out, int64_t* nz_used, double* x, double* l, double* L,
                        int32_t* final_status) {
  int n = desc->n, m = desc->m, N = desc->N;
  WSet ws;
  if (!prep_Wset(n, N, desc->W, desc->W_time_varying, ws)) return -2;
  Problem pr = make_problem(desc, 0);
  double mu = ce->mu_init, sigma = ce->sigma_init;  // initialize! :133-138
  double th_max = 0.0, th_min = INF;
  int iter = 0;
  int64_t zi = 0;
  double theta_opt;
  int S = ce->num_samples;
  if (kl_bound > 0) {
    while (iter < ce->iter_max) {  // step! :252-335
      iter++;
      vec th(S), costs(S);
      while (true) {
        double mm = iter == 1 ? ce->mu_init : mu, ss = iter == 1 ? ce->sigma_init : sigma;
        int cnt = 0;
        while (cnt < S) {  // get_positive_samples :233-246
          if (zi >= nz) return -5;
          double t = mm + ss * z[zi++];
          if (t > 0.0) th[cnt++] = t;
        }
        ratilqr_batch_in in; in.P = 1; in.K = S; in.x0 = x0; in.x0_count = 1; in.u_init = u_init; in.u_count = 1; in.theta = th.data();
        oracle_ce_costs(nullptr, desc, iopts, &in, kl_bound, costs.data(), nullptr);
        int num_inf = 0;
        for (int i = 0; i < S; ++i) num_inf += std::isinf(costs[i]) ? 1 : 0;
        int num_valid = S - num_inf;
        double thr = std::max((double)ce->num_elite, S * ce->lambda);
        if (iter == 1 && num_valid < thr) { ce->mu_init *= ce->lambda; ce->sigma_init *= ce->lambda; }  // :293-298
        else if (iter == 1 && num_valid == S) { ce->mu_init /= ce->lambda; ce->sigma_init /= ce->lambda; break; }  // :299-305
        else if (num_valid >= thr) break;  // :306-310
      }
      for (int i = 0; i < S; ++i) {  // :314-324
        if (std::isinf(costs[i])) continue;
        if (th[i] < th_min) th_min = th[i];
        else if (th[i] > th_max) th_max = th[i];
      }
      std::vector<int> idx(S);
      for (int i = 0; i < S; ++i) idx[i] = i;
      std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
        if (std::isnan(costs[a])) return false;
        if (std::isnan(costs[b])) return true;
        return costs[a] < costs[b];
      });
      double sum = 0.0;
      for (int e = 0; e < ce->num_elite; ++e) sum += th[idx[e]];
      double mu_new = sum / ce->num_elite;  // :329
      double ss2 = 0.0;
      for (int e = 0; e < ce->num_elite; ++e) ss2 += (th[idx[e]] - mu_new) * (th[idx[e]] - mu_new);
      double sigma_new = std::sqrt(ss2 / ce->num_elite);  // :330
      mu = mu_new; sigma = sigma_new;
    }
    theta_opt = ce->use_theta_max ? th_max : mu;  // :375-382
  } else {
    theta_opt = 0.0;  // :388
  }
  int guard = 0;
  while (true) {  // :390-414
    Solver s = make_solver(*iopts);
    int st = ileqg_solve(s, pr, ws, x0, u_init, theta_opt);
    if (st == 0) {
      *theta_opt_out = theta_opt;
      *value_out = kl_bound > 0 ? s.value + kl_bound / theta_opt : s.value;
      *theta_min_out = kl_bound > 0 ? th_min : 0.0;
      *theta_max_out = kl_bound > 0 ? th_max : 0.0;
      if (x) std::memcpy(x, s.x.data(), sizeof(double) * n * (N + 1));
      if (l) std::memcpy(l, s.l.data(), sizeof(double) * m * N);
      if (L) std::memcpy(L, s.L.data(), sizeof(double) * m * n * N);
      if (final_status) *final_status = 0;
      break;
    }
    theta_opt = std::max(0.0, theta_opt - sigma);  // :412
    if (++guard > 10000) { if (final_status) *final_status = st; break; }
  }
  *mu_out = mu; *sigma_out = sigma; *nz_used = zi;
  return 0;
}

typedef struct {
  double alpha, beta, gamma, eps, lambda; int32_t iter_max;
  double theta_high_init, theta_low_init;
  double c_high, c_low; int32_t has_c_high, has_c_low;  // persistent vertex costs (the NM quirk)
} oracle_nm_opts;  // nelder_mead_bilevel_optimization.jl:102-119

// solve!(nm_solver, ...)  nelder_mead_bilevel_optimization.jl:276-352
int32_t oracle_nm_solve(const ratilqr_problem_desc* desc, const ratilqr_ileqg_opts* iopts, oracle_nm_opts* nm,
                        const double* x0, const double* u_init, double kl_bound, double* theta_opt_out,
                        double* value_out, int32_t* nm_iters, int32_t* n_evals, double* x, double* l, double* L,
                        int32_t* final_status) {
  int n = desc->n, m = desc->m, N = desc->N;
  WSet ws;
  if (!prep_Wset(n, N, desc->W, desc->W_time_varying, ws)) return -2;
  Problem pr = make_problem(desc, 0);
  int evals = 0;
  auto cost_at = [&](double theta) {  // compute_cost_worker :134-158
    Solver s = make_solver(*iopts);
    evals++;
    int st = ileqg_solve(s, pr, ws, x0, u_init, theta);
    return st == 0 ? s.value + kl_bound / theta : INF;
  };
  int iter = 0;
  double th_low = nm->theta_low_init, th_high = nm->theta_high_init;  // initialize! :164-168
  double theta_opt;
  if (kl_bound > 0) {
    if (!nm->has_c_high) {  // :283-293
      while (true) {
        nm->c_high = cost_at(th_high);
        if (!std::isinf(nm->c_high)) break;
        th_high *= nm->lambda; nm->theta_high_init *= nm->lambda;
      }
      nm->has_c_high = 1;
    }
    if (!nm->has_c_low) {  // :294-304
      while (true) {
        nm->c_low = cost_at(th_low);
        if (!std::isinf(nm->c_low)) break;
        th_low *= nm->lambda; nm->theta_low_init *= nm->lambda;
      }
      nm->has_c_low = 1;
    }
    while (true) {
      // step! :174-252
      iter++;
      if (nm->c_high < nm->c_low) { std::swap(th_low, th_high); std::swap(nm->c_low, nm->c_high); }
      double th_m = th_low;
      double th_r = std::max(nm->theta_low_init, th_m + nm->alpha * (th_m - th_high));
      double c_r = cost_at(th_r);
      if (c_r < nm->c_low) {
        double th_e = std::max(nm->theta_low_init, th_m + nm->beta * (th_r - th_m));
        double c_e = cost_at(th_e);
        if (c_e < c_r) { th_high = th_e; nm->c_high = c_e; } else { th_high = th_r; nm->c_high = c_r; }
      } else {
        if (c_r < nm->c_high) { th_high = th_r; nm->c_high = c_r; }
        double th_c = std::max(nm->theta_low_init, th_m + nm->gamma * (th_high - th_m));
        double c_c = cost_at(th_c);
        if (c_c > nm->c_high) { th_high = (th_high + th_low) / 2; nm->c_high = cost_at(th_high); }
        else { th_high = th_c; nm->c_high = c_c; }
      }
      double c_mean = (nm->c_low + nm->c_high) / 2;  // :309-310
      double d1 = nm->c_high - c_mean, d2 = nm->c_low - c_mean;
      double stdev = std::sqrt(0.5 * (d1 * d1 + d2 * d2));
      if (stdev < nm->eps) break;
      if (iter == nm->iter_max) break;
    }
    theta_opt = th_low;  // :325
  } else {
    theta_opt = 0.0;
  }
  Solver s = make_solver(*iopts);  // :334-346 (no retry)
  int st = ileqg_solve(s, pr, ws, x0, u_init, theta_opt);
  *theta_opt_out = theta_opt;
  *value_out = st == 0 ? (kl_bound > 0 ? s.value + kl_bound / theta_opt : s.value) : INF;
  if (nm_iters) *nm_iters = iter;
  if (n_evals) *n_evals = evals;
  if (final_status) *final_status = st;
  if (st == 0) {
    if (x) std::memcpy(x, s.x.data(), sizeof(double) * n * (N + 1));
    if (l) std::memcpy(l, s.l.data(), sizeof(double) * m * N);
    if (L) std::memcpy(L, s.L.data(), sizeof(double) * m * n * N);
  }
  return 0;
}

// schedule helpers exposed for the reference's unit tests (ileqg_test.jl:137-148)
int32_t oracle_increase_mu_delta(double mu_min, double delta_0, double* mu, double* delta) {
  *delta = std::max(delta_0, *delta * delta_0);
  *mu = std::max(mu_min, *mu * *delta);
  return 0;
}
int32_t oracle_decrease_mu_delta(double mu_min, double delta_0, double* mu, double* delta) {  // ileqg.jl:480-488
  *delta = std::min(1 / delta_0, *delta / delta_0);
  double cand = *mu * *delta;
  *mu = cand >= mu_min ? cand : 0.0;
  return 0;
}

}  // extern "C"
