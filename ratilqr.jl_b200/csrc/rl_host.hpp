// rl_host.hpp -- host-side helpers of the C ABI: W pre-processing and the (model, cost)
// dispatch table.  Plain C++ (no CUDA), shared by rl_capi.cu and by tests/_hostemu.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/ratilqr.h"

namespace rlh {

// inv(W) (ileqg.jl:365), det(W) (logdet(W*M) = log(det W * det M), :387) and chol_lower(W)
// (noise sampling, rand(rng, MvNormal(0, W)), ileqg.jl:51,104) for every stage's W.
struct WPrep {
  std::vector<double> W, Winv, cholW, detW;  // n*n*cnt, n*n*cnt, n*n*cnt, cnt
  int cnt = 0;
};

inline bool chol_lower_host(int n, const double* M, double* C, double* invd, double* det) {
  double dprod = 1.0;
  for (int j = 0; j < n; ++j) {
    double d = M[j + j * n];
    for (int k = 0; k < j; ++k) d = std::fma(-C[j + k * n], C[j + k * n], d);
    if (!(d > 0.0)) return false;
    dprod = (j == 0) ? d : dprod * d;
    double cjj = std::sqrt(d), inv = 1.0 / cjj;
    C[j + j * n] = cjj;
    invd[j] = inv;
    for (int i = j + 1; i < n; ++i) {
      double a = M[j + i * n];
      for (int k = 0; k < j; ++k) a = std::fma(-C[i + k * n], C[j + k * n], a);
      C[i + j * n] = a * inv;
    }
    for (int i = 0; i < j; ++i) C[i + j * n] = 0.0;
  }
  *det = dprod;
  return true;
}

// returns false when some W(k) is not positive definite (the reference's inv(W)/MvNormal need PD)
inline bool prep_W(int n, int N, const double* W, int time_varying, WPrep& o) {
  o.cnt = time_varying ? N : 1;
  size_t nn = (size_t)n * n;
  o.W.assign(W, W + nn * o.cnt);
  o.Winv.assign(nn * o.cnt, 0.0);
  o.cholW.assign(nn * o.cnt, 0.0);
  o.detW.assign(o.cnt, 0.0);
  std::vector<double> invd(n), Y(nn);
  for (int c = 0; c < o.cnt; ++c) {
    double* C = &o.cholW[nn * c];
    if (!chol_lower_host(n, W + nn * c, C, invd.data(), &o.detW[c])) return false;
    for (int col = 0; col < n; ++col)  // Y = C^-1
      for (int i = 0; i < n; ++i) {
        double a = (i == col) ? 1.0 : 0.0;
        for (int k = 0; k < i; ++k) a = std::fma(-C[i + k * n], Y[k + col * n], a);
        Y[i + col * n] = a * invd[i];
      }
    double* Wi = &o.Winv[nn * c];
    for (int i = 0; i < n; ++i)  // inv(W) = Y'Y
      for (int j = i; j < n; ++j) {
        double e = Y[i * n] * Y[j * n];
        for (int k = 1; k < n; ++k) e = std::fma(Y[k + i * n], Y[k + j * n], e);
        Wi[i + j * n] = e;
        Wi[j + i * n] = e;
      }
  }
  return true;
}

// true-model noise mixture -> cumulative normalised weights + lower Cholesky factors; nullptr if fine, else a message
struct MixPrep { std::vector<double> cumw, mean, chol; int k = 0; };
inline const char* prep_mixture(int n, const ratilqr_noise_mixture* mx, MixPrep& o) {
  if (!mx || mx->n_components < 1 || !mx->weights || !mx->means || !mx->covs) return "bad noise mixture";
  o.k = mx->n_components;
  double tot = 0.0;
  for (int c = 0; c < o.k; ++c) { if (!(mx->weights[c] > 0.0)) return "mixture weights must be positive"; tot += mx->weights[c]; }
  o.cumw.resize(o.k);
  double acc = 0.0;
  for (int c = 0; c < o.k; ++c) { acc += mx->weights[c] / tot; o.cumw[c] = acc; }
  o.cumw[o.k - 1] = 1.0;
  o.mean.assign(mx->means, mx->means + (size_t)n * o.k);
  o.chol.assign((size_t)n * n * o.k, 0.0);
  std::vector<double> invd(n);
  double det;
  for (int c = 0; c < o.k; ++c)
    if (!chol_lower_host(n, mx->covs + (size_t)c * n * n, &o.chol[(size_t)c * n * n], invd.data(), &det))
      return "a mixture covariance is not positive definite";
  return nullptr;
}

inline bool model_dims(int id, int* n, int* m, int* np) {
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

inline int cost_param_count(int cost_id, int n, int m) {
  switch (cost_id) {
    case RATILQR_COST_QUADRATIC: return 5 + n + n * n + m * m + n * m + n * n;
    case RATILQR_COST_POWER_LAW: return 2;
    case RATILQR_COST_L1_CONTROL: return 1;
  }
  return -1;
}

// validates a problem description; returns nullptr if fine, else a message
inline const char* check_desc(const ratilqr_problem_desc* d, bool differentiable) {
  int n, m, np;
  if (!d) return "null problem description";
  if (!model_dims(d->model_id, &n, &m, &np)) return "unknown model_id";
  if (d->n != n || d->m != m) return "n/m do not match the registered model";
  if (d->n_model_params != np || !d->model_params) return "wrong number of model parameters";
  if (d->N < 1) return "N must be >= 1";
  int ncp = cost_param_count(d->cost_id, n, m);
  if (ncp < 0) return "unknown cost_id";
  if (d->n_cost_params != ncp || !d->cost_params) return "wrong number of cost parameters";
  if (d->cost_params_count < 1) return "cost_params_count must be >= 1";
  if (differentiable && d->cost_id == RATILQR_COST_L1_CONTROL) return "L1_CONTROL cost is rollout-only (PETS)";
  if (d->cost_id == RATILQR_COST_POWER_LAW && n != m) return "POWER_LAW cost needs n == m";
  if (!d->W) return "W is null";
  return nullptr;
}

}  // namespace rlh

// internal cost id (see rl_core.cuh): QUADRATIC with diagonal Q, R, Qf and Pc == 0
#ifndef RL_COST_QUAD_DIAG
#define RL_COST_QUAD_DIAG 0x101
#endif

namespace rlh {
// true when every parameter block of a QUADRATIC cost has diagonal Q, R, Qf and zero Pc
inline bool quad_is_diag(const ratilqr_problem_desc* d) {
  if (d->cost_id != RATILQR_COST_QUADRATIC) return false;
  const int n = d->n, m = d->m;
  for (int c = 0; c < d->cost_params_count; ++c) {
    const double* p = d->cost_params + (size_t)c * d->n_cost_params;
    const double *Q = p + 5 + n, *R = Q + n * n, *Pc = R + m * m, *Qf = Pc + n * m;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (i != j && (Q[i + j * n] != 0.0 || Qf[i + j * n] != 0.0)) return false;
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) if (i != j && R[i + j * m] != 0.0) return false;
    for (int i = 0; i < n * m; ++i) if (Pc[i] != 0.0) return false;
  }
  return true;
}
}  // namespace rlh

// (model, cost) pairs compiled into the library.  X(model_id, cost_id)
#define RL_FOR_EACH_ILEQG_COMBO(X)                                   \
  X(RATILQR_MODEL_SINGLE_INTEGRATOR, RATILQR_COST_QUADRATIC)         \
  X(RATILQR_MODEL_POWER_LAW, RATILQR_COST_POWER_LAW)                 \
  X(RATILQR_MODEL_POWER_LAW, RATILQR_COST_QUADRATIC)                 \
  X(RATILQR_MODEL_DOUBLE_INTEGRATOR, RATILQR_COST_QUADRATIC)         \
  X(RATILQR_MODEL_PENDULUM, RATILQR_COST_QUADRATIC)                  \
  X(RATILQR_MODEL_CARTPOLE, RATILQR_COST_QUADRATIC)                  \
  X(RATILQR_MODEL_UNICYCLE, RATILQR_COST_QUADRATIC)                  \
  X(RATILQR_MODEL_QUADROTOR, RATILQR_COST_QUADRATIC)

// structure-specialised solve kernels (host selects them when rlh::quad_is_diag)
#define RL_FOR_EACH_DIAG_COMBO(X)                                    \
  X(RATILQR_MODEL_SINGLE_INTEGRATOR, RL_COST_QUAD_DIAG)              \
  X(RATILQR_MODEL_DOUBLE_INTEGRATOR, RL_COST_QUAD_DIAG)              \
  X(RATILQR_MODEL_PENDULUM, RL_COST_QUAD_DIAG)                       \
  X(RATILQR_MODEL_CARTPOLE, RL_COST_QUAD_DIAG)                       \
  X(RATILQR_MODEL_UNICYCLE, RL_COST_QUAD_DIAG)

// rollout-only pairs (PETS / MC) in addition to the ones above
#define RL_FOR_EACH_ROLLOUT_ONLY_COMBO(X)                            \
  X(RATILQR_MODEL_SINGLE_INTEGRATOR, RATILQR_COST_L1_CONTROL)        \
  X(RATILQR_MODEL_CARTPOLE, RATILQR_COST_L1_CONTROL)
