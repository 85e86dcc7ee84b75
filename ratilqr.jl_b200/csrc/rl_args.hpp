// rl_args.hpp -- by-value kernel argument blocks of the model-templated kernels (rl_kernels_model.cuh).
// Plain structs (fixed-width types via rl_core.cuh): the same definitions are seen by nvcc (library build) and by NVRTC
// (user-extensible device models, rl_user.cuh), so a block filled on the host has the layout the kernel expects.
// All pointers are DEVICE pointers.
#pragma once
#include "rl_components.cuh"

namespace rll {

struct CompArgs {
  int model_id, cost_id, n, m, N, B;
  double mp[8];
  const double* cp;  // device, one block
  // rollouts / cost / linearize (host layout, instance slowest)
  const double *x0, *u, *xbar, *l, *L;
  double *x, *u_new, *cost;
  double *q, *qv, *Q, *r, *R, *Pm, *A, *Bm;
  int32_t* status;
};

struct McArgs {
  int model_id, cost_id, N, P, n_samples;
  double mp[8];
  const double* cp; int ncp, cp_count;
  const double *xbar, *l, *L;        // per problem
  const double* noise;               // n*N*n_samples*P or null
  const double* cholW; int W_tv;     // n*n [*N]
  uint64_t seed;
  rl::MixtureView mix;               // true-model noise (Philox mode) when mix.k > 0
  double* J; double* x_out;
};

struct PetsArgs {
  int model_id, cost_id, N, C, particles;
  double mp[8];
  const double* ens_params; int n_ens, n_mp;  // device, or null
  const double* cp;
  const double* x0;
  const double* controls;  // m*N*C
  const double* noise;     // n*N*particles*C or null
  const double* cholW;
  int noise_kind; double noise_scale;
  uint64_t seed; uint64_t stream_offset;
  rl::MixtureView mix;     // use_true_model: mixture noise when mix.k > 0
  double* cost;            // C
};

}  // namespace rll
