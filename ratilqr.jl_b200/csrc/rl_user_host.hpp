// rl_user_host.hpp -- user-extensible device models: the host half.
// A user snippet (dynamics and/or cost, see rl_user.cuh) is compiled with NVRTC for sm_100a against the headers
// embedded in the library, loaded with the driver API and launched through the same argument blocks as the
// registered models.  libnvrtc and libcuda are dlopen'ed on first use, so the library itself has no link-time
// dependency on either (it loads on a CPU-only box; only registering a user model needs them).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace rlu {

enum Kernel { K_SOLVE = 0, K_ROLLOUT_OPEN, K_ROLLOUT_CLOSED, K_INTEGRATE_COST, K_LINEARIZE, K_MC_ROLLOUT, K_PETS_COSTS, K_COUNT };

struct Spec {
  int n = 0, m = 0, n_model_params = 0, n_cost_params = 0;
  int base_model_id = 0, base_cost_id = 0;  // used when the corresponding snippet is empty
  std::string dynamics_src, cost_src;
  // declared structure (empty = dense): 0 zero, 1 one, 2 general; column-major.  a n*n, b n*m (dynamics snippet);
  // q n*n, r m*m, p m*n (cost snippet)
  std::vector<signed char> a_kind, b_kind, q_kind, r_kind, p_kind;
};

struct Compiled {
  std::string cubin, log;
  std::string lowered[K_COUNT];  // mangled kernel names
};

struct Module {
  Spec spec;
  int id = 0;
  int cost_id = 0;       // what desc->cost_id must be for this module
  bool differentiable = true;
  void* cu_module = nullptr;
  void* fn[K_COUNT] = {};
  size_t solve_smem = 0;
  int solve_threads = 64;  // launch shape of the solve kernel (threads per CTA; the register cap follows from min CTAs/SM)
};

// the translation unit handed to NVRTC (exposed for diagnostics / tests)
std::string make_source(const Spec& s);
// NVRTC only (no GPU needed): 0 ok, -20 NVRTC unavailable, -21 compilation failed (log says why), -1 bad spec
int compile(const Spec& s, Compiled& out);
// needs a current CUDA context on this thread: 0 ok, -22 driver unavailable, -23 load failed
int load(const Compiled& c, Module& m, std::string& err);
void unload(Module& m);
// one by-value argument block; returns 0 or -24 (err filled)
int launch(const Module& m, Kernel k, unsigned gx, unsigned gy, unsigned block, size_t smem, cudaStream_t st, void* args,
           std::string& err);

}  // namespace rlu
