// rl_user_host.cu -- NVRTC compilation + driver-API loading of user device models (see rl_user_host.hpp).
#include "rl_user_host.hpp"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "rl_embedded_headers.inc"
#include "rl_host.hpp"

namespace rlu {

namespace {

struct NvrtcApi {
  void* h = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
  nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
};

struct DriverApi {
  void* h = nullptr;
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
};

template <class F> bool sym(void* h, const char* name, F& f) {
  f = reinterpret_cast<F>(dlsym(h, name));
  return f != nullptr;
}

NvrtcApi* nvrtc_api() {
  static NvrtcApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<std::string> names = {"libnvrtc.so.12", "libnvrtc.so"};
    for (const char* env : {"CUDA_HOME", "CUDA_PATH"})
      if (const char* root = getenv(env)) names.push_back(std::string(root) + "/lib64/libnvrtc.so.12");
    names.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    names.push_back("/usr/local/cuda/lib64/libnvrtc.so");
    void* h = nullptr;
    for (const auto& nm : names)
      if ((h = dlopen(nm.c_str(), RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) return;
    bool ok = sym(h, "nvrtcCreateProgram", api.CreateProgram) && sym(h, "nvrtcDestroyProgram", api.DestroyProgram) &&
              sym(h, "nvrtcCompileProgram", api.CompileProgram) && sym(h, "nvrtcGetProgramLogSize", api.GetProgramLogSize) &&
              sym(h, "nvrtcGetProgramLog", api.GetProgramLog) && sym(h, "nvrtcGetCUBINSize", api.GetCUBINSize) &&
              sym(h, "nvrtcGetCUBIN", api.GetCUBIN) && sym(h, "nvrtcAddNameExpression", api.AddNameExpression) &&
              sym(h, "nvrtcGetLoweredName", api.GetLoweredName) && sym(h, "nvrtcGetErrorString", api.GetErrorString);
    if (ok) api.h = h;
  });
  return api.h ? &api : nullptr;
}

DriverApi* driver_api() {
  static DriverApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    bool ok = sym(h, "cuModuleLoadData", api.ModuleLoadData) && sym(h, "cuModuleUnload", api.ModuleUnload) &&
              sym(h, "cuModuleGetFunction", api.ModuleGetFunction) && sym(h, "cuFuncSetAttribute", api.FuncSetAttribute) &&
              sym(h, "cuLaunchKernel", api.LaunchKernel) && sym(h, "cuGetErrorString", api.GetErrorString);
    if (ok) api.h = h;
  });
  return api.h ? &api : nullptr;
}

std::string cu_err(DriverApi* d, CUresult r) {
  const char* s = nullptr;
  d->GetErrorString(r, &s);
  return s ? s : "unknown driver error";
}

// launch shape of the user solve kernel: 64 threads x 4 CTAs/SM (every register available).  The shape tuned for the
// registered unicycle (128 x 3 = 168 registers) was within +-3 % on user pairs (+7 % only with both snippets structured),
// so it stays an override for tuning: RATILQR_USER_SHAPE=128.
static int solve_threads_for(const Spec&) {
  if (const char* e = getenv("RATILQR_USER_SHAPE")) return atoi(e) == 128 ? 128 : 64;
  return 64;
}

const char* kernel_expr(Kernel k, int solve_threads = 64) {
  switch (k) {
    case K_SOLVE: return solve_threads == 128 ? "rll::k_ileqg_solve<RluD, RluC, 128, 3>" : "rll::k_ileqg_solve<RluD, RluC, 64, 4>";
    case K_ROLLOUT_OPEN: return "rll::k_rollout_open<RluD>";
    case K_ROLLOUT_CLOSED: return "rll::k_rollout_closed<RluD, RluC>";
    case K_INTEGRATE_COST: return "rll::k_integrate_cost<RluD, RluC>";
    case K_LINEARIZE: return "rll::k_linearize<RluD, RluC>";
    case K_MC_ROLLOUT: return "rll::k_mc_rollout<RluD, RluC>";
    case K_PETS_COSTS: return "rll::k_pets_costs<RluD, RluC>";
    default: return "";
  }
}

bool differentiable_cost(const Spec& s) { return !s.cost_src.empty() || s.base_cost_id != RATILQR_COST_L1_CONTROL; }

}  // namespace

// "(i == 0 && j == 2) ? 2 : (...) ? 1 : <most common value>" for a rows x cols column-major table of kinds
static std::string kind_expr(const std::vector<signed char>& k, int rows, int cols) {
  int cnt[3] = {0, 0, 0};
  for (signed char v : k) cnt[v < 0 ? 0 : (v > 2 ? 2 : v)]++;
  int dflt = 0;
  for (int v = 1; v < 3; ++v) if (cnt[v] > cnt[dflt]) dflt = v;
  std::string e;
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i) {
      int v = k[(size_t)i + (size_t)j * rows];
      v = v < 0 ? 0 : (v > 2 ? 2 : v);
      if (v != dflt) e += "(i == " + std::to_string(i) + " && j == " + std::to_string(j) + ") ? " + std::to_string(v) + " : ";
    }
  return e + std::to_string(dflt);
}

std::string make_source(const Spec& s) {
  std::string t;
  t += "#include \"rl_user.cuh\"\n#include \"rl_kernels_model.cuh\"\nusing rl::square;\n";
  const std::string N = std::to_string(s.n), M = std::to_string(s.m);
  if (!s.dynamics_src.empty()) {
    t += "namespace ratilqr_user_dynamics {\n#line 1 \"dynamics\"\n" + s.dynamics_src + "\n}\n";
    t += "struct RluBody { template <class T> __device__ void operator()(const double* p, const T* x, const T* u, T* xn) const "
         "{ ratilqr_user_dynamics::dynamics<T>(p, x, u, xn); } };\n";
    if (!s.a_kind.empty() || !s.b_kind.empty()) {
      const std::vector<signed char> dense_a((size_t)s.n * s.n, 2), dense_b((size_t)s.n * s.m, 2);
      t += "struct RluKinds {\n  static constexpr bool structured = true;\n"
           "  RL_HD static constexpr int a_kind(int i, int j) { return " + kind_expr(s.a_kind.empty() ? dense_a : s.a_kind, s.n, s.n) + "; }\n"
           "  RL_HD static constexpr int b_kind(int i, int j) { return " + kind_expr(s.b_kind.empty() ? dense_b : s.b_kind, s.n, s.m) + "; }\n};\n";
      t += "typedef rl::UserDyn<" + N + ", " + M + ", RluBody, RluKinds> RluD;\n";
    } else {
      t += "typedef rl::UserDyn<" + N + ", " + M + ", RluBody> RluD;\n";
    }
  } else {
    t += "typedef rl::Dyn<" + std::to_string(s.base_model_id) + "> RluD;\n";
    t += "static_assert(RluD::n == " + N + " && RluD::m == " + M + ", \"n/m do not match the registered model\");\n";
  }
  if (!s.cost_src.empty()) {
    t += "namespace ratilqr_user_cost {\n#line 1 \"cost\"\n" + s.cost_src + "\n}\n";
    t += "struct RluCostFn {\n"
         "  template <class T> __device__ T stage(const double* cp, int k, const T* x, const T* u) const { return ratilqr_user_cost::stage_cost<T>(cp, k, x, u); }\n"
         "  template <class T> __device__ T terminal(const double* cp, const T* x) const { return ratilqr_user_cost::terminal_cost<T>(cp, x); }\n"
         "};\n";
    if (!s.q_kind.empty() || !s.r_kind.empty() || !s.p_kind.empty()) {
      const std::vector<signed char> dq((size_t)s.n * s.n, 2), dr((size_t)s.m * s.m, 2), dp((size_t)s.m * s.n, 2);
      t += "struct RluCostKinds {\n"
           "  RL_HD static constexpr int q_kind(int i, int j) { return " + kind_expr(s.q_kind.empty() ? dq : s.q_kind, s.n, s.n) + "; }\n"
           "  RL_HD static constexpr int r_kind(int i, int j) { return " + kind_expr(s.r_kind.empty() ? dr : s.r_kind, s.m, s.m) + "; }\n"
           "  RL_HD static constexpr int p_kind(int i, int j) { return " + kind_expr(s.p_kind.empty() ? dp : s.p_kind, s.m, s.n) + "; }\n};\n";
      t += "typedef rl::UserCost<" + N + ", " + M + ", " + std::to_string(s.n_cost_params) + ", RluCostFn, RluCostKinds> RluC;\n";
    } else {
      t += "typedef rl::UserCost<" + N + ", " + M + ", " + std::to_string(s.n_cost_params) + ", RluCostFn> RluC;\n";
    }
  } else {
    t += "typedef rl::Cost<" + std::to_string(s.base_cost_id) + ", " + N + ", " + M + "> RluC;\n";
  }
  return t;
}

int compile(const Spec& s, Compiled& out) {
  out = Compiled();
  if (s.n < 1 || s.n > 16 || s.m < 1 || s.m > 4) { out.log = "user models need 1 <= n <= 16 and 1 <= m <= 4"; return -1; }
  if (s.dynamics_src.empty() && s.cost_src.empty()) { out.log = "neither a dynamics nor a cost snippet was given"; return -1; }
  if (s.n_model_params < 0 || s.n_model_params > 8) { out.log = "at most 8 model parameters"; return -1; }
  if (s.dynamics_src.empty()) {
    int n, m, np;
    if (!rlh::model_dims(s.base_model_id, &n, &m, &np) || n != s.n || m != s.m) { out.log = "base_model_id is not a registered model of this size"; return -1; }
  }
  if (s.cost_src.empty() && rlh::cost_param_count(s.base_cost_id, s.n, s.m) < 0) { out.log = "base_cost_id is not a registered cost"; return -1; }
  if (!s.cost_src.empty() && s.n_cost_params < 0) { out.log = "n_cost_params must be >= 0"; return -1; }
  for (const auto* k : {&s.q_kind, &s.r_kind, &s.p_kind})
    for (signed char v : *k) if (v != 0 && v != 2) { out.log = "q_kind / r_kind / p_kind entries must be 0 or 2"; return -1; }
  for (const auto* k : {&s.a_kind, &s.b_kind})
    for (signed char v : *k) if (v < 0 || v > 2) { out.log = "a_kind / b_kind entries must be 0, 1 or 2"; return -1; }
  NvrtcApi* nv = nvrtc_api();
  if (!nv) { out.log = "libnvrtc.so.12 could not be loaded (set CUDA_HOME or LD_LIBRARY_PATH)"; return -20; }
  const std::string src = make_source(s);
  nvrtcProgram prog = nullptr;
  nvrtcResult r = nv->CreateProgram(&prog, src.c_str(), "ratilqr_user_model.cu", rl_hdr_count, rl_hdr_src, rl_hdr_names);
  if (r != NVRTC_SUCCESS) { out.log = std::string("nvrtcCreateProgram: ") + nv->GetErrorString(r); return -21; }
  const bool diff = differentiable_cost(s);
  for (int k = 0; k < K_COUNT; ++k) {
    if (!diff && (k == K_SOLVE || k == K_LINEARIZE)) continue;
    nv->AddNameExpression(prog, kernel_expr((Kernel)k, solve_threads_for(s)));
  }
  // same code generation rules as the library build (csrc/Makefile): explicit fma only
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-default-device", "-lineinfo"};
  r = nv->CompileProgram(prog, 5, opts);
  size_t lsz = 0;
  nv->GetProgramLogSize(prog, &lsz);
  if (lsz > 1) { out.log.resize(lsz); nv->GetProgramLog(prog, &out.log[0]); out.log.resize(lsz - 1); }
  if (r != NVRTC_SUCCESS) {
    if (out.log.empty()) out.log = nv->GetErrorString(r);
    nv->DestroyProgram(&prog);
    return -21;
  }
  for (int k = 0; k < K_COUNT; ++k) {
    if (!diff && (k == K_SOLVE || k == K_LINEARIZE)) continue;
    const char* low = nullptr;
    if (nv->GetLoweredName(prog, kernel_expr((Kernel)k, solve_threads_for(s)), &low) == NVRTC_SUCCESS && low) out.lowered[k] = low;
  }
  size_t csz = 0;
  nv->GetCUBINSize(prog, &csz);
  out.cubin.resize(csz);
  if (csz) nv->GetCUBIN(prog, &out.cubin[0]);
  nv->DestroyProgram(&prog);
  if (!csz) { out.log += "\nNVRTC produced no cubin"; return -21; }
  return 0;
}

int load(const Compiled& c, Module& m, std::string& err) {
  DriverApi* d = driver_api();
  if (!d) { err = "libcuda.so.1 could not be loaded"; return -22; }
  CUmodule mod = nullptr;
  CUresult r = d->ModuleLoadData(&mod, c.cubin.data());
  if (r != CUDA_SUCCESS) { err = "cuModuleLoadData: " + cu_err(d, r); return -23; }
  m.cu_module = mod;
  for (int k = 0; k < K_COUNT; ++k) {
    m.fn[k] = nullptr;
    if (c.lowered[k].empty()) continue;
    CUfunction f = nullptr;
    r = d->ModuleGetFunction(&f, mod, c.lowered[k].c_str());
    if (r != CUDA_SUCCESS) { err = "cuModuleGetFunction(" + c.lowered[k] + "): " + cu_err(d, r); unload(m); return -23; }
    m.fn[k] = f;
  }
  // thread-private cp.async staging area of the solve kernel (rl::UseStage / rl::RL_STAGE_NV), 64-thread CTAs
  const int nv = m.spec.n + 2 * m.spec.m + m.spec.m * m.spec.n;
  m.solve_threads = solve_threads_for(m.spec);
  const int minb = m.solve_threads == 128 ? 3 : 4;
  m.solve_smem = (nv <= 16) ? (size_t)2 * 16 * m.solve_threads * sizeof(double) : 0;
  if (m.fn[K_SOLVE] && m.solve_smem) {
    int pct = (int)((minb * (m.solve_smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 5;
    d->FuncSetAttribute((CUfunction)m.fn[K_SOLVE], CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, pct > 100 ? 100 : pct);
  }
  return 0;
}

void unload(Module& m) {
  DriverApi* d = driver_api();
  if (d && m.cu_module) d->ModuleUnload((CUmodule)m.cu_module);
  m.cu_module = nullptr;
  for (auto& f : m.fn) f = nullptr;
}

int launch(const Module& m, Kernel k, unsigned gx, unsigned gy, unsigned block, size_t smem, cudaStream_t st, void* args,
           std::string& err) {
  DriverApi* d = driver_api();
  if (!d || !m.fn[k]) { err = "this kernel is not available for the user model (rollout-only cost?)"; return -24; }
  void* params[1] = {args};
  CUresult r = d->LaunchKernel((CUfunction)m.fn[k], gx, gy, 1, block, 1, 1, (unsigned)smem, (CUstream)st, params, nullptr);
  if (r != CUDA_SUCCESS) { err = "cuLaunchKernel: " + cu_err(d, r); return -24; }
  return 0;
}

}  // namespace rlu
