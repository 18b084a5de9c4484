// user_model.cu -- dynamics models supplied by the user at run time (row a5 of the scope table: the BaseModel plug-in contract,
// SCvx/models/base_model.py:16-88, where the reference accepts ANY model whose get_equations() yields f, A, B; its own models
// build them with sympy, unicycle_model.py:54-63).
//
// The Python mirror turns the user's sympy expressions into a `struct UserModel { NX, NU, D, eval, f_only }` in CUDA C++
// (scvx_b200/codegen.py), appends the text of foh_kernels.cuh -- the very templates the shipped models are compiled from -- and
// three extern "C" kernels that instantiate them.  scvx_user_model_register compiles that program for sm_100a with NVRTC
// (loaded with dlopen: the library has no link-time dependency on it), loads the cubin through the runtime's library API and
// hands back a model id >= SCVX_MODEL_USER_BASE that every stage-1 entry point accepts.  Stages 2-3 depend on a model's
// DIMENSIONS and constraint kind only, so a registered model runs through the existing kernels of its shape:
// (n_x, n_u, d) = (3, 2, 2) with the unicycle's input box, (3, 3, 3) with the single integrator's speed ball.
#include <dlfcn.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace scvx {

namespace {

struct UserModel {
  int nx, nu, d;
  cudaLibrary_t lib;
  cudaKernel_t foh, piecewise, full;
};

std::mutex g_mu;
std::vector<UserModel> g_models;
std::string g_log;

// the handful of NVRTC entry points used, resolved at run time
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* handle = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
} g_nvrtc;

bool load_nvrtc(const char* hint) {
  if (g_nvrtc.handle) return true;
  const char* names[] = {hint, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    g_nvrtc.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (g_nvrtc.handle) break;
  }
  if (!g_nvrtc.handle) return false;
#define SCVX_SYM(field, name) *(void**)(&g_nvrtc.field) = dlsym(g_nvrtc.handle, name)
  SCVX_SYM(CreateProgram, "nvrtcCreateProgram");
  SCVX_SYM(CompileProgram, "nvrtcCompileProgram");
  SCVX_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
  SCVX_SYM(GetProgramLog, "nvrtcGetProgramLog");
  SCVX_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
  SCVX_SYM(GetCUBIN, "nvrtcGetCUBIN");
  SCVX_SYM(DestroyProgram, "nvrtcDestroyProgram");
  SCVX_SYM(GetErrorString, "nvrtcGetErrorString");
#undef SCVX_SYM
  return g_nvrtc.CreateProgram && g_nvrtc.CompileProgram && g_nvrtc.GetCUBINSize && g_nvrtc.GetCUBIN && g_nvrtc.DestroyProgram &&
         g_nvrtc.GetProgramLogSize && g_nvrtc.GetProgramLog;
}

const UserModel* find(int model_id) {
  const int i = model_id - SCVX_MODEL_USER_BASE;
  if (i < 0 || i >= (int)g_models.size()) return nullptr;
  return &g_models[i];
}

}  // namespace

bool user_model_dims(int model_id, int* nx, int* nu, int* d) {
  std::lock_guard<std::mutex> lk(g_mu);
  const UserModel* m = find(model_id);
  if (!m) return false;
  *nx = m->nx; *nu = m->nu; *d = m->d;
  return true;
}

// launch one of the three stage-1 kernels of a registered model; `which`: 0 foh, 1 piecewise, 2 full
int user_model_launch(int model_id, int which, unsigned blocks, unsigned threads, void** args, cudaStream_t st, const char* where) {
  cudaKernel_t k;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    const UserModel* m = find(model_id);
    if (!m) return bad_arg("model_id");
    k = (which == 0) ? m->foh : (which == 1) ? m->piecewise : m->full;
  }
  const cudaError_t e = cudaLaunchKernel((const void*)k, dim3(blocks), dim3(threads), args, 0, st);
  if (e != cudaSuccess) return cuda_fail(e, where);
  return SCVX_OK;
}

}  // namespace scvx

using namespace scvx;

extern "C" const char* scvx_user_model_log(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_log.c_str();
}

extern "C" int scvx_user_model_register(const char* program, int n_x, int n_u, int d, const char* nvrtc_path, int* model_id) {
  if (!program || !model_id) return bad_arg("null pointer");
  if (n_x < 1 || n_x > 8 || n_u < 1 || n_u > 8 || d < 1 || d > 3 || d > n_x) return bad_arg("n_x/n_u/d (supported: 1..8, 1..8, 1..3)");
  std::lock_guard<std::mutex> lk(g_mu);
  g_log.clear();
  if (!load_nvrtc(nvrtc_path)) {
    const char* why = dlerror();
    snprintf(g_last_error, sizeof(g_last_error), "libnvrtc.so.12 could not be loaded (%s)", why ? why : "no such library");
    return SCVX_E_UNSUPPORTED;
  }
  nvrtcProgram prog = nullptr;
  int rc = g_nvrtc.CreateProgram(&prog, program, "scvx_user_model.cu", 0, nullptr, nullptr);
  if (rc != 0) {
    snprintf(g_last_error, sizeof(g_last_error), "nvrtcCreateProgram failed (%d)", rc);
    return SCVX_E_CUDA;
  }
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--fmad=true"};
  rc = g_nvrtc.CompileProgram(prog, 4, opts);
  size_t log_size = 0;
  g_nvrtc.GetProgramLogSize(prog, &log_size);
  if (log_size > 1) {
    g_log.resize(log_size);
    g_nvrtc.GetProgramLog(prog, &g_log[0]);
  }
  if (rc != 0) {
    g_nvrtc.DestroyProgram(&prog);
    snprintf(g_last_error, sizeof(g_last_error), "NVRTC could not compile the user model (%s); full log: scvx_user_model_log()",
             g_nvrtc.GetErrorString ? g_nvrtc.GetErrorString(rc) : "?");
    return SCVX_E_BADARG;
  }
  size_t cubin_size = 0;
  g_nvrtc.GetCUBINSize(prog, &cubin_size);
  std::vector<char> cubin(cubin_size);
  g_nvrtc.GetCUBIN(prog, cubin.data());
  g_nvrtc.DestroyProgram(&prog);

  UserModel m;
  m.nx = n_x; m.nu = n_u; m.d = d;
  cudaError_t e = cudaLibraryLoadData(&m.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLibraryLoadData");
  if ((e = cudaLibraryGetKernel(&m.foh, m.lib, "scvx_user_foh")) != cudaSuccess) return cuda_fail(e, "cudaLibraryGetKernel(scvx_user_foh)");
  if ((e = cudaLibraryGetKernel(&m.piecewise, m.lib, "scvx_user_piecewise")) != cudaSuccess) return cuda_fail(e, "cudaLibraryGetKernel(scvx_user_piecewise)");
  if ((e = cudaLibraryGetKernel(&m.full, m.lib, "scvx_user_full")) != cudaSuccess) return cuda_fail(e, "cudaLibraryGetKernel(scvx_user_full)");
  g_models.push_back(m);
  *model_id = SCVX_MODEL_USER_BASE + (int)g_models.size() - 1;
  return SCVX_OK;
}
