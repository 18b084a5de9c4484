// foh.cu -- stage 1 entry points: first-order-hold discretisation of the augmented ODE and the nonlinear propagators, fp64 RK4,
// sm_100a, for the two shipped models (device code: foh_kernels.cuh; user models: user_model.cu compiles the same text with NVRTC).
//
// Replaces FirstOrderHold.calculate_discretization / _ode_dVdt / integrate_nonlinear_* (SCvx/discretization/first_order_hold.py:52-162).
#include "common.cuh"
#include "foh_kernels.cuh"

namespace scvx {

thread_local char g_last_error[256] = "";

template <class M>
__global__ void __launch_bounds__(128)
foh_rk4_kernel(int n_agents, int K, int n_sub, const double* __restrict__ X, const double* __restrict__ U,
               const double* __restrict__ sigma_arr, double* __restrict__ A_bar, double* __restrict__ B_bar,
               double* __restrict__ C_bar, double* __restrict__ S_bar, double* __restrict__ z_bar) {
  foh_rk4_body<M>(n_agents, K, n_sub, X, U, sigma_arr, A_bar, B_bar, C_bar, S_bar, z_bar);
}

template <class M>
__global__ void __launch_bounds__(128)
integrate_piecewise_kernel(int n_agents, int K, int n_sub, const double* __restrict__ X_lin, const double* __restrict__ U,
                           const double* __restrict__ sigma_arr, double* __restrict__ X_nl) {
  integrate_piecewise_body<M>(n_agents, K, n_sub, X_lin, U, sigma_arr, X_nl);
}

template <class M>
__global__ void __launch_bounds__(128)
integrate_full_kernel(int n_agents, int K, int n_sub, const double* __restrict__ x0, const double* __restrict__ U,
                      const double* __restrict__ sigma_arr, double* __restrict__ X_nl) {
  integrate_full_body<M>(n_agents, K, n_sub, x0, U, sigma_arr, X_nl);
}

}  // namespace scvx

using namespace scvx;

extern "C" int scvx_abi_version(void) { return SCVX_ABI_VERSION; }
extern "C" const char* scvx_last_error(void) { return g_last_error; }
extern "C" int scvx_model_dims(int model_id, int* n_x, int* n_u, int* d) {
  int a, b, c;
  if (!model_dims(model_id, &a, &b, &c)) return bad_arg("model_id");
  if (n_x) *n_x = a;
  if (n_u) *n_u = b;
  if (d) *d = c;
  return SCVX_OK;
}

extern "C" int scvx_foh_batched(int model_id, int n_agents, int K, int n_sub, const double* X, const double* U,
                                const double* sigma, double* A_bar, double* B_bar, double* C_bar, double* S_bar,
                                double* z_bar, void* stream) {
  if (n_agents < 0 || K < 2 || n_sub < 0) return bad_arg("n_agents/K/n_sub");
  if (n_agents == 0) return SCVX_OK;
  if (!X || !U || !sigma || !A_bar || !B_bar || !C_bar || !S_bar || !z_bar) return bad_arg("null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n_agents * (K - 1);
  const int threads = 128;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  switch (model_id) {
    case SCVX_MODEL_UNICYCLE:
      foh_rk4_kernel<Unicycle><<<blocks, threads, 0, st>>>(n_agents, K, n_sub, X, U, sigma, A_bar, B_bar, C_bar,
                                                           S_bar, z_bar);
      break;
    case SCVX_MODEL_SINGLE_INTEGRATOR:
      foh_rk4_kernel<SingleIntegrator><<<blocks, threads, 0, st>>>(n_agents, K, n_sub, X, U, sigma, A_bar, B_bar,
                                                                   C_bar, S_bar, z_bar);
      break;
    default: {
      void* args[] = {&n_agents, &K, &n_sub, &X, &U, &sigma, &A_bar, &B_bar, &C_bar, &S_bar, &z_bar};
      return user_model_launch(model_id, 0, blocks, threads, args, st, "scvx_foh_batched (user model)");
    }
  }
  SCVX_CHECK_LAUNCH("scvx_foh_batched");
  return SCVX_OK;
}

extern "C" int scvx_integrate_piecewise_batched(int model_id, int n_agents, int K, int n_sub, const double* X_lin,
                                                const double* U, const double* sigma, double* X_nl, void* stream) {
  if (n_agents < 0 || K < 2 || n_sub < 0) return bad_arg("n_agents/K/n_sub");
  if (n_agents == 0) return SCVX_OK;
  if (!X_lin || !U || !sigma || !X_nl) return bad_arg("null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n_agents * K;
  const int threads = 128;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  switch (model_id) {
    case SCVX_MODEL_UNICYCLE:
      integrate_piecewise_kernel<Unicycle><<<blocks, threads, 0, st>>>(n_agents, K, n_sub, X_lin, U, sigma, X_nl);
      break;
    case SCVX_MODEL_SINGLE_INTEGRATOR:
      integrate_piecewise_kernel<SingleIntegrator><<<blocks, threads, 0, st>>>(n_agents, K, n_sub, X_lin, U, sigma, X_nl);
      break;
    default: {
      void* args[] = {&n_agents, &K, &n_sub, &X_lin, &U, &sigma, &X_nl};
      return user_model_launch(model_id, 1, blocks, threads, args, st, "scvx_integrate_piecewise_batched (user model)");
    }
  }
  SCVX_CHECK_LAUNCH("scvx_integrate_piecewise_batched");
  return SCVX_OK;
}

extern "C" int scvx_integrate_full_batched(int model_id, int n_agents, int K, int n_sub, const double* x0,
                                           const double* U, const double* sigma, double* X_nl, void* stream) {
  if (n_agents < 0 || K < 2 || n_sub < 0) return bad_arg("n_agents/K/n_sub");
  if (n_agents == 0) return SCVX_OK;
  if (!x0 || !U || !sigma || !X_nl) return bad_arg("null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 64;
  const unsigned blocks = (unsigned)((n_agents + threads - 1) / threads);
  switch (model_id) {
    case SCVX_MODEL_UNICYCLE:
      integrate_full_kernel<Unicycle><<<blocks, threads, 0, st>>>(n_agents, K, n_sub, x0, U, sigma, X_nl);
      break;
    case SCVX_MODEL_SINGLE_INTEGRATOR:
      integrate_full_kernel<SingleIntegrator><<<blocks, threads, 0, st>>>(n_agents, K, n_sub, x0, U, sigma, X_nl);
      break;
    default: {
      void* args[] = {&n_agents, &K, &n_sub, &x0, &U, &sigma, &X_nl};
      return user_model_launch(model_id, 2, blocks, threads, args, st, "scvx_integrate_full_batched (user model)");
    }
  }
  SCVX_CHECK_LAUNCH("scvx_integrate_full_batched");
  return SCVX_OK;
}
