// common.cuh -- shared definitions for the sm_100a SCvx kernels (fp64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include "../../include/scvx_b200.h"

namespace scvx {

// ---- error plumbing ---------------------------------------------------------------------------
extern thread_local char g_last_error[256];
inline int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", where, cudaGetErrorString(e));
  return SCVX_E_CUDA;
}
inline int bad_arg(const char* what) {
  snprintf(g_last_error, sizeof(g_last_error), "bad argument: %s", what);
  return SCVX_E_BADARG;
}
#define SCVX_CHECK_LAUNCH(where)                                   \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return ::scvx::cuda_fail(e__, where);  \
  } while (0)

// ---- dynamics models as device functions --------------------------------------------------------
// Each model exposes NX, NU, D (position dimension) and eval(x, u, f, A, B) with A, B dense
// row-major [NX][NX], [NX][NU].  These are the device twins of the reference's
// Model.get_equations() lambdas.

// SCvx/models/unicycle_model.py:54-63   f = [v cos th, v sin th, w]
struct Unicycle {
  static constexpr int NX = 3, NU = 2, D = 2;
  __device__ __forceinline__ static void eval(const double* x, const double* u, double* f, double (*A)[3],
                                              double (*B)[2]) {
    double s, c;
    sincos(x[2], &s, &c);
    f[0] = u[0] * c; f[1] = u[0] * s; f[2] = u[1];
    A[0][0] = 0.0; A[0][1] = 0.0; A[0][2] = -u[0] * s;
    A[1][0] = 0.0; A[1][1] = 0.0; A[1][2] = u[0] * c;
    A[2][0] = 0.0; A[2][1] = 0.0; A[2][2] = 0.0;
    B[0][0] = c;   B[0][1] = 0.0;
    B[1][0] = s;   B[1][1] = 0.0;
    B[2][0] = 0.0; B[2][1] = 1.0;
  }
  __device__ __forceinline__ static void f_only(const double* x, const double* u, double* f) {
    double s, c;
    sincos(x[2], &s, &c);
    f[0] = u[0] * c; f[1] = u[0] * s; f[2] = u[1];
  }
};

// SCvx/models/single_integrator_model.py:53-57   f = u, A = 0, B = I
struct SingleIntegrator {
  static constexpr int NX = 3, NU = 3, D = 3;
  __device__ __forceinline__ static void eval(const double* x, const double* u, double* f, double (*A)[3],
                                              double (*B)[3]) {
    (void)x;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      f[i] = u[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) { A[i][j] = 0.0; B[i][j] = (i == j) ? 1.0 : 0.0; }
    }
  }
  __device__ __forceinline__ static void f_only(const double* x, const double* u, double* f) {
    (void)x;
#pragma unroll
    for (int i = 0; i < 3; ++i) f[i] = u[i];
  }
};

// models registered at run time (user_model.cu)
bool user_model_dims(int model_id, int* nx, int* nu, int* d);
int user_model_launch(int model_id, int which, unsigned blocks, unsigned threads, void** args, cudaStream_t st, const char* where);

inline bool model_dims(int model_id, int* nx, int* nu, int* d) {
  switch (model_id) {
    case SCVX_MODEL_UNICYCLE: *nx = 3; *nu = 2; *d = 2; return true;
    case SCVX_MODEL_SINGLE_INTEGRATOR: *nx = 3; *nu = 3; *d = 3; return true;
    default: return user_model_dims(model_id, nx, nu, d);
  }
}

// Stages 2-3 depend on a model's dimensions and constraint kind only: a registered model runs through the kernels of the shipped
// model with its shape.  Returns SCVX_MODEL_UNICYCLE / SCVX_MODEL_SINGLE_INTEGRATOR, or -1 when no kernel of that shape exists.
inline int solver_shape_of(int model_id) {
  if (model_id == SCVX_MODEL_UNICYCLE || model_id == SCVX_MODEL_SINGLE_INTEGRATOR) return model_id;
  int nx, nu, d;
  if (!user_model_dims(model_id, &nx, &nu, &d)) return -1;
  if (nx == 3 && nu == 2 && d == 2) return SCVX_MODEL_UNICYCLE;
  if (nx == 3 && nu == 3 && d == 3) return SCVX_MODEL_SINGLE_INTEGRATOR;
  return -1;
}

}  // namespace scvx
