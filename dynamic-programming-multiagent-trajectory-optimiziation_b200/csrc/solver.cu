// solver.cu -- stage 3 (placeholder while the IPM kernel is being brought up)
#include "common.cuh"
using namespace scvx;
extern "C" unsigned long long scvx_solve_workspace_bytes(int, int, int, int, int) { return 0ull; }
extern "C" int scvx_solve_batched(const scvx_solve_args*, void*) {
  snprintf(g_last_error, sizeof(g_last_error), "scvx_solve_batched: not built yet");
  return SCVX_E_UNSUPPORTED;
}
