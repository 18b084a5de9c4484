// solver_unicycle_fixed.cu -- ipm_kernel<Unicycle> with the sizes of the BASELINE configs as compile-time constants (solver_kernel.cuh,
// template parameters KT / MT): K = 100 nodes on a block of 128 threads (configs 2, 3) and, for the plain sub-problem of SCProblem, M = 8
// obstacle rows per stage (config 2); K = 50, M = 3 (config 1); K = 200 with the Jacobians in global memory (config 5).
// Same code as the generic kernel, results equal to round-off (not bitwise: other multiply-add pairs get contracted); 37 % of the generic kernel's executed instructions were integer arithmetic
// and register moves, most of it address computation with K -- here every row-state / shared-memory address is base + k + constant.
#include "solver_kernel.cuh"

namespace scvx {
template <>
int launch_ipm_fixed<Unicycle>(const scvx_solve_args& a, cudaStream_t st, size_t smem, int threads, size_t jac_off, bool jsm, int G, int C) {
  if (G != 1 || C != 1) return IPM_NOT_FIXED;
  if (jsm && a.K == 100 && threads == 128) {            // configs 2 (plain, M = 8) and 3 (inter-agent rows)
    if (a.M == 8 && ipm_args_plain(a)) return launch_ipm_kernel(ipm_kernel<Unicycle, true, 1, 1, false, 100, 8>, 1, a, st, smem, threads, jac_off);
    return launch_ipm_kernel(ipm_kernel<Unicycle, true, 1, 1, false, 100>, 1, a, st, smem, threads, jac_off);
  }
  if (jsm && a.K == 50 && threads == 64 && a.M == 3 && ipm_args_plain(a))      // config 1: the shipped single-agent problem
    return launch_ipm_kernel(ipm_kernel<Unicycle, true, 1, 1, false, 50, 3>, 1, a, st, smem, threads, jac_off);
  if (!jsm && a.K == 200 && threads == 224)             // config 5: K = 200, interval Jacobians in the global workspace
    return launch_ipm_kernel(ipm_kernel<Unicycle, false, 1, 1, false, 200>, 1, a, st, smem, threads, jac_off);
  return IPM_NOT_FIXED;
}
}  // namespace scvx
