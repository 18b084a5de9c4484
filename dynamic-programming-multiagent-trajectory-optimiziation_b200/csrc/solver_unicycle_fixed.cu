// solver_unicycle_fixed.cu -- ipm_kernel<Unicycle> with the sizes of BASELINE config 2 as compile-time constants (solver_kernel.cuh, template
// parameters KT / MT): K = 100 nodes on a block of 128 threads; and, for the plain sub-problem of SCProblem, M = 8 obstacle rows per stage.
// Same code as the generic kernel, results equal to round-off (not bitwise: other multiply-add pairs get contracted); 37 % of the generic kernel's executed instructions were integer arithmetic
// and register moves, most of it address computation with K -- here every row-state / shared-memory address is base + k + constant.
#include "solver_kernel.cuh"

namespace scvx {
template <>
int launch_ipm_fixed<Unicycle>(const scvx_solve_args& a, cudaStream_t st, size_t smem, int threads, size_t jac_off, bool jsm, int G, int C) {
  if (!(jsm && G == 1 && C == 1 && a.K == 100 && threads == 128)) return IPM_NOT_FIXED;
  if (a.M == 8 && ipm_args_plain(a)) return launch_ipm_kernel(ipm_kernel<Unicycle, true, 1, 1, false, 100, 8>, 1, a, st, smem, threads, jac_off);
  return launch_ipm_kernel(ipm_kernel<Unicycle, true, 1, 1, false, 100>, 1, a, st, smem, threads, jac_off);
}
}  // namespace scvx
