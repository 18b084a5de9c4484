// solver_si_fixed.cu -- ipm_kernel<SingleIntegrator> with K = 100 (BASELINE config 4) as a compile-time constant, for the launch shapes
// config 4 uses: two hinge groups per block, alone or in clusters of 2 / 4 blocks per agent (see solver_unicycle_fixed.cu).
#include "solver_kernel.cuh"

namespace scvx {
template <>
int launch_ipm_fixed<SingleIntegrator>(const scvx_solve_args& a, cudaStream_t st, size_t smem, int threads, size_t jac_off, bool jsm, int G, int C) {
  if (!(jsm && a.K == 100)) return IPM_NOT_FIXED;
  if (G == 2 && C == 1) return launch_ipm_kernel(ipm_kernel<SingleIntegrator, true, 2, 1, false, 100>, 1, a, st, smem, threads, jac_off);
  if (G == 2 && C == 2) return launch_ipm_kernel(ipm_kernel<SingleIntegrator, true, 2, 2, false, 100>, 2, a, st, smem, threads, jac_off);
  if (G == 2 && C == 4) return launch_ipm_kernel(ipm_kernel<SingleIntegrator, true, 2, 4, false, 100>, 4, a, st, smem, threads, jac_off);
  if (G == 1 && C == 1 && threads == 128) return launch_ipm_kernel(ipm_kernel<SingleIntegrator, true, 1, 1, false, 100>, 1, a, st, smem, threads, jac_off);
  return IPM_NOT_FIXED;
}
}  // namespace scvx
