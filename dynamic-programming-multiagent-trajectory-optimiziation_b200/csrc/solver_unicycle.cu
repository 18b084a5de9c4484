// solver_unicycle.cu -- the sub-problem kernels (solver_kernel.cuh) instantiated for the Unicycle shape: a translation unit of its own
// so that the models' kernels compile in parallel.
#include "solver_kernel.cuh"

namespace scvx {
int launch_ipm_unicycle(const scvx_solve_args& a, cudaStream_t st) { return launch_ipm<Unicycle>(a, st); }
unsigned long long solver_ws_total_doubles_unicycle(int n_agents, int K, int NH) { return (unsigned long long)solver_ws_total_doubles<Unicycle>(n_agents, K, NH); }
int phase_cycles_unicycle(unsigned long long* out32, int reset) { return phase_cycles_of_this_unit(out32, reset); }
}  // namespace scvx
