// solver.cu -- stage 3 entry points: SCProblem.solve / AgentSolver.solve for a batch (SCvx/optimization/sc_problem.py:15-105,
// agent_solver.py:43-117).  The kernel is solver_kernel.cuh, instantiated per model shape in solver_unicycle.cu / solver_si.cu.
#include "common.cuh"

namespace scvx {
int launch_ipm_unicycle(const scvx_solve_args& a, cudaStream_t st);
int launch_ipm_si(const scvx_solve_args& a, cudaStream_t st);
unsigned long long solver_ws_total_doubles_unicycle(int n_agents, int K, int NH);
unsigned long long solver_ws_total_doubles_si(int n_agents, int K, int NH);
int phase_cycles_unicycle(unsigned long long* out32, int reset);
int phase_cycles_si(unsigned long long* out32, int reset);
}  // namespace scvx

using namespace scvx;

// Debug aid (tools/phase_timing.py): cycles block 0 spent per kernel phase since the last call; all zero unless the
// library was built with -DSCVX_PHASE_TIMING.
extern "C" int scvx_debug_phase_cycles(unsigned long long* out32, int reset) {
  unsigned long long u[32], v[32];
  int rc = phase_cycles_unicycle(u, reset);
  if (rc != SCVX_OK) return rc;
  rc = phase_cycles_si(v, reset);
  if (rc != SCVX_OK) return rc;
  for (int i = 0; i < 32; ++i) out32[i] = u[i] + v[i];
  return SCVX_OK;
}

// rank of agent i in the longest-first order = #{j : iters_j > iters_i} + #{j < i : iters_j == iters_i}; order[rank] = i.
__global__ void __launch_bounds__(256) order_by_iters_kernel(int n, const int* __restrict__ iters, int* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int mine = iters[i];
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    const int v = iters[j];
    rank += (v > mine || (v == mine && j < i)) ? 1 : 0;
  }
  order[rank] = i;
}

__global__ void __launch_bounds__(256) mu0_from_iters_kernel(int n, const int* __restrict__ iters, int easy_max, double mu_easy,
                                                             double mu_hard, double* __restrict__ mu0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mu0[i] = (iters[i] > 0 && iters[i] <= easy_max) ? mu_easy : mu_hard;
}

extern "C" int scvx_mu0_from_iters(int n_agents, const int* iters, int easy_max_iters, double mu0_easy, double mu0_hard, double* mu0,
                                   void* stream) {
  if (n_agents < 0) return bad_arg("n_agents");
  if (n_agents == 0) return SCVX_OK;
  if (!iters || !mu0) return bad_arg("null pointer");
  if (!(mu0_easy > 0.0) || !(mu0_hard > 0.0)) return bad_arg("mu0 values must be positive");
  mu0_from_iters_kernel<<<(n_agents + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_agents, iters, easy_max_iters, mu0_easy, mu0_hard, mu0);
  SCVX_CHECK_LAUNCH("scvx_mu0_from_iters");
  return SCVX_OK;
}

extern "C" int scvx_order_by_iters(int n_agents, const int* iters, int* order, void* stream) {
  if (n_agents < 0) return bad_arg("n_agents");
  if (n_agents == 0) return SCVX_OK;
  if (!iters || !order) return bad_arg("null pointer");
  order_by_iters_kernel<<<(n_agents + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_agents, iters, order);
  SCVX_CHECK_LAUNCH("scvx_order_by_iters");
  return SCVX_OK;
}

extern "C" unsigned long long scvx_solve_workspace_bytes(int model_id, int n_agents, int K, int M, int n_nbr) {
  if (n_agents < 0 || K < 3 || M < 0 || n_nbr < 0) return 0ull;
  size_t tot;
  switch (solver_shape_of(model_id)) {
    case SCVX_MODEL_UNICYCLE: tot = (size_t)solver_ws_total_doubles_unicycle(n_agents, K, M + n_nbr); break;
    case SCVX_MODEL_SINGLE_INTEGRATOR: tot = (size_t)solver_ws_total_doubles_si(n_agents, K, M + n_nbr); break;
    default: return 0ull;
  }
  return (unsigned long long)tot * sizeof(double);
}

extern "C" int scvx_solve_batched(const scvx_solve_args* a, void* stream) {
  if (!a) return bad_arg("args");
  if (a->n_agents < 0 || a->K < 3 || a->M < 0 || a->n_nbr < 0) return bad_arg("n_agents/K/M/n_nbr");
  if (a->n_agents == 0) return SCVX_OK;
  if (!a->norm1_induced) {
    snprintf(g_last_error, sizeof(g_last_error), "entry-wise L1 mode is not implemented (the reference uses the induced norm)");
    return SCVX_E_UNSUPPORTED;
  }
  if (!a->A_bar || !a->B_bar || !a->C_bar || !a->S_bar || !a->z_bar || !a->X_ref || !a->U_ref || !a->sigma_ref ||
      !a->tr_radius || !a->x_init || !a->x_final || !a->pos_lo || !a->pos_hi || !a->v_max || !a->X || !a->U || !a->nu ||
      !a->sigma || !a->objective || !a->status || !a->iters)
    return bad_arg("null pointer");
  if (a->M > 0 && (!a->obs_a || !a->obs_b || !a->s_prime)) return bad_arg("obstacle tables");
  if (a->n_nbr > 0 && (!a->col_a || !a->col_b)) return bad_arg("collision tables");
  const unsigned long long need = scvx_solve_workspace_bytes(a->model_id, a->n_agents, a->K, a->M, a->n_nbr);
  if (need == 0ull) return bad_arg("model_id");
  if (!a->workspace || a->workspace_bytes < need) {
    snprintf(g_last_error, sizeof(g_last_error), "workspace too small: need %llu bytes", need);
    return SCVX_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (solver_shape_of(a->model_id)) {     // a registered user model runs through the kernel of its shape
    case SCVX_MODEL_UNICYCLE:
      if (!a->w_max) return bad_arg("w_max");
      return launch_ipm_unicycle(*a, st);
    case SCVX_MODEL_SINGLE_INTEGRATOR:
      return launch_ipm_si(*a, st);
    default:
      return bad_arg("model_id");
  }
}
