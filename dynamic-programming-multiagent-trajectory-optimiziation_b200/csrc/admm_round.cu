// admm_round.cu -- everything an ADMM consensus round needs between the position exchange and the sub-problem solve, in the
// solver's own layout, as ONE kernel (plus the neighbour selection for the k-nearest-neighbour coupling).
//
// Replaces, for the local agents i of a rank and all their neighbour slots at once:
//   * MultiAgentModel.linearize_collision / SI_MultiAgentModel.linearize_inter_agent_collision
//     (SCvx/models/multi_agent_model.py:61-79, SI_multi_agent_model.py:49-74): a_ijk = (p_ik - p_jk) / (||.|| + 1e-6);
//   * the right-hand side of AgentSolver's collision rows, a_ijk.(p_ik - Y_jk) + S_jk >= d_min  <=>  a.p + S >= d_min + a.Y_jk
//     (SCvx/optimization/agent_solver.py:79-90);
//   * the augmented-Lagrangian terms of agent_solver.py:92-95 collapsed over the neighbours:
//     sum_j <Lambda_j, P - Y_j> + rho/2 ||P - Y_j||^2 = rho n_act / 2 ||P||^2 + <sum_j Lambda_j - rho sum_j Y_j, P> + const,
//     const = rho/2 sum ||Y_j||^2 - sum <Lambda_j, Y_j>.
// Round 1 did the last two with eager PyTorch (an (n_local, N, d, K) broadcast product, four einsum/cuBLAS calls, topk) that
// re-read the N^2 table the linearisation kernel had just written.
#include "common.cuh"
#include "reduce.cuh"

namespace scvx {

constexpr double EPS_NORMAL_ADMM = 1e-6;     // the "+1e-6" of every normal's denominator (multi_agent_model.py:74)

// One block per local agent, threads strided over the K nodes, a sequential walk over the neighbour slots (every load and
// store coalesced over k).  nbr_idx == NULL: slot q is agent q (all pairs, n_slots == n_agents), else slot q is agent
// nbr_idx[i][q] (-1 = empty).  The own slot and slots with mask_in == 0 are inactive: zero rows, mask_out 0.
template <int D>
__global__ void __launch_bounds__(128)
admm_prep_kernel(int i0, int n_agents, int K, int n_x, int n_slots, double d_min, double rho, const double* __restrict__ X_own,
                 const double* __restrict__ X_all, const double* __restrict__ Y, const double* __restrict__ Lam,
                 const int* __restrict__ nbr_idx, const unsigned char* __restrict__ mask_in, double* __restrict__ col_a,
                 double* __restrict__ col_b, unsigned char* __restrict__ mask_out, double* __restrict__ lin_p,
                 double* __restrict__ quad_rho, double* __restrict__ aug_const) {
  __shared__ double red[9 * 2];
  const int i = blockIdx.x, tid = threadIdx.x;
  double part[2] = {0.0, 0.0};      // sum ||Y_j||^2, sum <Lambda_j, Y_j> over the active neighbours (this thread's nodes)
  int n_act = 0;
  for (int k0 = 0; k0 < K; k0 += blockDim.x) {
    const int k = k0 + tid;
    const bool in = k < K;
    double p[D], sl[D];
#pragma unroll
    for (int c = 0; c < D; ++c) { p[c] = in ? X_own[((size_t)i * n_x + c) * K + k] : 0.0; sl[c] = 0.0; }
    n_act = 0;
    for (int q = 0; q < n_slots; ++q) {
      const size_t slot = (size_t)i * n_slots + q;
      const int j = nbr_idx ? nbr_idx[slot] : q;
      const bool on = j >= 0 && j != i0 + i && (!mask_in || mask_in[slot]);
      n_act += on ? 1 : 0;
      if (k0 == 0 && tid == 0) mask_out[slot] = on ? 1 : 0;
      if (!in) continue;
      if (!on) {
#pragma unroll
        for (int c = 0; c < D; ++c) col_a[(slot * D + c) * K + k] = 0.0;
        col_b[slot * K + k] = 0.0;
        continue;
      }
      double diff[D], y[D], nrm2 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double qv = X_all[((size_t)j * n_x + c) * K + k];
        y[c] = Y[((size_t)j * D + c) * K + k];
        diff[c] = p[c] - qv;
        nrm2 += diff[c] * diff[c];
      }
      const double inv = 1.0 / (sqrt(nrm2) + EPS_NORMAL_ADMM);
      double dot = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double a = diff[c] * inv;
        const double lam = Lam[((size_t)j * D + c) * K + k];
        col_a[(slot * D + c) * K + k] = a;
        dot += a * y[c];
        sl[c] += lam - rho * y[c];
        part[0] += y[c] * y[c];
        part[1] += lam * y[c];
      }
      col_b[slot * K + k] = d_min + dot;
    }
    if (in) {
#pragma unroll
      for (int c = 0; c < D; ++c) lin_p[((size_t)i * D + c) * K + k] = sl[c];
    }
  }
  const int ops[2] = {0, 0};
  block_reduce<2>(part, ops, red);
  if (tid == 0) {
    quad_rho[i] = rho * (double)n_act;
    aug_const[i] = 0.5 * rho * red[0] - red[1];
  }
}

// k_sel nearest neighbours of every local agent from the (n_local, n_agents) table of squared minimum distances: one warp per
// row, k_sel rounds of (lane-local minimum over the lane's strided entries, warp arg-min with ties to the smaller index, the
// winner is struck from the table).  Entries beyond radius2 (> 0) and the own column are never selected; unused slots get -1.
__global__ void __launch_bounds__(128)
knn_select_kernel(int n_local, int i0, int n_agents, int k_sel, double radius2, double* __restrict__ d2, int* __restrict__ nbr_idx) {
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n_local) return;
  double* r = d2 + (size_t)row * n_agents;
  if (lane == 0) r[i0 + row] = INFINITY;
  __syncwarp();
  for (int s = 0; s < k_sel; ++s) {
    double best = INFINITY;
    int arg = n_agents;
    for (int j = lane; j < n_agents; j += 32) {
      const double v = r[j];
      if (v < best) { best = v; arg = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    const bool ok = arg < n_agents && isfinite(best) && (radius2 <= 0.0 || best <= radius2);
    if (lane == 0) {
      nbr_idx[(size_t)row * k_sel + s] = ok ? arg : -1;
      if (arg < n_agents) r[arg] = INFINITY;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
radius_mask_kernel(long long n, double radius2, const double* __restrict__ d2, unsigned char* __restrict__ mask) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) mask[e] = d2[e] <= radius2 ? 1 : 0;
}

}  // namespace scvx

using namespace scvx;

extern "C" int scvx_admm_round_prep(int model_id, int n_local, int i0, int n_agents, int K, int n_slots, double d_min, double rho,
                                    const double* X_own, const double* X_all, const double* Y, const double* Lambda,
                                    const int* nbr_idx, const unsigned char* mask_in, double* col_a, double* col_b,
                                    unsigned char* mask_out, double* lin_p, double* quad_rho, double* aug_const, void* stream) {
  int nx, nu, d;
  if (!model_dims(model_id, &nx, &nu, &d)) return bad_arg("model_id");
  if (n_local < 0 || n_agents < 1 || K < 1 || i0 < 0 || n_slots < 1) return bad_arg("n_local/n_agents/K/i0/n_slots");
  if (!nbr_idx && n_slots != n_agents) return bad_arg("all-pairs mode needs n_slots == n_agents");
  if (n_local == 0) return SCVX_OK;
  if (!X_own || !X_all || !Y || !Lambda || !col_a || !col_b || !mask_out || !lin_p || !quad_rho || !aug_const) return bad_arg("null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (d == 2)
    admm_prep_kernel<2><<<n_local, 128, 0, st>>>(i0, n_agents, K, nx, n_slots, d_min, rho, X_own, X_all, Y, Lambda, nbr_idx, mask_in,
                                                 col_a, col_b, mask_out, lin_p, quad_rho, aug_const);
  else
    admm_prep_kernel<3><<<n_local, 128, 0, st>>>(i0, n_agents, K, nx, n_slots, d_min, rho, X_own, X_all, Y, Lambda, nbr_idx, mask_in,
                                                 col_a, col_b, mask_out, lin_p, quad_rho, aug_const);
  SCVX_CHECK_LAUNCH("scvx_admm_round_prep");
  return SCVX_OK;
}

extern "C" int scvx_knn_select(int n_local, int i0, int n_agents, int k_sel, double radius2, double* d2, int* nbr_idx, void* stream) {
  if (n_local < 0 || n_agents < 1 || k_sel < 1 || i0 < 0 || i0 + n_local > n_agents) return bad_arg("n_local/n_agents/k_sel/i0");
  if (n_local == 0) return SCVX_OK;
  if (!d2 || !nbr_idx) return bad_arg("null pointer");
  const long long threads = (long long)n_local * 32;
  knn_select_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n_local, i0, n_agents, k_sel, radius2, d2, nbr_idx);
  SCVX_CHECK_LAUNCH("scvx_knn_select");
  return SCVX_OK;
}

extern "C" int scvx_radius_mask(int n_local, int n_agents, double radius2, const double* d2, unsigned char* mask, void* stream) {
  if (n_local < 0 || n_agents < 1) return bad_arg("n_local/n_agents");
  if (n_local == 0) return SCVX_OK;
  if (!d2 || !mask) return bad_arg("null pointer");
  const long long n = (long long)n_local * n_agents;
  radius_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, radius2, d2, mask);
  SCVX_CHECK_LAUNCH("scvx_radius_mask");
  return SCVX_OK;
}
