// linearize.cu -- stage 2 (obstacle / inter-agent half-space linearisation), the consensus round and the
// on-device outer-loop bookkeeping.  All HBM-bound elementwise / small-reduction kernels, fp64, sm_100a;
// every global access is coalesced over k (the fastest index of every trajectory array).
#include "common.cuh"

namespace scvx {

constexpr double EPS_NORMAL = 1e-6;   // the "+1e-6" in the reference's normals (unicycle_model.py:111)

// a_jk = (p_k - c_j)/(||p_k - c_j|| + 1e-6), b_jk = clearance_j + a_jk.c_j      (unicycle_model.py:103-114)
template <int D>
__global__ void __launch_bounds__(256)
linearize_obstacles_kernel(int n_agents, int K, int M, int n_x, const double* __restrict__ X_ref,
                           const double* __restrict__ obs_c, const double* __restrict__ obs_clear,
                           double* __restrict__ obs_a, double* __restrict__ obs_b) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_agents * M * K;
  if (gid >= total) return;
  const int k = (int)(gid % K);
  const long long am = gid / K;          // agent*M + j
  const int agent = (int)(am / M);
  const double* Xa = X_ref + (size_t)agent * n_x * K;
  const double* c = obs_c + (size_t)am * D;
  double diff[D], nrm2 = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    diff[i] = Xa[(size_t)i * K + k] - c[i];
    nrm2 += diff[i] * diff[i];
  }
  const double inv = 1.0 / (sqrt(nrm2) + EPS_NORMAL);
  double dot = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const double a = diff[i] * inv;
    obs_a[((size_t)am * D + i) * K + k] = a;
    dot += a * c[i];
  }
  obs_b[(size_t)am * K + k] = obs_clear[am] + dot;
}

// a_ijk = (p_ik - q_jk)/(|| || + 1e-6), b_ijk = d_min + a_ijk.q_jk             (multi_agent_model.py:61-79)
// grid: (ceil(K/32), n_agents_nbr tiles, n_local); each thread handles one k for a tile of neighbours so the
// own position is loaded once and reused.
template <int D>
__global__ void __launch_bounds__(128)
linearize_collision_kernel(int n_local, int i0, int n_agents, int K, int n_x, double d_min,
                           const double* __restrict__ X_own, const double* __restrict__ X_nbr,
                           double* __restrict__ col_a, double* __restrict__ col_b) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.z;
  if (k >= K) return;
  double p[D];
#pragma unroll
  for (int c = 0; c < D; ++c) p[c] = X_own[((size_t)i * n_x + c) * K + k];
  const int j0 = blockIdx.y * 16;
  const int j1 = min(j0 + 16, n_agents);
  for (int j = j0; j < j1; ++j) {
    const size_t slot = (size_t)i * n_agents + j;
    if (j == i0 + i) {
#pragma unroll
      for (int c = 0; c < D; ++c) col_a[(slot * D + c) * K + k] = 0.0;
      col_b[slot * K + k] = 0.0;
      continue;
    }
    double q[D], diff[D], nrm2 = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      q[c] = X_nbr[((size_t)j * n_x + c) * K + k];
      diff[c] = p[c] - q[c];
      nrm2 += diff[c] * diff[c];
    }
    const double inv = 1.0 / (sqrt(nrm2) + EPS_NORMAL);
    double dot = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const double a = diff[c] * inv;
      col_a[(slot * D + c) * K + k] = a;
      dot += a * q[c];
    }
    col_b[slot * K + k] = d_min + dot;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum for blockDim.x = 128 (4 warps); result valid in every thread
__device__ __forceinline__ double block_sum128(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return sh[0] + sh[1] + sh[2] + sh[3];
}
__device__ __forceinline__ double block_max128(double v, double* sh) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmax(fmax(sh[0], sh[1]), fmax(sh[2], sh[3]));
}

// Y+ = (Y+P)/2; Lambda += rho (P - Y+); residual norms per agent      (admm_coordinator.py:80-96)
__global__ void __launch_bounds__(128)
consensus_kernel(int dK, size_t p_stride, double rho, const double* __restrict__ P, double* __restrict__ Y,
                 double* __restrict__ Lambda, double* __restrict__ pr, double* __restrict__ du) {
  __shared__ double sh[4];
  const int j = blockIdx.x;
  const size_t base = (size_t)j * dK;
  double spr = 0.0, sdu = 0.0;
  for (int e = threadIdx.x; e < dK; e += blockDim.x) {
    const double p = P[(size_t)j * p_stride + e], yo = Y[base + e];
    const double yn = 0.5 * (yo + p);
    Y[base + e] = yn;
    Lambda[base + e] += rho * (p - yn);
    spr += (p - yn) * (p - yn);
    sdu += (yn - yo) * (yn - yo);
  }
  spr = block_sum128(spr, sh);
  sdu = block_sum128(sdu, sh);
  if (threadIdx.x == 0) { pr[j] = sqrt(spr); du[j] = sqrt(sdu); }
}

// One outer SCvx bookkeeping step per agent (scvx_solver.py:82-111, :125-133).
__global__ void __launch_bounds__(128)
outer_update_kernel(int K, int n_x, int n_u, int M, double conv_tol, const double* __restrict__ X_new,
                    const double* __restrict__ U_new, const double* __restrict__ nu_new,
                    const double* __restrict__ sigma_new, const double* __restrict__ s_prime,
                    double* __restrict__ X, double* __restrict__ U, double* __restrict__ sigma,
                    double* __restrict__ tr_radius, int* __restrict__ active, double* __restrict__ metrics) {
  __shared__ double sh[4];
  const int a = blockIdx.x;
  if (!active[a]) return;                      // uniform per block
  const size_t xo = (size_t)a * n_x * K, uo = (size_t)a * n_u * K, no = (size_t)a * n_x * (K - 1);
  // nu_norm = max_k sum_i |nu_ik|  (np.linalg.norm(nu, 1), scvx_solver.py:82)
  double cmax = 0.0;
  for (int k = threadIdx.x; k < K - 1; k += blockDim.x) {
    double s = 0.0;
    for (int i = 0; i < n_x; ++i) s += fabs(nu_new[no + (size_t)i * (K - 1) + k]);
    cmax = fmax(cmax, s);
  }
  const double nu_norm = block_max128(cmax, sh);
  double ssl = 0.0;
  for (int e = threadIdx.x; e < M * K; e += blockDim.x) ssl += s_prime[(size_t)a * M * K + e];
  const double slack = block_sum128(ssl, sh);
  double sdx = 0.0, sdu = 0.0;
  for (int e = threadIdx.x; e < n_x * K; e += blockDim.x) { const double d = X_new[xo + e] - X[xo + e]; sdx += d * d; }
  for (int e = threadIdx.x; e < n_u * K; e += blockDim.x) { const double d = U_new[uo + e] - U[uo + e]; sdu += d * d; }
  const double dx = sqrt(block_sum128(sdx, sh));
  const double du = sqrt(block_sum128(sdu, sh));
  const double sn = sigma_new[a];
  const double ds = fabs(sn - sigma[a]);
  __syncthreads();
  if (threadIdx.x == 0) {
    double* m = metrics + (size_t)a * 6;
    m[0] = nu_norm; m[1] = slack; m[2] = dx; m[3] = du; m[4] = ds; m[5] = sn;
  }
  const bool conv = nu_norm < conv_tol && slack < conv_tol && dx < conv_tol && ds < conv_tol;
  if (conv) {                                   // break BEFORE accepting the new iterate (scvx_solver.py:104-105)
    if (threadIdx.x == 0) active[a] = 0;
    return;
  }
  for (int e = threadIdx.x; e < n_x * K; e += blockDim.x) X[xo + e] = X_new[xo + e];
  for (int e = threadIdx.x; e < n_u * K; e += blockDim.x) U[uo + e] = U_new[uo + e];
  if (threadIdx.x == 0) {
    sigma[a] = sn;
    double tr = tr_radius[a];
    tr = (nu_norm < 1e-2 && slack < 1e-2) ? fmin(tr * 1.5, 50.0) : fmin(tr * 1.2, 50.0);
    tr_radius[a] = fmax(tr, 1e-3);
  }
}


// Slab normals of the Nash best response (GameUnicycleModel.update_slabs + the slab rows of get_cost_function,
// SCvx/models/game_model.py:56-67,118-124):  z_jk = d / ||d||, d = p_ik - q_jk  (zero when ||d|| < 1e-6, no +1e-6 in the
// denominator), row  z_jk.(p_ik - Y_jk) >= radius_i  <=>  a.p >= b  with a = z, b = radius_i + z.Y_jk.
// degenerate[i] counts the (j, k) whose normal vanished: the reference's row 0 >= radius is infeasible there.
template <int D>
__global__ void __launch_bounds__(128)
slab_normals_kernel(int n_local, int i0, int n_agents, int K, int n_x, const double* __restrict__ radius,
                    const double* __restrict__ P_own, const double* __restrict__ X_dir, const double* __restrict__ X_off,
                    double* __restrict__ col_a, double* __restrict__ col_b, int* __restrict__ degenerate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.z;
  if (k >= K) return;
  double p[D];
#pragma unroll
  for (int c = 0; c < D; ++c) p[c] = P_own[((size_t)i * n_x + c) * K + k];
  const int j0 = blockIdx.y * 16;
  const int j1 = min(j0 + 16, n_agents);
  int bad = 0;
  for (int j = j0; j < j1; ++j) {
    const size_t slot = (size_t)i * n_agents + j;
    double z[D], n2 = 0.0, zy = 0.0;
    if (j == i0 + i) {
#pragma unroll
      for (int c = 0; c < D; ++c) col_a[(slot * D + c) * K + k] = 0.0;
      col_b[slot * K + k] = 0.0;
      continue;
    }
#pragma unroll
    for (int c = 0; c < D; ++c) { z[c] = p[c] - X_dir[((size_t)j * n_x + c) * K + k]; n2 = (c == 0) ? __dmul_rn(z[c], z[c]) : __dadd_rn(n2, __dmul_rn(z[c], z[c])); }
    const double nrm = sqrt(n2);
    const bool zero = nrm < 1e-6;
    bad += zero ? 1 : 0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      z[c] = zero ? 0.0 : z[c] / nrm;
      col_a[(slot * D + c) * K + k] = z[c];
      zy += z[c] * X_off[((size_t)j * n_x + c) * K + k];
    }
    col_b[slot * K + k] = radius[i] + zy;
  }
  if (bad && degenerate) atomicAdd(degenerate + i, bad);
}


// ---- neighbour culling for large N (a documented DEVIATION from the reference, which couples all pairs) ----------------
// cross_min_dist2_kernel: one warp per (local agent i, agent j): min over k of || p_ik - q_jk ||^2 on the position rows.
// linearize_collision_indexed_kernel: the half-spaces of multi_agent_model.py:61-79 for a per-agent LIST of neighbours
// (nbr_idx [n_local][n_sel], -1 = empty slot): compact tables [n_local][n_sel][d][K] instead of [n_local][N][d][K].
template <int D>
__global__ void __launch_bounds__(128)
cross_min_dist2_kernel(int n_local, int n_agents, int K, int n_x, const double* __restrict__ X_own,
                       const double* __restrict__ X_all, double* __restrict__ d2) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)n_local * n_agents) return;
  const int i = (int)(w / n_agents), j = (int)(w - (long long)i * n_agents);
  const double* Xi = X_own + (size_t)i * n_x * K;
  const double* Xj = X_all + (size_t)j * n_x * K;
  double m = INFINITY;
  for (int k = lane; k < K; k += 32) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) { const double d = Xi[(size_t)c * K + k] - Xj[(size_t)c * K + k]; s += d * d; }
    m = fmin(m, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) d2[w] = m;
}

template <int D>
__global__ void __launch_bounds__(128)
linearize_collision_indexed_kernel(int n_local, int n_sel, int K, int n_x, double d_min, const double* __restrict__ X_own,
                                   const double* __restrict__ X_all, const int* __restrict__ nbr_idx,
                                   double* __restrict__ col_a, double* __restrict__ col_b) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.z;
  if (k >= K) return;
  double p[D];
#pragma unroll
  for (int c = 0; c < D; ++c) p[c] = X_own[((size_t)i * n_x + c) * K + k];
  const int q0 = blockIdx.y * 16;
  const int q1 = min(q0 + 16, n_sel);
  for (int q = q0; q < q1; ++q) {
    const size_t slot = (size_t)i * n_sel + q;
    const int j = nbr_idx[slot];
    if (j < 0) {
#pragma unroll
      for (int c = 0; c < D; ++c) col_a[(slot * D + c) * K + k] = 0.0;
      col_b[slot * K + k] = 0.0;
      continue;
    }
    double qq[D], diff[D], nrm2 = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      qq[c] = X_all[((size_t)j * n_x + c) * K + k];
      diff[c] = p[c] - qq[c];
      nrm2 += diff[c] * diff[c];
    }
    const double inv = 1.0 / (sqrt(nrm2) + EPS_NORMAL);
    double dot = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const double a = diff[c] * inv;
      col_a[(slot * D + c) * K + k] = a;
      dot += a * qq[c];
    }
    col_b[slot * K + k] = d_min + dot;
  }
}

}  // namespace scvx

using namespace scvx;

extern "C" int scvx_linearize_obstacles_batched(int model_id, int n_agents, int K, int M, const double* X_ref,
                                                const double* obs_c, const double* obs_clear, double* obs_a,
                                                double* obs_b, void* stream) {
  int nx, nu, d;
  if (!model_dims(model_id, &nx, &nu, &d)) return bad_arg("model_id");
  if (n_agents < 0 || K < 1 || M < 0) return bad_arg("n_agents/K/M");
  if (n_agents == 0 || M == 0) return SCVX_OK;
  if (!X_ref || !obs_c || !obs_clear || !obs_a || !obs_b) return bad_arg("null pointer");
  const long long total = (long long)n_agents * M * K;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)stream;
  if (d == 2)
    linearize_obstacles_kernel<2><<<blocks, threads, 0, st>>>(n_agents, K, M, nx, X_ref, obs_c, obs_clear, obs_a, obs_b);
  else
    linearize_obstacles_kernel<3><<<blocks, threads, 0, st>>>(n_agents, K, M, nx, X_ref, obs_c, obs_clear, obs_a, obs_b);
  SCVX_CHECK_LAUNCH("scvx_linearize_obstacles_batched");
  return SCVX_OK;
}

extern "C" int scvx_linearize_collision_batched(int model_id, int n_local, int i0, int n_agents, int K, double d_min,
                                                const double* X_own, const double* X_nbr, double* col_a,
                                                double* col_b, void* stream) {
  int nx, nu, d;
  if (!model_dims(model_id, &nx, &nu, &d)) return bad_arg("model_id");
  if (n_local < 0 || n_agents < 0 || K < 1 || i0 < 0) return bad_arg("n_local/n_agents/K/i0");
  if (n_local == 0 || n_agents == 0) return SCVX_OK;
  if (n_local > 65535) return bad_arg("n_local > 65535 (split the call)");
  if (!X_own || !X_nbr || !col_a || !col_b) return bad_arg("null pointer");
  dim3 grid((K + 127) / 128, (n_agents + 15) / 16, n_local);
  cudaStream_t st = (cudaStream_t)stream;
  if (d == 2)
    linearize_collision_kernel<2><<<grid, 128, 0, st>>>(n_local, i0, n_agents, K, nx, d_min, X_own, X_nbr, col_a, col_b);
  else
    linearize_collision_kernel<3><<<grid, 128, 0, st>>>(n_local, i0, n_agents, K, nx, d_min, X_own, X_nbr, col_a, col_b);
  SCVX_CHECK_LAUNCH("scvx_linearize_collision_batched");
  return SCVX_OK;
}

extern "C" int scvx_slab_normals_batched(int model_id, int n_local, int i0, int n_agents, int K, const double* radius,
                                         const double* P_own, const double* X_dir, const double* X_off, double* col_a,
                                         double* col_b, int* degenerate, void* stream) {
  int nx, nu, d;
  if (!model_dims(model_id, &nx, &nu, &d)) return bad_arg("model_id");
  if (n_local < 0 || n_agents < 0 || K < 1 || i0 < 0) return bad_arg("n_local/n_agents/K/i0");
  if (n_local == 0 || n_agents == 0) return SCVX_OK;
  if (n_local > 65535) return bad_arg("n_local > 65535 (split the call)");
  if (!radius || !P_own || !X_dir || !X_off || !col_a || !col_b) return bad_arg("null pointer");
  dim3 grid((K + 127) / 128, (n_agents + 15) / 16, n_local);
  cudaStream_t st = (cudaStream_t)stream;
  if (degenerate) {
    cudaError_t e = cudaMemsetAsync(degenerate, 0, (size_t)n_local * sizeof(int), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
  }
  if (d == 2)
    slab_normals_kernel<2><<<grid, 128, 0, st>>>(n_local, i0, n_agents, K, nx, radius, P_own, X_dir, X_off, col_a, col_b, degenerate);
  else
    slab_normals_kernel<3><<<grid, 128, 0, st>>>(n_local, i0, n_agents, K, nx, radius, P_own, X_dir, X_off, col_a, col_b, degenerate);
  SCVX_CHECK_LAUNCH("scvx_slab_normals_batched");
  return SCVX_OK;
}

extern "C" int scvx_cross_min_dist2(int model_id, int n_local, int n_agents, int K, const double* X_own, const double* X_all,
                                    double* d2, void* stream) {
  int nx, nu, d;
  if (!model_dims(model_id, &nx, &nu, &d)) return bad_arg("model_id");
  if (n_local < 0 || n_agents < 0 || K < 1) return bad_arg("n_local/n_agents/K");
  const long long items = (long long)n_local * n_agents;
  if (items == 0) return SCVX_OK;
  if (!X_own || !X_all || !d2) return bad_arg("null pointer");
  const unsigned blocks = (unsigned)((items + 3) / 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (d == 2) cross_min_dist2_kernel<2><<<blocks, 128, 0, st>>>(n_local, n_agents, K, nx, X_own, X_all, d2);
  else cross_min_dist2_kernel<3><<<blocks, 128, 0, st>>>(n_local, n_agents, K, nx, X_own, X_all, d2);
  SCVX_CHECK_LAUNCH("scvx_cross_min_dist2");
  return SCVX_OK;
}

extern "C" int scvx_linearize_collision_indexed(int model_id, int n_local, int n_sel, int n_agents, int K, double d_min,
                                                const double* X_own, const double* X_all, const int* nbr_idx, double* col_a,
                                                double* col_b, void* stream) {
  int nx, nu, d;
  if (!model_dims(model_id, &nx, &nu, &d)) return bad_arg("model_id");
  if (n_local < 0 || n_sel < 0 || n_agents < 0 || K < 1) return bad_arg("n_local/n_sel/n_agents/K");
  if (n_local == 0 || n_sel == 0) return SCVX_OK;
  if (n_local > 65535) return bad_arg("n_local > 65535 (split the call)");
  if (!X_own || !X_all || !nbr_idx || !col_a || !col_b) return bad_arg("null pointer");
  dim3 grid((K + 127) / 128, (n_sel + 15) / 16, n_local);
  cudaStream_t st = (cudaStream_t)stream;
  if (d == 2)
    linearize_collision_indexed_kernel<2><<<grid, 128, 0, st>>>(n_local, n_sel, K, nx, d_min, X_own, X_all, nbr_idx, col_a, col_b);
  else
    linearize_collision_indexed_kernel<3><<<grid, 128, 0, st>>>(n_local, n_sel, K, nx, d_min, X_own, X_all, nbr_idx, col_a, col_b);
  SCVX_CHECK_LAUNCH("scvx_linearize_collision_indexed");
  return SCVX_OK;
}

extern "C" int scvx_consensus_update(int n_agents, int d, int K, double rho, const double* P, double* Y,
                                     double* Lambda, double* pr, double* du, void* stream) {
  if (n_agents < 0 || d < 1 || K < 1) return bad_arg("n_agents/d/K");
  if (n_agents == 0) return SCVX_OK;
  if (!P || !Y || !Lambda || !pr || !du) return bad_arg("null pointer");
  consensus_kernel<<<n_agents, 128, 0, (cudaStream_t)stream>>>(d * K, (size_t)d * K, rho, P, Y, Lambda, pr, du);
  SCVX_CHECK_LAUNCH("scvx_consensus_update");
  return SCVX_OK;
}

extern "C" int scvx_consensus_update_x(int n_agents, int n_x, int d, int K, double rho, const double* X, double* Y, double* Lambda,
                                       double* pr, double* du, void* stream) {
  if (n_agents < 0 || d < 1 || K < 1 || n_x < d) return bad_arg("n_agents/n_x/d/K");
  if (n_agents == 0) return SCVX_OK;
  if (!X || !Y || !Lambda || !pr || !du) return bad_arg("null pointer");
  // positions = the first d rows of every agent's (n_x, K) state block: read in place, no packed copy
  consensus_kernel<<<n_agents, 128, 0, (cudaStream_t)stream>>>(d * K, (size_t)n_x * K, rho, X, Y, Lambda, pr, du);
  SCVX_CHECK_LAUNCH("scvx_consensus_update_x");
  return SCVX_OK;
}

extern "C" int scvx_outer_update(int model_id, int n_agents, int K, int M, double conv_tol, const double* X_new,
                                 const double* U_new, const double* nu_new, const double* sigma_new,
                                 const double* s_prime, double* X, double* U, double* sigma, double* tr_radius,
                                 int* active, double* metrics, void* stream) {
  int nx, nu, d;
  if (!model_dims(model_id, &nx, &nu, &d)) return bad_arg("model_id");
  if (n_agents < 0 || K < 2 || M < 0) return bad_arg("n_agents/K/M");
  if (n_agents == 0) return SCVX_OK;
  if (!X_new || !U_new || !nu_new || !sigma_new || !X || !U || !sigma || !tr_radius || !active || !metrics ||
      (M > 0 && !s_prime))
    return bad_arg("null pointer");
  outer_update_kernel<<<n_agents, 128, 0, (cudaStream_t)stream>>>(K, nx, nu, M, conv_tol, X_new, U_new, nu_new,
                                                                  sigma_new, s_prime, X, U, sigma, tr_radius, active,
                                                                  metrics);
  SCVX_CHECK_LAUNCH("scvx_outer_update");
  return SCVX_OK;
}
