// utils.cu -- the callers either side of the SCvx loop (SURVEY section 8 f, ranks 2 and 4), batched, fp64, sm_100a:
//   * warm-start generators        SCvx/utils/initial_guess.py:61-107 (unicycle, tangent way-points around discs)
//                                   SCvx/utils/IS_initial_guess.py:87-126 (single integrator, detour way-points around spheres)
//   * analysis metrics             SCvx/utils/analysis.py:10-31 (min_inter_agent_distance), :34-62 (min_agent_obstacle_distance)
// The way-point list of an agent is tiny and sequential (obstacle order matters): thread 0 of the agent's block builds it in
// shared memory, then all threads sample the K points of the piecewise-linear path with numpy.linspace's exact arithmetic.
// Sums of squares use __dmul_rn/__dadd_rn in numpy's order (no FMA contraction), so everything that does not go through
// libm's asin/atan2/sincos reproduces the reference bit for bit.
#include "common.cuh"

namespace scvx {

__device__ __forceinline__ double sq_sum2(double a, double b) { return __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)); }
__device__ __forceinline__ double sq_sum3(double a, double b, double c) { return __dadd_rn(sq_sum2(a, b), __dmul_rn(c, c)); }

template <int D>
__device__ __forceinline__ double dot_d(const double* a, const double* b) {
  double s = __dmul_rn(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < D; ++i) s = __dadd_rn(s, __dmul_rn(a[i], b[i]));
  return s;
}

// line_circle_intersect / line_sphere_intersect (initial_guess.py:4-19, IS_initial_guess.py:6-23)
template <int D>
__device__ bool segment_hits_ball(const double* p, const double* q, const double* c, double r) {
  double d[D], f[D];
#pragma unroll
  for (int i = 0; i < D; ++i) { d[i] = q[i] - p[i]; f[i] = p[i] - c[i]; }
  const double a = dot_d<D>(d, d);
  const double b = __dmul_rn(2.0, dot_d<D>(f, d));
  const double cc = __dadd_rn(dot_d<D>(f, f), -__dmul_rn(r, r));
  const double disc = __dadd_rn(__dmul_rn(b, b), -__dmul_rn(__dmul_rn(4.0, a), cc));
  if (disc < 0.0) return false;
  const double sq = sqrt(disc);
  const double t1 = (-b + sq) / __dmul_rn(2.0, a);
  const double t2 = (-b - sq) / __dmul_rn(2.0, a);
  return (0.0 < t1 && t1 < 1.0) || (0.0 < t2 && t2 < 1.0);
}

// compute_tangent_points (initial_guess.py:22-38); false when p is inside / on the circle (the reference raises)
__device__ bool tangent_points(const double* p, const double* c, double r, double* T1, double* T2) {
  const double v0 = p[0] - c[0], v1 = p[1] - c[1];
  const double d = sqrt(sq_sum2(v0, v1));
  if (d <= r) return false;
  const double alpha = asin(r / d);
  const double theta = atan2(v1, v0);
  const double t1 = theta + alpha, t2 = theta - alpha;
  T1[0] = c[0] + __dmul_rn(r, cos(t1)); T1[1] = c[1] + __dmul_rn(r, sin(t1));
  T2[0] = c[0] + __dmul_rn(r, cos(t2)); T2[1] = c[1] + __dmul_rn(r, sin(t2));
  return true;
}

__device__ __forceinline__ double dist2(const double* a, const double* b) { return sqrt(sq_sum2(a[0] - b[0], a[1] - b[1])); }

// One agent per block.  smem: pts [(2 M_max + 2)][D], seg_n [2 M_max + 1] ints, seg_off likewise.
// status: 0 ok; 1 start/goal inside an inflated obstacle (ValueError "Point inside/on circle"); 2 negative sample count for the
// last segment (numpy's ValueError in linspace); 3 start and goal closer than 1e-6 (ValueError in compute_detour_waypoints).
template <int D>
__global__ void __launch_bounds__(128)
warm_start_kernel(int K, int M_max, const double* __restrict__ p0_all, const double* __restrict__ p1_all,
                  const double* __restrict__ obs_c, const double* __restrict__ obs_r, const int* __restrict__ obs_count,
                  double clearance, double* __restrict__ X0, double* __restrict__ U0, int* __restrict__ status) {
  constexpr int NU = (D == 2) ? 2 : 3;
  extern __shared__ __align__(16) double sm[];
  const int agent = blockIdx.x, tid = threadIdx.x;
  const int P_max = 2 * M_max + 2;
  double* pts = sm;                                   // [P_max][D]
  int* seg_n = (int*)(pts + (size_t)P_max * D);       // [P_max - 1]
  int* seg_off = seg_n + P_max;                       // [P_max]
  __shared__ int s_np, s_status;

  if (tid == 0) {
    const double* p0 = p0_all + (size_t)agent * 3;
    const double* p1 = p1_all + (size_t)agent * 3;
    const int M = obs_count ? obs_count[agent] : M_max;
    int np = 0, st = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) pts[i] = p0[i];
    np = 1;
    for (int j = 0; j < M && st == 0; ++j) {
      const double* c = obs_c + ((size_t)agent * M_max + j) * D;
      const double r = obs_r[(size_t)agent * M_max + j] + clearance;
      if (!segment_hits_ball<D>(p0, p1, c, r)) continue;
      if (D == 2) {
        double T[2][2], G[2][2];
        if (!tangent_points(p0, c, r, T[0], T[1]) || !tangent_points(p1, c, r, G[0], G[1])) { st = 1; break; }
        double best = 0.0; int bi = 0, bj = 0; bool first = true;
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) {
            const double L = __dadd_rn(__dadd_rn(dist2(p0, T[a]), dist2(T[a], G[b])), dist2(G[b], p1));
            if (first || L < best) { best = L; bi = a; bj = b; first = false; }
          }
        pts[np * D + 0] = T[bi][0]; pts[np * D + 1] = T[bi][1]; ++np;
        pts[np * D + 0] = G[bj][0]; pts[np * D + 1] = G[bj][1]; ++np;
      } else {
        // compute_detour_waypoints (IS_initial_guess.py:26-55)
        double d[3], du[3], u[3];
        for (int i = 0; i < 3; ++i) d[i] = p1[i] - p0[i];
        const double nd = sqrt(sq_sum3(d[0], d[1], d[2]));
        if (nd < 1e-6) { st = 3; break; }
        for (int i = 0; i < 3; ++i) du[i] = d[i] / nd;
        double tmp[3] = {1.0, 0.0, 0.0};
        if (!(fabs(du[0]) < 0.9)) { tmp[0] = 0.0; tmp[1] = 1.0; }
        u[0] = __dadd_rn(__dmul_rn(du[1], tmp[2]), -__dmul_rn(du[2], tmp[1]));
        u[1] = __dadd_rn(__dmul_rn(du[2], tmp[0]), -__dmul_rn(du[0], tmp[2]));
        u[2] = __dadd_rn(__dmul_rn(du[0], tmp[1]), -__dmul_rn(du[1], tmp[0]));
        const double nu_ = sqrt(sq_sum3(u[0], u[1], u[2]));
        for (int i = 0; i < 3; ++i) u[i] = u[i] / nu_;
        double cp[3];
        for (int i = 0; i < 3; ++i) cp[i] = c[i] - p0[i];
        const double t = dot_d<3>(cp, du);
        for (int i = 0; i < 3; ++i) {
          const double proj = __dadd_rn(p0[i], __dmul_rn(t, du[i]));
          pts[np * D + i] = __dadd_rn(proj, __dmul_rn(r, u[i]));
          pts[(np + 1) * D + i] = __dadd_rn(proj, -__dmul_rn(r, u[i]));
        }
        np += 2;
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) pts[np * D + i] = p1[i];
    ++np;
    if (st == 0) {
      // generate_piecewise_linear: samples per segment (initial_guess.py:41-58)
      const int ns = np - 1;
      double total = 0.0;
      for (int s = 0; s < ns; ++s) {
        double l2 = 0.0;
        if (D == 2) l2 = sq_sum2(pts[(s + 1) * D] - pts[s * D], pts[(s + 1) * D + 1] - pts[s * D + 1]);
        else l2 = sq_sum3(pts[(s + 1) * D] - pts[s * D], pts[(s + 1) * D + 1] - pts[s * D + 1], pts[(s + 1) * D + 2] - pts[s * D + 2]);
        total = __dadd_rn(total, sqrt(l2));
      }
      int used = 0;
      for (int s = 0; s < ns; ++s) {
        double l2 = 0.0;
        if (D == 2) l2 = sq_sum2(pts[(s + 1) * D] - pts[s * D], pts[(s + 1) * D + 1] - pts[s * D + 1]);
        else l2 = sq_sum3(pts[(s + 1) * D] - pts[s * D], pts[(s + 1) * D + 1] - pts[s * D + 1], pts[(s + 1) * D + 2] - pts[s * D + 2]);
        const double q = __dmul_rn((double)K, sqrt(l2)) / total;
        int n = (int)rint(q);                    // Python's round(): half to even
        if (n < 2) n = 2;
        if (s == ns - 1) n = K - used;
        seg_off[s] = used; seg_n[s] = n;
        used += n;
      }
      seg_off[ns] = used;
      if (seg_n[ns - 1] < 0) st = 2;
    }
    s_np = np; s_status = st;
    status[agent] = st;
  }
  __syncthreads();
  double* Xa = X0 + (size_t)agent * 3 * K;
  double* Ua = U0 + (size_t)agent * NU * K;
  if (s_status != 0) {
    for (int k = tid; k < K; k += blockDim.x) {
      for (int i = 0; i < 3; ++i) Xa[(size_t)i * K + k] = 0.0;
      for (int i = 0; i < NU; ++i) Ua[(size_t)i * K + k] = 0.0;
    }
    return;
  }
  const int ns = s_np - 1;
  // numpy.linspace(start, stop, num, endpoint) evaluated at global sample index k
  auto sample = [&](int k, double* out) {
    int s = 0;
    while (s < ns - 1 && k >= seg_off[s + 1]) ++s;
    const int i = k - seg_off[s], num = seg_n[s];
    const bool endpoint = (s == ns - 1);
    const int div = endpoint ? num - 1 : num;
    const double* a = pts + s * D;
    const double* b = pts + (s + 1) * D;
    double delta[D], step[D];
    bool any_zero = false;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      delta[c] = b[c] - a[c];
      step[c] = (div > 0) ? delta[c] / (double)div : 0.0;
      any_zero = any_zero || (step[c] == 0.0);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
      double y;
      if (div > 0) y = any_zero ? __dmul_rn((double)i / (double)div, delta[c]) : __dmul_rn((double)i, step[c]);
      else y = __dmul_rn((double)i, delta[c]);
      y = __dadd_rn(y, a[c]);
      if (endpoint && num > 1 && i == num - 1) y = b[c];
      out[c] = y;
    }
  };
  const double dt = 1.0 / (double)(K - 1);
  for (int k = tid; k < K; k += blockDim.x) {
    double pk[D], pn[D], pm[D];
    sample(k, pk);
#pragma unroll
    for (int c = 0; c < D; ++c) Xa[(size_t)c * K + k] = pk[c];
    if (D == 2) {
      // orientations from forward differences; the last one repeats (initial_guess.py:100-104); U0 = 0
      if (k < K - 1) { sample(k + 1, pn); Xa[(size_t)2 * K + k] = atan2(pn[1] - pk[1], pn[0] - pk[0]); }
      else if (K >= 2) { sample(k - 1, pm); Xa[(size_t)2 * K + k] = atan2(pk[1] - pm[1], pk[0] - pm[0]); }
      else Xa[(size_t)2 * K + k] = 0.0;
      Ua[k] = 0.0; Ua[(size_t)K + k] = 0.0;
    } else {
      // velocity warm start (IS_initial_guess.py:121-125)
      if (k < K - 1) { sample(k + 1, pn); for (int c = 0; c < 3; ++c) Ua[(size_t)c * K + k] = (pn[c] - pk[c]) / dt; }
      else if (K >= 2) { sample(k - 1, pm); for (int c = 0; c < 3; ++c) Ua[(size_t)c * K + k] = (pk[c] - pm[c]) / dt; }
      else for (int c = 0; c < 3; ++c) Ua[(size_t)c * K + k] = 0.0;
    }
  }
}

// ---- analysis metrics ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One warp per unordered pair (i < j): d_mat[i][j] = d_mat[j][i] = min_k || X_i[0:R, k] - X_j[0:R, k] ||   (analysis.py:10-31)
__global__ void __launch_bounds__(128)
min_pair_distance_kernel(int N, int K, int n_x, int R, const double* __restrict__ X, double* __restrict__ d_mat,
                         double* __restrict__ d_min) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long n_pairs = (long long)N * (N - 1) / 2;
  if (w >= n_pairs) return;
  // unrank w -> (i, j), i < j, row-major over the strict upper triangle
  int i = (int)((2.0 * N - 1.0 - sqrt((2.0 * N - 1.0) * (2.0 * N - 1.0) - 8.0 * (double)w)) * 0.5);
  while ((long long)i * (2LL * N - i - 1) / 2 > w) --i;
  while ((long long)(i + 1) * (2LL * N - i - 2) / 2 <= w) ++i;
  const int j = (int)(w - (long long)i * (2LL * N - i - 1) / 2) + i + 1;
  const double* Xi = X + (size_t)i * n_x * K;
  const double* Xj = X + (size_t)j * n_x * K;
  double m = INFINITY;
  for (int k = lane; k < K; k += 32) {
    double s = 0.0;
    for (int c = 0; c < R; ++c) {
      const double d = Xi[(size_t)c * K + k] - Xj[(size_t)c * K + k];
      s = (c == 0) ? __dmul_rn(d, d) : __dadd_rn(s, __dmul_rn(d, d));
    }
    m = fmin(m, sqrt(s));
  }
  m = warp_min(m);
  if (lane == 0) {
    d_mat[(size_t)i * N + j] = m; d_mat[(size_t)j * N + i] = m;
    if (m > 0.0) atomic_min_double(d_min, m);         // "ignore zeros" (analysis.py:29)
  }
}

// One warp per (agent, obstacle): d_mat[i][j] = min_k ( || p_i(k) - c_j || - (robot_radius + r_j) )      (analysis.py:34-62)
__global__ void __launch_bounds__(128)
min_obstacle_distance_kernel(int N, int K, int n_x, int R, int M, const double* __restrict__ X, const double* __restrict__ obs_c,
                             const double* __restrict__ obs_r, double robot_radius, double* __restrict__ d_mat,
                             double* __restrict__ d_min) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)N * M) return;
  const int i = (int)(w / M), j = (int)(w - (long long)i * M);
  const double* Xi = X + (size_t)i * n_x * K;
  const double rr = robot_radius + obs_r[j];
  double m = INFINITY;
  for (int k = lane; k < K; k += 32) {
    double s = 0.0;
    for (int c = 0; c < R; ++c) {
      const double d = Xi[(size_t)c * K + k] - obs_c[(size_t)j * R + c];
      s = (c == 0) ? __dmul_rn(d, d) : __dadd_rn(s, __dmul_rn(d, d));
    }
    m = fmin(m, sqrt(s) - rr);
  }
  m = warp_min(m);
  if (lane == 0) { d_mat[(size_t)i * M + j] = m; atomic_min_double(d_min, m); }
}

__global__ void fill_kernel(double* p, long long n, double v) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) p[g] = v;
}

}  // namespace scvx

using namespace scvx;

extern "C" int scvx_warm_start_batched(int model_id, int n_agents, int K, int M_max, const double* p0, const double* p1,
                                       const double* obs_c, const double* obs_r, const int* obs_count, double clearance,
                                       double* X0, double* U0, int* status, void* stream) {
  if (n_agents < 0 || K < 2 || M_max < 0) return bad_arg("n_agents/K/M_max");
  if (n_agents == 0) return SCVX_OK;
  if (!p0 || !p1 || !X0 || !U0 || !status) return bad_arg("null pointer");
  if (M_max > 0 && (!obs_c || !obs_r)) return bad_arg("obstacle tables");
  if (M_max > 1024) return bad_arg("M_max > 1024");
  cudaStream_t st = (cudaStream_t)stream;
  const int P_max = 2 * M_max + 2;
  switch (solver_shape_of(model_id)) {
    case SCVX_MODEL_UNICYCLE: {
      const size_t smem = (size_t)P_max * 2 * sizeof(double) + (size_t)(2 * P_max + 2) * sizeof(int);
      warm_start_kernel<2><<<n_agents, 128, smem, st>>>(K, M_max, p0, p1, obs_c, obs_r, obs_count, clearance, X0, U0, status);
      break;
    }
    case SCVX_MODEL_SINGLE_INTEGRATOR: {
      const size_t smem = (size_t)P_max * 3 * sizeof(double) + (size_t)(2 * P_max + 2) * sizeof(int);
      warm_start_kernel<3><<<n_agents, 128, smem, st>>>(K, M_max, p0, p1, obs_c, obs_r, obs_count, clearance, X0, U0, status);
      break;
    }
    default:
      return bad_arg("model_id");
  }
  SCVX_CHECK_LAUNCH("scvx_warm_start_batched");
  return SCVX_OK;
}

extern "C" int scvx_min_inter_agent_distance(int n_agents, int K, int n_x, int n_rows, const double* X, double* d_mat,
                                             double* d_min, void* stream) {
  if (n_agents < 0 || K < 1 || n_x < 1 || n_rows < 1 || n_rows > n_x) return bad_arg("n_agents/K/n_x/n_rows");
  if (!X || !d_mat || !d_min) return bad_arg("null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nn = (long long)n_agents * n_agents;
  if (nn > 0) fill_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(d_mat, nn, 0.0);
  fill_kernel<<<1, 32, 0, st>>>(d_min, 1, INFINITY);
  const long long pairs = (long long)n_agents * (n_agents - 1) / 2;
  if (pairs > 0) min_pair_distance_kernel<<<(unsigned)((pairs + 3) / 4), 128, 0, st>>>(n_agents, K, n_x, n_rows, X, d_mat, d_min);
  SCVX_CHECK_LAUNCH("scvx_min_inter_agent_distance");
  return SCVX_OK;
}

extern "C" int scvx_min_agent_obstacle_distance(int n_agents, int K, int n_x, int n_rows, int M, const double* X,
                                                const double* obs_c, const double* obs_r, double robot_radius, double* d_mat,
                                                double* d_min, void* stream) {
  if (n_agents < 0 || K < 1 || n_x < 1 || n_rows < 1 || n_rows > n_x || M < 0) return bad_arg("n_agents/K/n_x/n_rows/M");
  if (!X || !d_mat || !d_min || (M > 0 && (!obs_c || !obs_r))) return bad_arg("null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  fill_kernel<<<1, 32, 0, st>>>(d_min, 1, INFINITY);
  const long long items = (long long)n_agents * M;
  if (items > 0)
    min_obstacle_distance_kernel<<<(unsigned)((items + 3) / 4), 128, 0, st>>>(n_agents, K, n_x, n_rows, M, X, obs_c, obs_r,
                                                                           robot_radius, d_mat, d_min);
  SCVX_CHECK_LAUNCH("scvx_min_agent_obstacle_distance");
  return SCVX_OK;
}
