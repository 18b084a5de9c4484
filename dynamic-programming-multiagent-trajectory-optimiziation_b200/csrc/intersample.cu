// intersample.cu -- inter-sample obstacle clearance (SURVEY section 8 f rank 3), batched, fp64, sm_100a.
//
// Replaces, for the segment flow the reference builds with make_segment_f (SCvx/utils/intersample_collision.py:100-125:
// x(t) = state after integrating xdot = f(x, u(tau)) over [0, t * dt * sigma], u linear between the segment's end controls):
//   h_i                  intersample_collision.py:7-26    h = || T x(t) - p_c || - r,   T = first `m` rows of the identity
//   find_critical_times  :29-72   phi = central difference of h in t (eps), sign changes of phi on a grid of num_samples points
//                                 in [eps, dt - eps], bisection (<= 30 halvings, |b - a| < tol), keep 0 < t* < dt with phi2 > 0
//   linearize_h          :75-97   h0 and the central-difference gradient in x_k.  grad_u is identically ZERO in the reference
//                                 (the segment flow ignores its control argument), and is written as zeros here.
// One WARP per (agent, segment, obstacle): lanes share the grid samples; each bisection runs five halvings per round on the
// 32 lanes (the same decisions as the reference's sequential bisection), phi2 and the linearisation one flow per lane.
// The flow is RK4 in the normalised interval time (the FOH kernel's integrator, stopped at the fraction t).
#include "common.cuh"

namespace scvx {

constexpr int IS_MAX_SAMPLES = 128;

template <class M>
__device__ void flow_partial(const double* x0, const double* u0, const double* u1, double sigma, double dt, double frac, double* x) {
  constexpr int NX = M::NX, NU = M::NU;
  double du[NU], um = 1.0;
#pragma unroll
  for (int j = 0; j < NU; ++j) { du[j] = u1[j] - u0[j]; um = fmax(um, fmax(fabs(u0[j]), fabs(u1[j]))); }
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = x0[i];
  const double lam = fabs(sigma * dt * frac) * um;
  int ns = (int)ceil(fmax(100.0 * sqrt(lam), 130.0 * pow(lam, 1.25)));
  ns = ns < 4 ? 4 : (ns > 4096 ? 4096 : ns);
  const double h = frac * dt / (double)ns, inv_dt = 1.0 / dt;
  for (int s = 0; s < ns; ++s) {
    const double t = (double)s * h;
    double k1[NX], k2[NX], k3[NX], k4[NX], xs[NX], u[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = u0[j] + (t * inv_dt) * du[j];
    M::f_only(x, u, k1);
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = u0[j] + ((t + 0.5 * h) * inv_dt) * du[j];
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + 0.5 * h * sigma * k1[i];
    M::f_only(xs, u, k2);
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + 0.5 * h * sigma * k2[i];
    M::f_only(xs, u, k3);
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = u0[j] + ((t + h) * inv_dt) * du[j];
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + h * sigma * k3[i];
    M::f_only(xs, u, k4);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] += (h * sigma / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
  }
}

template <class M>
struct Segment {
  double x[M::NX], u0[M::NU], u1[M::NU], c[3], r, sigma, dt_foh;
  int m;
  // clearance at the fraction t of the segment, from the start state xs
  __device__ double h(const double* xs, double t) const {
    double xt[M::NX];
    flow_partial<M>(xs, u0, u1, sigma, dt_foh, t, xt);
    double s = 0.0;
    for (int i = 0; i < m; ++i) { const double d = xt[i] - c[i]; s += d * d; }
    return sqrt(s) - r;
  }
  __device__ double phi(double t, double eps) const { return (h(x, t + eps) - h(x, t - eps)) / (2.0 * eps); }
  __device__ double phi2(double t, double eps) const { return (phi(t + eps, eps) - phi(t - eps, eps)) / (2.0 * eps); }
};

template <class M>
__global__ void __launch_bounds__(128)
intersample_kernel(int n_agents, int K, int Mobs, int m, const double* __restrict__ X, const double* __restrict__ U,
                   const double* __restrict__ sigma, const double* __restrict__ obs_c, const double* __restrict__ obs_r,
                   double t_range, int num_samples, double eps, double tol, int max_roots, int* __restrict__ n_roots,
                   double* __restrict__ t_star, double* __restrict__ h0, double* __restrict__ grad_x) {
  constexpr int NX = M::NX, NU = M::NU;
  __shared__ double phis_all[4][IS_MAX_SAMPLES];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * 4 + wib;
  const long long items = (long long)n_agents * (K - 1) * Mobs;
  if (w >= items) return;
  const int j = (int)(w % Mobs);
  const long long ak = w / Mobs;
  const int k = (int)(ak % (K - 1)), agent = (int)(ak / (K - 1));
  Segment<M> sg;
#pragma unroll
  for (int i = 0; i < NX; ++i) sg.x[i] = X[((size_t)agent * NX + i) * K + k];
#pragma unroll
  for (int q = 0; q < NU; ++q) { sg.u0[q] = U[((size_t)agent * NU + q) * K + k]; sg.u1[q] = U[((size_t)agent * NU + q) * K + k + 1]; }
  for (int i = 0; i < 3; ++i) sg.c[i] = (i < m) ? obs_c[((size_t)agent * Mobs + j) * m + i] : 0.0;
  sg.r = obs_r[(size_t)agent * Mobs + j]; sg.sigma = sigma[agent]; sg.dt_foh = 1.0 / (double)(K - 1); sg.m = m;
  double* phis = phis_all[wib];
  // grid of phi values: ts = linspace(eps, t_range - eps, num_samples)
  const double lo = eps, hi = t_range - eps, step = (hi - lo) / (double)(num_samples - 1);
  auto ts = [&](int i) { return (i == num_samples - 1) ? hi : lo + (double)i * step; };
  for (int i = lane; i < num_samples; i += 32) phis[i] = sg.phi(ts(i), eps);
  __syncwarp();
  int found = 0;
  double* ts_out = t_star + (size_t)w * max_roots;
  double* h_out = h0 + (size_t)w * max_roots;
  double* g_out = grad_x + (size_t)w * max_roots * NX;
  for (int i = 0; i + 1 < num_samples; ++i) {           // warp-uniform: every lane walks the same candidates
    const double p0 = phis[i], p1 = phis[i + 1];
    if (!(p0 == 0.0 || p0 * p1 < 0.0)) continue;
    // The reference's bisection (<= 30 halvings, stop once |b - a| < tol), five halvings per round: the midpoints five
    // halvings can visit are the 31 interior points of a 32-way split of [a, b], so lane j evaluates phi there (lane 0: phi(a))
    // and the decisions "phi(a) phi(c) <= 0 ? b = c : a = c" are then replayed on the stored values.
    double a = ts(i), b = ts(i + 1);
    {
      int halvings = 0;
      bool done = false;
      while (!done && halvings < 30) {
        const double wd = (b - a) * (1.0 / 32.0);
        const double P = sg.phi(a + (double)lane * wd, eps);
        int lo = 0, hi = 32;
        for (int lvl = 0; lvl < 5 && halvings < 30; ++lvl) {
          const int c = (lo + hi) >> 1;
          const double pl = __shfl_sync(0xffffffffu, P, lo), pc = __shfl_sync(0xffffffffu, P, c);
          if (pl * pc <= 0.0) hi = c; else lo = c;
          ++halvings;
          if (fabs((double)(hi - lo) * wd) < tol) { done = true; break; }
        }
        const double a2 = a + (double)lo * wd, b2 = (hi == 32) ? b : a + (double)hi * wd;
        a = a2; b = b2;
      }
    }
    const double r = 0.5 * (a + b);
    if (!(0.0 < r && r < t_range)) continue;
    {
      // phi2(r) > 0 (intersample_collision.py:66-70): its four clearance evaluations on lanes 0..3
      const double tq = (lane & 2) ? r - eps : r + eps;
      const double hq = (lane < 4) ? sg.h(sg.x, (lane & 1) ? tq - eps : tq + eps) : 0.0;
      const double h0p = __shfl_sync(0xffffffffu, hq, 0), h1p = __shfl_sync(0xffffffffu, hq, 1);
      const double h2p = __shfl_sync(0xffffffffu, hq, 2), h3p = __shfl_sync(0xffffffffu, hq, 3);
      const double ph_p = (h0p - h1p) / (2.0 * eps), ph_m = (h2p - h3p) / (2.0 * eps);
      if (!((ph_p - ph_m) / (2.0 * eps) > 0.0)) continue;
    }
    if (found < max_roots) {
      // linearize_h at (x_k, t*): lanes 0..2*NX-1 take the perturbed states, lane 31 the nominal one
      double xs[NX];
#pragma unroll
      for (int q = 0; q < NX; ++q) xs[q] = sg.x[q];
      const int comp = lane >> 1;
      if (lane < 2 * NX) xs[comp] += (lane & 1) ? -eps : eps;
      const double hv = sg.h(xs, r);
      for (int q = 0; q < NX; ++q) {
        const double hp = __shfl_sync(0xffffffffu, hv, 2 * q), hm = __shfl_sync(0xffffffffu, hv, 2 * q + 1);
        if (lane == 0) g_out[found * NX + q] = (hp - hm) / (2.0 * eps);
      }
      const double hn = __shfl_sync(0xffffffffu, hv, 31);
      if (lane == 0) { ts_out[found] = r; h_out[found] = hn; }
    }
    ++found;
  }
  if (lane == 0) n_roots[w] = found;
}

}  // namespace scvx

using namespace scvx;

extern "C" int scvx_intersample_batched(int model_id, int n_agents, int K, int M, int proj_dim, const double* X, const double* U,
                                        const double* sigma, const double* obs_c, const double* obs_r, double t_range,
                                        int num_samples, double eps, double tol, int max_roots, int* n_roots, double* t_star,
                                        double* h0, double* grad_x, void* stream) {
  int nx, nu, d;
  if (!model_dims(model_id, &nx, &nu, &d)) return bad_arg("model_id");
  if (n_agents < 0 || K < 2 || M < 0 || proj_dim < 1 || proj_dim > nx) return bad_arg("n_agents/K/M/proj_dim");
  if (num_samples < 2 || num_samples > IS_MAX_SAMPLES || max_roots < 1 || !(eps > 0.0) || !(tol > 0.0)) return bad_arg("num_samples/max_roots/eps/tol");
  const long long items = (long long)n_agents * (K - 1) * M;
  if (items == 0) return SCVX_OK;
  if (!X || !U || !sigma || !obs_c || !obs_r || !n_roots || !t_star || !h0 || !grad_x) return bad_arg("null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((items + 3) / 4);
  if (model_id == SCVX_MODEL_UNICYCLE)
    intersample_kernel<Unicycle><<<blocks, 128, 0, st>>>(n_agents, K, M, proj_dim, X, U, sigma, obs_c, obs_r, t_range, num_samples, eps,
                                                         tol, max_roots, n_roots, t_star, h0, grad_x);
  else
    intersample_kernel<SingleIntegrator><<<blocks, 128, 0, st>>>(n_agents, K, M, proj_dim, X, U, sigma, obs_c, obs_r, t_range,
                                                                 num_samples, eps, tol, max_roots, n_roots, t_star, h0, grad_x);
  SCVX_CHECK_LAUNCH("scvx_intersample_batched");
  return SCVX_OK;
}

// Clearance samples along every segment (compute_intersample_clearance, SCvx/utils/analysis.py:64-104): h at the fractions
// t = i / resolution, i = 0..resolution-1, of each of the K-1 segments, for ONE obstacle per agent.
template <class M>
__global__ void __launch_bounds__(128)
clearance_samples_kernel(int n_agents, int K, int m, int resolution, const double* __restrict__ X, const double* __restrict__ U,
                         const double* __restrict__ sigma, const double* __restrict__ obs_c, const double* __restrict__ total_r,
                         double* __restrict__ h_cont) {
  constexpr int NX = M::NX, NU = M::NU;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_agents * (K - 1) * resolution;
  if (g >= total) return;
  const int i = (int)(g % resolution);
  const long long ak = g / resolution;
  const int k = (int)(ak % (K - 1)), agent = (int)(ak / (K - 1));
  Segment<M> sg;
#pragma unroll
  for (int q = 0; q < NX; ++q) sg.x[q] = X[((size_t)agent * NX + q) * K + k];
#pragma unroll
  for (int q = 0; q < NU; ++q) { sg.u0[q] = U[((size_t)agent * NU + q) * K + k]; sg.u1[q] = U[((size_t)agent * NU + q) * K + k + 1]; }
  for (int q = 0; q < 3; ++q) sg.c[q] = (q < m) ? obs_c[(size_t)agent * m + q] : 0.0;
  sg.r = total_r[agent]; sg.sigma = sigma[agent]; sg.dt_foh = 1.0 / (double)(K - 1); sg.m = m;
  h_cont[g] = sg.h(sg.x, (double)i / (double)resolution);
}

extern "C" int scvx_clearance_samples_batched(int model_id, int n_agents, int K, int proj_dim, int resolution, const double* X,
                                              const double* U, const double* sigma, const double* obs_c, const double* total_r,
                                              double* h_cont, void* stream) {
  int nx, nu, d;
  if (!model_dims(model_id, &nx, &nu, &d)) return bad_arg("model_id");
  if (n_agents < 0 || K < 2 || proj_dim < 1 || proj_dim > nx || resolution < 1) return bad_arg("n_agents/K/proj_dim/resolution");
  const long long total = (long long)n_agents * (K - 1) * resolution;
  if (total == 0) return SCVX_OK;
  if (!X || !U || !sigma || !obs_c || !total_r || !h_cont) return bad_arg("null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((total + 127) / 128);
  if (model_id == SCVX_MODEL_UNICYCLE)
    clearance_samples_kernel<Unicycle><<<blocks, 128, 0, st>>>(n_agents, K, proj_dim, resolution, X, U, sigma, obs_c, total_r, h_cont);
  else
    clearance_samples_kernel<SingleIntegrator><<<blocks, 128, 0, st>>>(n_agents, K, proj_dim, resolution, X, U, sigma, obs_c, total_r, h_cont);
  SCVX_CHECK_LAUNCH("scvx_clearance_samples_batched");
  return SCVX_OK;
}
