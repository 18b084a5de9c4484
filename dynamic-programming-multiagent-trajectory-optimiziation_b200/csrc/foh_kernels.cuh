// foh_kernels.cuh -- device code of stage 1 (first-order-hold discretisation and the nonlinear propagators), templated on the
// dynamics model.  DEVICE-ONLY and free of host headers on purpose: foh.cu instantiates it for the two shipped models at build
// time, and user_model.cu compiles THE SAME TEXT at run time with NVRTC for a model generated from a user's sympy expressions
// (BaseModel plug-in contract, SCvx/models/base_model.py:16-88; a model only has to provide NX, NU, D, eval and f_only).
//
// Replaces FirstOrderHold.calculate_discretization / _ode_dVdt / integrate_nonlinear_piecewise / integrate_nonlinear_full
// (SCvx/discretization/first_order_hold.py:52-162) for a batch of agents.
//
// Mapping: ONE THREAD per (agent, interval).  The augmented state V = [x, Phi, B-, B+, S, z] splits into a genuinely dynamic
// part (x, Phi: NX + NX^2 values) and pure quadratures (B-, B+, S, z) whose derivatives depend only on (t, x, Phi).  Classical
// RK4 on the full V is therefore identical to RK4 on (x, Phi) with the quadrature integrands accumulated at the four stage
// points with weights h/6, h/3, h/3, h/6 -- which keeps the per-thread live state at ~70 doubles (no spills) instead of
// 4 x 30.  Global loads/stores are coalesced over k (the fastest index of every array).
#pragma once

namespace scvx {

// Inverse of the NX x NX state-transition matrix: closed form for NX = 3 (the shipped models), Gauss-Jordan with partial
// pivoting (fully unrolled, register resident) for any other size a user model brings.
template <int N>
__device__ __forceinline__ void inv_small(const double (*P)[N], double (*Pi)[N]) {
  if constexpr (N == 3) {
    const double c00 = P[1][1] * P[2][2] - P[1][2] * P[2][1];
    const double c01 = P[1][2] * P[2][0] - P[1][0] * P[2][2];
    const double c02 = P[1][0] * P[2][1] - P[1][1] * P[2][0];
    const double det = P[0][0] * c00 + P[0][1] * c01 + P[0][2] * c02;
    const double r = 1.0 / det;
    Pi[0][0] = c00 * r;
    Pi[1][0] = c01 * r;
    Pi[2][0] = c02 * r;
    Pi[0][1] = (P[0][2] * P[2][1] - P[0][1] * P[2][2]) * r;
    Pi[1][1] = (P[0][0] * P[2][2] - P[0][2] * P[2][0]) * r;
    Pi[2][1] = (P[0][1] * P[2][0] - P[0][0] * P[2][1]) * r;
    Pi[0][2] = (P[0][1] * P[1][2] - P[0][2] * P[1][1]) * r;
    Pi[1][2] = (P[0][2] * P[1][0] - P[0][0] * P[1][2]) * r;
    Pi[2][2] = (P[0][0] * P[1][1] - P[0][1] * P[1][0]) * r;
  } else {
    double Aug[N][2 * N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) { Aug[i][j] = P[i][j]; Aug[i][N + j] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
    for (int c = 0; c < N; ++c) {
      // partial pivoting by compare-and-swap of whole rows (static indices only: everything stays in registers)
#pragma unroll
      for (int r = c + 1; r < N; ++r) {
        const bool sw = fabs(Aug[r][c]) > fabs(Aug[c][c]);
#pragma unroll
        for (int j = 0; j < 2 * N; ++j) {
          const double t = Aug[c][j];
          Aug[c][j] = sw ? Aug[r][j] : t;
          Aug[r][j] = sw ? t : Aug[r][j];
        }
      }
      const double piv = 1.0 / Aug[c][c];
#pragma unroll
      for (int j = 0; j < 2 * N; ++j) Aug[c][j] *= piv;
#pragma unroll
      for (int r = 0; r < N; ++r) {
        if (r == c) continue;
        const double f = Aug[r][c];
#pragma unroll
        for (int j = 0; j < 2 * N; ++j) Aug[r][j] -= f * Aug[c][j];
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) Pi[i][j] = Aug[i][N + j];
  }
}

// One evaluation of _ode_dVdt (first_order_hold.py:89-125) at time t with state (x, Phi):
// returns kx = sigma f, kP = A_s Phi and ADDS w * {alpha Phi^-1 B_s, beta Phi^-1 B_s, Phi^-1 f,
// Phi^-1(-A_s x - B_s u)} into the quadrature accumulators.
template <class M>
__device__ __forceinline__ void rhs_stage(double t, double inv_dt, double sigma, const double* u0, const double* du,
                                          const double* x, const double (*Phi)[M::NX], double* kx,
                                          double (*kP)[M::NX], double w, double (*Bm)[M::NU], double (*Bp)[M::NU],
                                          double* S, double* Z) {
  constexpr int NX = M::NX, NU = M::NU;
  const double beta = t * inv_dt;        // t / dt
  const double alpha = 1.0 - beta;       // (dt - t) / dt
  double u[NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) u[j] = u0[j] + beta * du[j];
  double f[NX], A[NX][NX], B[NX][NU];
  M::eval(x, u, f, A, B);
  // A_s = sigma A, B_s = sigma B
  double zt[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    kx[i] = sigma * f[i];
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      A[i][j] *= sigma;
      acc -= A[i][j] * x[j];
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      B[i][j] *= sigma;
      acc -= B[i][j] * u[j];
    }
    zt[i] = acc;   // -A_s x - B_s u
  }
#pragma unroll
  for (int i = 0; i < NX; ++i)
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < NX; ++l) acc += A[i][l] * Phi[l][j];
      kP[i][j] = acc;
    }
  if (w == 0.0) return;      // a stage whose quadrature weight is zero (stage 2 of the 6th-order scheme) only feeds (x, Phi)
  double Pi[NX][NX];
  inv_small<NX>(Phi, Pi);
  const double wa = w * alpha, wb = w * beta;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double sf = 0.0, sz = 0.0;
#pragma unroll
    for (int l = 0; l < NX; ++l) {
      sf += Pi[i][l] * f[l];
      sz += Pi[i][l] * zt[l];
    }
    S[i] += w * sf;
    Z[i] += w * sz;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double pb = 0.0;
#pragma unroll
      for (int l = 0; l < NX; ++l) pb += Pi[i][l] * B[l][j];
      Bm[i][j] += wa * pb;
      Bp[i][j] += wb * pb;
    }
  }
}

__device__ __forceinline__ int auto_substeps(double sigma, double dt, const double* u0, const double* u1, int nu) {
  // Measured on the unicycle (tests/golden, K=50, |u|<=1): the error of B_bar/C_bar relative to their own scale
  // behaves like 8.6e-3 * lambda^2 / n^4 for small lambda = sigma*dt*max(1,|u|) (first-order-hold ramp of u inside the
  // interval) and like lambda^5/(28 n^4) for large lambda.  n below puts both at <= ~1.5e-10.
  double um = 1.0;
  for (int j = 0; j < nu; ++j) um = fmax(um, fmax(fabs(u0[j]), fabs(u1[j])));
  const double lam = fabs(sigma) * dt * um;
  double n = ceil(fmax(100.0 * sqrt(lam), 130.0 * pow(lam, 1.25)));
  n = fmin(fmax(n, 8.0), 8192.0);
  return (int)n;
}

// Sub-steps of the 6th-order scheme for the same accuracy target (measured against the reference right-hand side integrated at
// rtol 1e-13 over lambda = 0.02 .. 4, 24 random (x, u0, u1) each: the smallest n with error <= 1e-10 of scale was
// 2, 3, 4, 5, 8, 10, 16, 20 at lambda = 0.02, 0.05, 0.12, 0.25, 0.5, 1, 2, 4  ~ 10 sqrt(lambda); the law adds ~25 % in n,
// i.e. a factor ~4 in error, on top).
__device__ __forceinline__ int auto_substeps_rk6(double sigma, double dt, const double* u0, const double* u1, int nu) {
  double um = 1.0;
  for (int j = 0; j < nu; ++j) um = fmax(um, fmax(fabs(u0[j]), fabs(u1[j])));
  const double lam = fabs(sigma) * dt * um;
  double n = ceil(11.5 * sqrt(lam)) + 1.0;
  n = fmin(fmax(n, 2.0), 4096.0);
  return (int)n;
}

template <class M>
__device__ __forceinline__ void foh_rk4_body(int n_agents, int K, int n_sub, const double* __restrict__ X, const double* __restrict__ U,
                                             const double* __restrict__ sigma_arr, double* __restrict__ A_bar, double* __restrict__ B_bar,
                                             double* __restrict__ C_bar, double* __restrict__ S_bar, double* __restrict__ z_bar) {
  constexpr int NX = M::NX, NU = M::NU;
  const int Km1 = K - 1;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n_agents * Km1) return;
  const int agent = (int)(gid / Km1);
  const int k = (int)(gid - (long long)agent * Km1);
  const double sigma = sigma_arr[agent];
  const double dt = 1.0 / (double)Km1;
  const double inv_dt = (double)Km1;

  const double* Xa = X + (size_t)agent * NX * K;
  const double* Ua = U + (size_t)agent * NU * K;
  double x[NX], u0[NU], du[NU], u1[NU];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = Xa[(size_t)i * K + k];
#pragma unroll
  for (int j = 0; j < NU; ++j) {
    u0[j] = Ua[(size_t)j * K + k];
    u1[j] = Ua[(size_t)j * K + k + 1];
    du[j] = u1[j] - u0[j];
  }
  double Phi[NX][NX], Bm[NX][NU], Bp[NX][NU], S[NX], Z[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    S[i] = 0.0; Z[i] = 0.0;
#pragma unroll
    for (int j = 0; j < NX; ++j) Phi[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < NU; ++j) { Bm[i][j] = 0.0; Bp[i][j] = 0.0; }
  }
  if (n_sub <= 0) {
    // Sub-step count chosen on the device: Butcher's 6th-order, 7-stage explicit Runge-Kutta scheme (1964).  Same splitting of
    // the augmented state as below (the quadratures ride on the stage evaluations of (x, Phi) with the scheme's weights), and
    // 4-5x fewer right-hand-side evaluations than RK4 for the same 1e-10 (50 x 4 vs 7 x 7 at lambda = 0.25).  A caller that
    // asks for n_sub steps explicitly gets classical RK4, the scheme the interface names.
    constexpr double c6[7] = {0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0 / 3.0, 0.5, 0.5, 1.0};
    constexpr double a6[7][6] = {{0, 0, 0, 0, 0, 0},
                                 {1.0 / 3.0, 0, 0, 0, 0, 0},
                                 {0, 2.0 / 3.0, 0, 0, 0, 0},
                                 {1.0 / 12.0, 1.0 / 3.0, -1.0 / 12.0, 0, 0, 0},
                                 {-1.0 / 16.0, 9.0 / 8.0, -3.0 / 16.0, -3.0 / 8.0, 0, 0},
                                 {0, 9.0 / 8.0, -3.0 / 8.0, -3.0 / 4.0, 0.5, 0},
                                 {9.0 / 44.0, -9.0 / 11.0, 63.0 / 44.0, 18.0 / 11.0, -16.0 / 11.0, 0}};
    constexpr double b6[7] = {11.0 / 120.0, 0.0, 27.0 / 40.0, 27.0 / 40.0, -4.0 / 15.0, -4.0 / 15.0, 11.0 / 120.0};
    const int ns6 = auto_substeps_rk6(sigma, dt, u0, u1, NU);
    const double h = dt / (double)ns6;
    for (int s = 0; s < ns6; ++s) {
      const double t = (double)s * h;
      double kx[7][NX], kP[7][NX][NX];
#pragma unroll
      for (int st = 0; st < 7; ++st) {
        double xs[NX], Ps[NX][NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double ax = 0.0;
#pragma unroll
          for (int j = 0; j < st; ++j)
            if (a6[st][j] != 0.0) ax += a6[st][j] * kx[j][i];
          xs[i] = x[i] + h * ax;
#pragma unroll
          for (int l = 0; l < NX; ++l) {
            double ap = 0.0;
#pragma unroll
            for (int j = 0; j < st; ++j)
              if (a6[st][j] != 0.0) ap += a6[st][j] * kP[j][i][l];
            Ps[i][l] = Phi[i][l] + h * ap;
          }
        }
        rhs_stage<M>(t + c6[st] * h, inv_dt, sigma, u0, du, xs, Ps, kx[st], kP[st], b6[st] * h, Bm, Bp, S, Z);
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double ax = 0.0;
#pragma unroll
        for (int st = 0; st < 7; ++st)
          if (b6[st] != 0.0) ax += b6[st] * kx[st][i];
        x[i] += h * ax;
#pragma unroll
        for (int l = 0; l < NX; ++l) {
          double ap = 0.0;
#pragma unroll
          for (int st = 0; st < 7; ++st)
            if (b6[st] != 0.0) ap += b6[st] * kP[st][i][l];
          Phi[i][l] += h * ap;
        }
      }
    }
  }
  const int ns = (n_sub > 0) ? n_sub : 0;
  const double h = (n_sub > 0) ? dt / (double)ns : 0.0;
  const double h2 = 0.5 * h, h6 = h / 6.0, h3 = h / 3.0;

  for (int s = 0; s < ns; ++s) {
    const double t = (double)s * h;
    double kx[NX], kP[NX][NX], ax[NX], aP[NX][NX], xs[NX], Ps[NX][NX];
    // stage 1
    rhs_stage<M>(t, inv_dt, sigma, u0, du, x, Phi, kx, kP, h6, Bm, Bp, S, Z);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      ax[i] = kx[i]; xs[i] = x[i] + h2 * kx[i];
#pragma unroll
      for (int j = 0; j < NX; ++j) { aP[i][j] = kP[i][j]; Ps[i][j] = Phi[i][j] + h2 * kP[i][j]; }
    }
    // stage 2
    rhs_stage<M>(t + h2, inv_dt, sigma, u0, du, xs, Ps, kx, kP, h3, Bm, Bp, S, Z);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      ax[i] += 2.0 * kx[i]; xs[i] = x[i] + h2 * kx[i];
#pragma unroll
      for (int j = 0; j < NX; ++j) { aP[i][j] += 2.0 * kP[i][j]; Ps[i][j] = Phi[i][j] + h2 * kP[i][j]; }
    }
    // stage 3
    rhs_stage<M>(t + h2, inv_dt, sigma, u0, du, xs, Ps, kx, kP, h3, Bm, Bp, S, Z);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      ax[i] += 2.0 * kx[i]; xs[i] = x[i] + h * kx[i];
#pragma unroll
      for (int j = 0; j < NX; ++j) { aP[i][j] += 2.0 * kP[i][j]; Ps[i][j] = Phi[i][j] + h * kP[i][j]; }
    }
    // stage 4
    rhs_stage<M>(t + h, inv_dt, sigma, u0, du, xs, Ps, kx, kP, h6, Bm, Bp, S, Z);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      x[i] += h6 * (ax[i] + kx[i]);
#pragma unroll
      for (int j = 0; j < NX; ++j) Phi[i][j] += h6 * (aP[i][j] + kP[i][j]);
    }
  }

  // epilogue (first_order_hold.py:76-85): A_bar = Phi, B_bar = Phi B-, C_bar = Phi B+, S_bar = Phi S, z_bar = Phi z,
  // matrices flattened column-major into column k.
  double* Ao = A_bar + (size_t)agent * NX * NX * Km1 + k;
  double* Bo = B_bar + (size_t)agent * NX * NU * Km1 + k;
  double* Co = C_bar + (size_t)agent * NX * NU * Km1 + k;
  double* So = S_bar + (size_t)agent * NX * Km1 + k;
  double* Zo = z_bar + (size_t)agent * NX * Km1 + k;
#pragma unroll
  for (int j = 0; j < NX; ++j)
#pragma unroll
    for (int i = 0; i < NX; ++i) Ao[(size_t)(j * NX + i) * Km1] = Phi[i][j];
#pragma unroll
  for (int j = 0; j < NU; ++j)
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double b = 0.0, c = 0.0;
#pragma unroll
      for (int l = 0; l < NX; ++l) { b += Phi[i][l] * Bm[l][j]; c += Phi[i][l] * Bp[l][j]; }
      Bo[(size_t)(j * NX + i) * Km1] = b;
      Co[(size_t)(j * NX + i) * Km1] = c;
    }
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double s = 0.0, z = 0.0;
#pragma unroll
    for (int l = 0; l < NX; ++l) { s += Phi[i][l] * S[l]; z += Phi[i][l] * Z[l]; }
    So[(size_t)i * Km1] = s;
    Zo[(size_t)i * Km1] = z;
  }
}

// xdot = f(x, u(t)) over [0, dt*sigma]  (first_order_hold.py:157-162), integrated in tau = t/sigma.
template <class M>
__device__ __forceinline__ void flow_interval(double* x, const double* u0, const double* u1, double sigma, double dt,
                                              int n_sub) {
  constexpr int NX = M::NX, NU = M::NU;
  double du[NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) du[j] = u1[j] - u0[j];
  const int ns = (n_sub > 0) ? n_sub : auto_substeps(sigma, dt, u0, u1, NU);
  const double h = dt / (double)ns, inv_dt = 1.0 / dt;
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
__device__ __forceinline__ void integrate_piecewise_body(int n_agents, int K, int n_sub, const double* __restrict__ X_lin,
                                                         const double* __restrict__ U, const double* __restrict__ sigma_arr,
                                                         double* __restrict__ X_nl) {
  constexpr int NX = M::NX, NU = M::NU;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n_agents * K) return;
  const int agent = (int)(gid / K);
  const int k = (int)(gid - (long long)agent * K);     // output column k
  const double* Xa = X_lin + (size_t)agent * NX * K;
  const double* Ua = U + (size_t)agent * NU * K;
  double* Xo = X_nl + (size_t)agent * NX * K;
  if (k == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) Xo[(size_t)i * K] = Xa[(size_t)i * K];
    return;
  }
  double x[NX], u0[NU], u1[NU];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = Xa[(size_t)i * K + k - 1];
#pragma unroll
  for (int j = 0; j < NU; ++j) { u0[j] = Ua[(size_t)j * K + k - 1]; u1[j] = Ua[(size_t)j * K + k]; }
  flow_interval<M>(x, u0, u1, sigma_arr[agent], 1.0 / (double)(K - 1), n_sub);
#pragma unroll
  for (int i = 0; i < NX; ++i) Xo[(size_t)i * K + k] = x[i];
}

template <class M>
__device__ __forceinline__ void integrate_full_body(int n_agents, int K, int n_sub, const double* __restrict__ x0,
                                                    const double* __restrict__ U, const double* __restrict__ sigma_arr,
                                                    double* __restrict__ X_nl) {
  constexpr int NX = M::NX, NU = M::NU;
  const int agent = blockIdx.x * blockDim.x + threadIdx.x;
  if (agent >= n_agents) return;
  const double* Ua = U + (size_t)agent * NU * K;
  double* Xo = X_nl + (size_t)agent * NX * K;
  double x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) { x[i] = x0[(size_t)agent * NX + i]; Xo[(size_t)i * K] = x[i]; }
  const double sigma = sigma_arr[agent], dt = 1.0 / (double)(K - 1);
  for (int k = 0; k < K - 1; ++k) {
    double u0[NU], u1[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) { u0[j] = Ua[(size_t)j * K + k]; u1[j] = Ua[(size_t)j * K + k + 1]; }
    flow_interval<M>(x, u0, u1, sigma, dt, n_sub);
#pragma unroll
    for (int i = 0; i < NX; ++i) Xo[(size_t)i * K + k + 1] = x[i];
  }
}

}  // namespace scvx
