// lti_qp.cu -- the per-robot QP of the Distributed_opt scripts, batched, fp64, sm_100a.
//
// Replaces the cvxpy -> CLARABEL solves of
//   Distributed_opt/ADMM_decentralized.py:52-98   (per-robot perturbation QP of the 2-D double integrator, n=4, m=2)
//   Distributed_opt/dist_scvx_3d.py:51-111        (3-D twin, n=6, m=3, collision rows with one slack per time step inside)
// and, with `scvx_sbar_qp_batched`, the per-time-step consensus QP of ADMM_decentralized.py:106-139.
//
//   min  c_w sum_{t<T-1} |u_t + w_t|^2 + sum_t [ lin_t . dpos_t + rho/2 |dpos_t - sbar_t|^2 ] + c_S sum_t S_t
//   s.t. d_0 = 0, d_{T-1} = x_des - x_{T-1}, x_{t+1} + d_{t+1} = A (x_t + d_t) + B (u_t + w_t)       (hard equalities)
//        |w_t|_1 <= tr, box on (x_t + d_t)[0:2],  h_tq - g_tq . d_t[0:dc] <= S_t, S_t >= 0            (t = 0..T-2)
//
// Algorithm (oracle/lti_ipm.py is the numpy twin): primal-dual interior point, Mehrotra predictor-corrector, separate
// primal/dual step lengths.  The Newton step is an equality-constrained LQ problem: a Riccati recursion handles the
// dynamics exactly; the terminal equality is enforced through n extra right-hand sides (one per terminal multiplier) and an
// n x n solve.  The shared slack S_t of a time step is eliminated analytically (rank-one correction of the stage Hessian).
// One thread block per robot; row passes are parallel over t; the Riccati sweeps are sequential in t but cooperative within
// warp 0: every small matrix product (4x4 / 6x6 blocks, N+1 right-hand sides) is one entry per lane between __syncwarp's, and
// the corrector solve reuses the predictor's matrix recursion and terminal-multiplier columns (only column 0 is redone).
#include "common.cuh"
#include "reduce.cuh"

namespace scvx {

constexpr double LTI_TINY = 1e-14;

template <int N, int M>
struct LtiDims {
  static constexpr int DC = N / 2;            // position dimension
  static constexpr int NE = 1 << M;           // trust-region sign rows per step
  static constexpr int NR = N + 1;            // right-hand sides of the vector Riccati pass
  // shared-memory doubles per time step
  static constexpr int PER_T = 2 * N + 2 * M      // d, pi | w, (pad)
                               + 2 * N + 2 * M    // dd, dd_aff | dw, dw_aff
                               + N                // pi_new
                               + DC * DC + M * M + N + M   // Q (position block), R, qv, rv
                               + N                // ctil (equality residual)
                               + N * N + M * N + M * N + M * M   // P, K, Qux, Quu^-1
                               + M * NR + N * NR                // k, p (all right-hand sides)
                               + N + M                          // best primal-feasible iterate seen
                               + N * NR;                        // state response to every right-hand-side column (xsT)
  // scratch of the warp-cooperative Riccati sweeps: PA, PB, Quu, pc, qu, qx, xs (two buffers), us, Mt, nu
  static constexpr int SCR = N * N + N * M + M * M + N * NR + M * NR + N * NR + 2 * N * NR + M * NR + N * N + N;
};

// row-state layout in the global workspace, per robot: lamT[NE][T], sT[NE][T], lamB[4][T], sB[4][T],
// lamC[nq][T], sC[nq][T], lam0[T], s0[T], xi[T]
template <int N, int M>
__host__ __device__ inline size_t lti_ws_doubles(int T, int nq) {
  return (size_t)T * (2 * LtiDims<N, M>::NE + 8 + 2 * (size_t)nq + 3);
}

template <int N, int M>
__global__ void __launch_bounds__(128, 1) lti_qp_kernel(scvx_lti_args a) {
  using Dm = LtiDims<N, M>;
  constexpr int DC = Dm::DC, NE = Dm::NE, NR = Dm::NR;
  const int T = a.T, robot = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, nq = a.nq;
  extern __shared__ __align__(16) double sm[];
  double* d = sm;                         // [T][N]
  double* pi = d + (size_t)T * N;         // [T][N]
  double* w = pi + (size_t)T * N;         // [T][M]
  double* dd = w + (size_t)T * M;         // [T][N]
  double* dda = dd + (size_t)T * N;       // [T][N]  affine direction
  double* dw = dda + (size_t)T * N;       // [T][M]
  double* dwa = dw + (size_t)T * M;       // [T][M]
  double* pin = dwa + (size_t)T * M;      // [T][N]  costates of the last LQ solve
  double* Qp = pin + (size_t)T * N;       // [T][DC*DC]
  double* Rm = Qp + (size_t)T * DC * DC;  // [T][M*M]
  double* qv = Rm + (size_t)T * M * M;    // [T][N]
  double* rv = qv + (size_t)T * N;        // [T][M]
  double* ct = rv + (size_t)T * M;        // [T][N]
  double* Pm = ct + (size_t)T * N;        // [T][N*N]
  double* Km = Pm + (size_t)T * N * N;    // [T][M*N]
  double* Qux = Km + (size_t)T * M * N;   // [T][M*N]
  double* Qui = Qux + (size_t)T * M * N;  // [T][M*M]
  double* kk = Qui + (size_t)T * M * M;   // [T][M*NR]
  double* pp = kk + (size_t)T * M * NR;   // [T][N*NR]
  double* dbest = pp + (size_t)T * N * NR;   // [T][N]
  double* wbest = dbest + (size_t)T * N;     // [T][M]
  double* xsT = wbest + (size_t)T * M;    // [T][N*NR]
  double* gl = xsT + (size_t)T * N * NR;  // [32]
  double* red = gl + 32;                  // [9][12]
  double* AB = red + 9 * 12;              // A [N*N], B [N*M]
  double* Am = AB;
  double* Bm = AB + N * N;
  double* sPA = Bm + N * M;               // Riccati scratch (warp 0)
  double* sPB = sPA + N * N;
  double* sQuu = sPB + N * M;
  double* sPc = sQuu + M * M;
  double* sQu = sPc + N * NR;
  double* sQv = sQu + M * NR;
  double* sXs = sQv + N * NR;
  double* sUs = sXs + 2 * N * NR;
  double* sMt = sUs + M * NR;
  double* sNu = sMt + N * N;

  const double* x = a.x + (size_t)robot * T * N;
  const double* u = a.u + (size_t)robot * T * M;
  const double* lin = a.lin ? a.lin + (size_t)robot * T * 2 : nullptr;
  const double* sbar = a.sbar ? a.sbar + (size_t)robot * T * 2 : nullptr;
  const double* colh = nq ? a.col_h + (size_t)robot * nq * (T - 1) : nullptr;
  const double* colg = nq ? a.col_g + (size_t)robot * nq * DC * (T - 1) : nullptr;
  double* base = (double*)a.workspace + (size_t)robot * lti_ws_doubles<N, M>(T, nq);
  double* lamT = base;              double* sT = lamT + (size_t)NE * T;
  double* lamB = sT + (size_t)NE * T; double* sB = lamB + 4 * (size_t)T;
  double* lamC = sB + 4 * (size_t)T;  double* sC = lamC + (size_t)nq * T;
  double* lam0 = sC + (size_t)nq * T; double* s0 = lam0 + T; double* xiv = s0 + T;
  const double lo[2] = {a.box_lo0, a.box_lo1}, hi[2] = {a.box_hi0, a.box_hi1};
  const double cS = a.c_S, cw = a.c_w, rho = a.rho, trr = a.tr;
  const double mu0 = fmax(1.0, cS);
  auto H = [&](int q, int t) -> double { return colh[(size_t)q * (T - 1) + t]; };
  auto G = [&](int q, int c, int t) -> double { return colg[((size_t)q * DC + c) * (T - 1) + t]; };
  auto LIN = [&](int t, int i) -> double { return lin ? lin[t * 2 + i] : 0.0; };
  auto SBAR = [&](int t, int i) -> double { return sbar ? sbar[t * 2 + i] : 0.0; };

  for (int i = tid; i < N * N + N * M; i += nthr) AB[i] = (i < N * N) ? a.Ad[i] : a.Bd[i - N * N];
  if (tid < 32) gl[tid] = 0.0;
  // ---- start: d = 0 except the fixed terminal value, w = 0 ----------------------------------------------------
  for (int t = tid; t < T; t += nthr) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      d[t * N + i] = (t == T - 1) ? a.x_des[(size_t)robot * N + i] - x[(T - 1) * N + i] : 0.0;
      pi[t * N + i] = 0.0; dd[t * N + i] = 0.0; dda[t * N + i] = 0.0; pin[t * N + i] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < M; ++j) { w[t * M + j] = 0.0; dw[t * M + j] = 0.0; dwa[t * M + j] = 0.0; }
  }
  __syncthreads();
  for (int t = tid; t < T - 1; t += nthr) {
#pragma unroll
    for (int e = 0; e < NE; ++e) { sT[(size_t)e * T + t] = trr; lamT[(size_t)e * T + t] = mu0 / trr; }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double p = x[t * N + i];
      double s = fmax(hi[i] - p, 1e-2); sB[(size_t)i * T + t] = s; lamB[(size_t)i * T + t] = mu0 / s;
      s = fmax(p - lo[i], 1e-2); sB[(size_t)(2 + i) * T + t] = s; lamB[(size_t)(2 + i) * T + t] = mu0 / s;
    }
    if (nq) {
      // dual-feasible central start of the stage slack: sum_q mu0/(xi - viol_q) + mu0/xi = c_S (bisection)
      double vmax = -1e300;
      for (int q = 0; q < nq; ++q) vmax = fmax(vmax, H(q, t));
      const double L = fmax(vmax, 0.0);
      double lo_ = L, hi_ = L + (nq + 1) * mu0 / cS;
      for (int itb = 0; itb < 60; ++itb) {
        const double mid = 0.5 * (lo_ + hi_);
        double f = mu0 / mid - cS;
        for (int q = 0; q < nq; ++q) f += mu0 / (mid - H(q, t));
        if (f > 0.0) lo_ = mid; else hi_ = mid;
      }
      const double xi = hi_;
      xiv[t] = xi; s0[t] = xi; lam0[t] = mu0 / xi;
      for (int q = 0; q < nq; ++q) { const double s = xi - H(q, t); sC[(size_t)q * T + t] = s; lamC[(size_t)q * T + t] = mu0 / s; }
    }
  }
  __syncthreads();

  const int n_act = T - 2;                                  // stages 1..T-2 carry box / collision rows
  const double n_rows = (double)(T - 1) * NE + (double)n_act * 4 + (nq ? (double)n_act * (nq + 1) : 0.0);
  int status = SCVX_ST_MAXITER, it = 0;
  const int max_iter = a.max_iter > 0 ? a.max_iter : 60;
  double obj_prev = 0.0;

  // ---- Riccati LQ solve on warp 0, lanes spread over the entries of each small matrix product ------------------------
  // with_matrix = true  (predictor): matrix recursion + all NR right-hand sides (column 0 = the gradient, columns 1..N = the
  //                      terminal multipliers); the terminal response matrix Mt is kept in shared memory.
  // with_matrix = false (corrector): P, K, Qux, Quu^-1 and columns 1..N of p / k / Mt are unchanged -- only column 0 is redone.
  // Each time step is a handful of __syncwarp-separated stages; every stage is one entry (<= N multiply-adds) per lane.
  auto riccati = [&](bool with_matrix, double* out_d, double* out_w) {
    const int lane = tid;                               // called by tid < 32 only
    const int nc = with_matrix ? NR : 1;                // right-hand-side columns handled in this call
    if (with_matrix) {
      for (int e = lane; e < N * NR; e += 32) { const int i = e / NR, c = e - i * NR; pp[(size_t)(T - 1) * N * NR + e] = (c >= 1 && c - 1 == i) ? 1.0 : 0.0; }
      for (int e = lane; e < N * N; e += 32) Pm[(size_t)(T - 1) * N * N + e] = 0.0;
    }
    __syncwarp();
    // ---- backward
    for (int t = T - 2; t >= 0; --t) {
      const double* Pn = Pm + (size_t)(t + 1) * N * N;
      double* Pt = Pm + (size_t)t * N * N;
      double* Kt = Km + (size_t)t * M * N;
      double* Qx = Qux + (size_t)t * M * N;
      double* Qi = Qui + (size_t)t * M * M;
      const double* pn = pp + (size_t)(t + 1) * N * NR;
      double* pt = pp + (size_t)t * N * NR;
      double* kt = kk + (size_t)t * M * NR;
      // stage a: PA = P A, PB = P B | pc = p_next (+ P c~ in column 0)
      if (with_matrix) {
        for (int e = lane; e < N * (N + M); e += 32) {
          if (e < N * N) {
            const int i = e / N, j = e - i * N;
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < N; ++l) s += Pn[i * N + l] * Am[l * N + j];
            sPA[e] = s;
          } else {
            const int e2 = e - N * N, i = e2 / M, j = e2 - i * M;
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < N; ++l) s += Pn[i * N + l] * Bm[l * M + j];
            sPB[e2] = s;
          }
        }
      }
      for (int e = lane; e < N * nc; e += 32) {
        const int i = e / nc, c = e - i * nc;
        double v = pn[i * NR + c];
        if (c == 0) {
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < N; ++l) s += Pn[i * N + l] * ct[(size_t)t * N + l];
          v += s;
        }
        sPc[i * NR + c] = v;
      }
      __syncwarp();
      // stage b: Quu = R + B'PB, Qux = B'PA | qu = B'pc (+ r), qx = A'pc (+ q)
      if (with_matrix) {
        for (int e = lane; e < M * (M + N); e += 32) {
          if (e < M * M) {
            const int i = e / M, j = e - i * M;
            double s = Rm[(size_t)t * M * M + e];
#pragma unroll
            for (int l = 0; l < N; ++l) s += Bm[l * M + i] * sPB[l * M + j];
            sQuu[e] = s;
          } else {
            const int e2 = e - M * M, i = e2 / N, j = e2 - i * N;
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < N; ++l) s += Bm[l * M + i] * sPA[l * N + j];
            Qx[e2] = s;
          }
        }
      }
      for (int e = lane; e < (M + N) * nc; e += 32) {
        const int r = e / nc, c = e - r * nc;
        if (r < M) {
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < N; ++l) s += Bm[l * M + r] * sPc[l * NR + c];
          sQu[r * NR + c] = s + ((c == 0) ? rv[(size_t)t * M + r] : 0.0);
        } else {
          const int i = r - M;
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < N; ++l) s += Am[l * N + i] * sPc[l * NR + c];
          sQv[i * NR + c] = s + ((c == 0) ? qv[(size_t)t * N + i] : 0.0);
        }
      }
      __syncwarp();
      if (with_matrix) {
        // stage c: inverse of the SPD M x M block by Gauss-Jordan (M = 2 or 3), every lane redundantly in registers
        double Ai[M][2 * M];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j < M; ++j) { Ai[i][j] = sQuu[i * M + j]; Ai[i][M + j] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
        for (int c = 0; c < M; ++c) {
          const double r = 1.0 / Ai[c][c];
#pragma unroll
          for (int j = 0; j < 2 * M; ++j) Ai[c][j] *= r;
#pragma unroll
          for (int i = 0; i < M; ++i)
            if (i != c) {
              const double f = Ai[i][c];
#pragma unroll
              for (int j = 0; j < 2 * M; ++j) Ai[i][j] -= f * Ai[c][j];
            }
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) Qi[i * M + j] = Ai[i][M + j];
        }
        __syncwarp();
      }
      // stage d: K = -Quu^-1 Qux | k = -Quu^-1 qu
      if (with_matrix) {
        for (int e = lane; e < M * N; e += 32) {
          const int i = e / N, j = e - i * N;
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < M; ++l) s -= Qi[i * M + l] * Qx[l * N + j];
          Kt[e] = s;
        }
      }
      for (int e = lane; e < M * nc; e += 32) {
        const int i = e / nc, c = e - i * nc;
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < M; ++l) s -= Qi[i * M + l] * sQu[l * NR + c];
        kt[i * NR + c] = s;
      }
      __syncwarp();
      // stage e: P_t = Q_t + A'PA + Qux'K (symmetrised) | p = qx + Qux'k
      if (with_matrix) {
        for (int e = lane; e < N * N; e += 32) {
          const int i = e / N, j = e - i * N;
          double s = 0.0, s2 = 0.0;
#pragma unroll
          for (int l = 0; l < N; ++l) { s += Am[l * N + i] * sPA[l * N + j]; s2 += Am[l * N + j] * sPA[l * N + i]; }
#pragma unroll
          for (int l = 0; l < M; ++l) { s += Qx[l * N + i] * Kt[l * N + j]; s2 += Qx[l * N + j] * Kt[l * N + i]; }
          if (i < DC && j < DC) { s += Qp[(size_t)t * DC * DC + i * DC + j]; s2 += Qp[(size_t)t * DC * DC + j * DC + i]; }
          Pt[e] = 0.5 * (s + s2);
        }
      }
      for (int e = lane; e < N * nc; e += 32) {
        const int i = e / nc, c = e - i * nc;
        double s = sQv[i * NR + c];
#pragma unroll
        for (int l = 0; l < M; ++l) s += Qx[l * N + i] * kt[l * NR + c];
        pt[i * NR + c] = s;
      }
      __syncwarp();
    }
    // ---- forward: response of the state to every right-hand-side column, KEPT for all t (xsT); the terminal one gives nu
    if (with_matrix)
      for (int e = lane; e < N * NR; e += 32) xsT[e] = 0.0;                 // x_0 = 0 in every column
    __syncwarp();
    for (int t = 0; t < T - 1; ++t) {
      const double* Kt = Km + (size_t)t * M * N;
      const double* kt = kk + (size_t)t * M * NR;
      const double* xs = xsT + (size_t)t * N * NR;
      double* xn = xsT + (size_t)(t + 1) * N * NR;
      for (int e = lane; e < M * nc; e += 32) {
        const int i = e / nc, c = e - i * nc;
        double s = kt[i * NR + c];
#pragma unroll
        for (int l = 0; l < N; ++l) s += Kt[i * N + l] * xs[l * NR + c];
        sUs[i * NR + c] = s;
      }
      __syncwarp();
      for (int e = lane; e < N * nc; e += 32) {
        const int i = e / nc, c = e - i * nc;
        double s = (c == 0) ? ct[(size_t)t * N + i] : 0.0;
#pragma unroll
        for (int l = 0; l < N; ++l) s += Am[i * N + l] * xs[l * NR + c];
#pragma unroll
        for (int l = 0; l < M; ++l) s += Bm[i * M + l] * sUs[l * NR + c];
        xn[i * NR + c] = s;
      }
      __syncwarp();
    }
    {
      const double* xs = xsT + (size_t)(T - 1) * N * NR;
      // solve  Mt nu = -xs[:,0], Mt = xs[:,1:]   (terminal step must be zero: d_{T-1} is exact from the start), partial pivoting
      if (lane == 0) {
        double Aug[N][N + 1], nu[N];
        for (int i = 0; i < N; ++i) { for (int j = 0; j < N; ++j) Aug[i][j] = xs[i * NR + 1 + j]; Aug[i][N] = -xs[i * NR]; }
        for (int c = 0; c < N; ++c) {
          int pv = c; double best = fabs(Aug[c][c]);
          for (int r = c + 1; r < N; ++r) if (fabs(Aug[r][c]) > best) { best = fabs(Aug[r][c]); pv = r; }
          if (pv != c) for (int j = 0; j <= N; ++j) { const double tt = Aug[c][j]; Aug[c][j] = Aug[pv][j]; Aug[pv][j] = tt; }
          const double dgn = (best > 1e-300) ? Aug[c][c] : 1e-300;
          for (int r = c + 1; r < N; ++r) { const double f = Aug[r][c] / dgn; for (int j = c; j <= N; ++j) Aug[r][j] -= f * Aug[c][j]; }
          Aug[c][c] = dgn;
        }
        for (int i = N - 1; i >= 0; --i) { double v = Aug[i][N]; for (int j = i + 1; j < N; ++j) v -= Aug[i][j] * nu[j]; nu[i] = v / Aug[i][i]; }
        for (int i = 0; i < N; ++i) sNu[i] = nu[i];
      }
      __syncwarp();
    }
    // ---- superposition, parallel over t (no sequential rollout): d_t = xs_t[:,0] + xs_t[:,1:] nu, then
    //      w_t = k_t[:,0] + k_t[:,1:] nu + K_t d_t  and  pi_t = p_t[:,0] + p_t[:,1:] nu + P_t d_t
    {
      double nu[N];
#pragma unroll
      for (int i = 0; i < N; ++i) nu[i] = sNu[i];
      for (int e = lane; e < T * N; e += 32) {
        const double* xr = xsT + (size_t)e * NR;            // row (t, i) of the stored responses
        double s = xr[0];
#pragma unroll
        for (int c = 0; c < N; ++c) s += xr[1 + c] * nu[c];
        out_d[e] = s;
      }
      __syncwarp();
      for (int e = lane; e < T * (M + N); e += 32) {
        const int t = e / (M + N), r = e - t * (M + N);
        const double* xc = out_d + t * N;
        if (r < M) {
          double s = 0.0;
          if (t < T - 1) {
            const double* kt = kk + (size_t)t * M * NR + r * NR;
            const double* Kt = Km + (size_t)t * M * N + r * N;
            s = kt[0];
#pragma unroll
            for (int c = 0; c < N; ++c) s += kt[1 + c] * nu[c];
#pragma unroll
            for (int l = 0; l < N; ++l) s += Kt[l] * xc[l];
          }
          out_w[t * M + r] = s;
        } else {
          const int i = r - M;
          const double* pt = pp + (size_t)t * N * NR + i * NR;
          const double* Pt = Pm + (size_t)t * N * N + i * N;
          double s = pt[0];
#pragma unroll
          for (int c = 0; c < N; ++c) s += pt[1 + c] * nu[c];
#pragma unroll
          for (int l = 0; l < N; ++l) s += Pt[l] * xc[l];
          pin[t * N + i] = s;
        }
      }
      __syncwarp();
    }
  };

  // ---- IPM iterations ---------------------------------------------------------------------------------------------
  // pass modes: 0 = residuals + Hessians + predictor gradient; 1 = affine statistics; 2 = corrector gradient;
  //             3 = step lengths; 4 = apply
  for (it = 0; it < max_iter; ++it) {
    double sigmu = 0.0, al_p = 0.0, al_d = 0.0;
    for (int mode = 0; mode <= 4; ++mode) {
      if (mode >= 2) sigmu = gl[3];
      if (mode == 4) { al_p = gl[1]; al_d = gl[2]; }
      double pr[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) pr[i] = 0.0;
      // pr: 0 comp, 1 rp_inf, 2 -, 3 objective, 4 qp (max -ds/s), 5 qd (max -dl/l), 6 sum ds*l, 7 sum s*dl, 8 sum ds*dl, 9 slack part of the objective
      for (int t = tid; t < T; t += nthr) {
        const bool stage = (t < T - 1);                // has a control and trust rows
        const bool act = (t >= 1 && t < T - 1);        // has box / collision rows
        double wt[M], dt_[N];
#pragma unroll
        for (int i = 0; i < N; ++i) dt_[i] = d[t * N + i];
#pragma unroll
        for (int j = 0; j < M; ++j) wt[j] = stage ? w[t * M + j] : 0.0;
        const double* dA = dda + t * N; const double* wA = dwa + t * M;
        const double* dZ = dd + t * N;  const double* wZ = dw + t * M;
        double Rl[M][M], rvl[M], Ql[DC][DC], qvl[N];
#pragma unroll
        for (int i = 0; i < M; ++i) { rvl[i] = 0.0;
#pragma unroll
          for (int j = 0; j < M; ++j) Rl[i][j] = (i == j) ? 2.0 * cw : 0.0; }
#pragma unroll
        for (int i = 0; i < N; ++i) qvl[i] = 0.0;
#pragma unroll
        for (int i = 0; i < DC; ++i)
#pragma unroll
          for (int j = 0; j < DC; ++j) Ql[i][j] = 0.0;
        // generic row update: returns tau for the gradient; handles statistics / apply
        auto row = [&](double* ps, double* pl, double gz_h, double gda, double gdz, double& tau, double& wgt) {
          const double s = *ps, l = *pl, rs = 1.0 / s;
          const double rp = gz_h + s;
          wgt = l * rs;
          if (mode == 0) { pr[0] += s * l; pr[1] = fmax(pr[1], fabs(rp)); tau = wgt * rp; return; }
          const double dsa = -rp - gda, dla = -l - wgt * dsa;
          if (mode == 1) { pr[4] = fmax(pr[4], -dsa * rs); pr[5] = fmax(pr[5], -dla / l); pr[6] += dsa * l; pr[7] += s * dla; pr[8] += dsa * dla; return; }
          const double c2 = dsa * dla;
          if (mode == 2) { tau = (sigmu - c2 + l * rp) * rs; return; }
          const double ds = -rp - gdz, dl = -l + (sigmu - c2) * rs - wgt * ds;
          if (mode == 3) { pr[4] = fmax(pr[4], -ds * rs); pr[5] = fmax(pr[5], -dl / l); return; }
          *ps = s + al_p * ds; *pl = l + al_d * dl;
        };
        if (stage) {
          // trust rows
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            double f = -trr, fa = 0.0, fz = 0.0;
#pragma unroll
            for (int j = 0; j < M; ++j) { const double sg = ((e >> j) & 1) ? -1.0 : 1.0; f += sg * wt[j]; fa += sg * wA[j]; fz += sg * wZ[j]; }
            double tau = 0.0, wgt = 0.0;
            row(sT + (size_t)e * T + t, lamT + (size_t)e * T + t, f, fa, fz, tau, wgt);
            if (mode == 0 || mode == 2) {
#pragma unroll
              for (int i = 0; i < M; ++i) {
                const double si = ((e >> i) & 1) ? -1.0 : 1.0;
                rvl[i] += tau * si;
                if (mode == 0) {
#pragma unroll
                  for (int j = 0; j < M; ++j) Rl[i][j] += wgt * si * (((e >> j) & 1) ? -1.0 : 1.0);
                }
              }
            }
          }
        }
        if (act) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double p = x[t * N + i] + dt_[i];
            double tau = 0.0, wgt = 0.0;
            row(sB + (size_t)i * T + t, lamB + (size_t)i * T + t, p - hi[i], dA[i], dZ[i], tau, wgt);
            if (mode == 0 || mode == 2) { qvl[i] += tau; if (mode == 0) Ql[i][i] += wgt; }
            row(sB + (size_t)(2 + i) * T + t, lamB + (size_t)(2 + i) * T + t, lo[i] - p, -dA[i], -dZ[i], tau, wgt);
            if (mode == 0 || mode == 2) { qvl[i] -= tau; if (mode == 0) Ql[i][i] += wgt; }
          }
          if (nq) {
            // collision rows with the shared slack xi_t (eliminated): rows  h - g.d - xi <= 0  and  -xi <= 0
            const double xi = xiv[t];
            double Wsum = 0.0, tsum = 0.0, gW[DC], gT[DC], gWa = 0.0, gWz = 0.0, Hc[DC][DC];
#pragma unroll
            for (int c = 0; c < DC; ++c) { gW[c] = 0.0; gT[c] = 0.0;
#pragma unroll
              for (int e2 = 0; e2 < DC; ++e2) Hc[c][e2] = 0.0; }
            // first sweep: weights and sums (needed before the per-row steps because dxi couples the rows)
            // affine quantities need dxi_aff, final ones dxi: both require sum_q W_q g_q.dd and the tau sums -> two sweeps
            double sumWgda = 0.0, sumWgdz = 0.0, rhs_aff = -cS, rhs_fin = -cS;
            const double s0v = s0[t], l0v = lam0[t], r0 = -xi + s0v, W0 = l0v / s0v;
            for (int q = 0; q < nq; ++q) {
              const double s = sC[(size_t)q * T + t], l = lamC[(size_t)q * T + t], wq = l / s;
              double gd = 0.0, gda = 0.0, gdz = 0.0, gq[DC];
#pragma unroll
              for (int c = 0; c < DC; ++c) { gq[c] = G(q, c, t); gd += gq[c] * dt_[c]; gda += gq[c] * dA[c]; gdz += gq[c] * dZ[c]; }
              const double rC = (H(q, t) - gd) - xi + s;
              Wsum += wq; sumWgda += wq * gda; sumWgdz += wq * gdz;
              rhs_aff += wq * rC;                                  // tau_aff = lam*r/s
#pragma unroll
              for (int c = 0; c < DC; ++c) gW[c] += wq * gq[c];
              if (mode == 0) {
                pr[0] += s * l; pr[1] = fmax(pr[1], fabs(rC));
#pragma unroll
                for (int c = 0; c < DC; ++c)
#pragma unroll
                  for (int e2 = 0; e2 < DC; ++e2) Hc[c][e2] += wq * gq[c] * gq[e2];
              }
            }
            Wsum += W0; rhs_aff += W0 * r0;
            if (mode == 0) { pr[0] += s0v * l0v; pr[1] = fmax(pr[1], fabs(r0)); pr[3] += cS * xi; pr[9] += cS * xi; }
            const double rW = 1.0 / Wsum;
            const double dxia = (rhs_aff - sumWgda) * rW;
            // second sweep: per-row steps
            double tau0 = 0.0;
            {
              const double ds0a = -r0 + dxia, dl0a = -l0v - W0 * ds0a, c0 = ds0a * dl0a;
              if (mode == 1) { pr[4] = fmax(pr[4], -ds0a / s0v); pr[5] = fmax(pr[5], -dl0a / l0v); pr[6] += ds0a * l0v; pr[7] += s0v * dl0a; pr[8] += c0; }
              tau0 = (mode == 0) ? W0 * r0 : (sigmu - c0 + l0v * r0) / s0v;
            }
            rhs_fin += tau0;
            // taus of the collision rows for the gradient (mode 0: affine taus; mode >= 2: corrector taus)
            for (int q = 0; q < nq; ++q) {
              const double s = sC[(size_t)q * T + t], l = lamC[(size_t)q * T + t], wq = l / s;
              double gd = 0.0, gda = 0.0, gq[DC];
#pragma unroll
              for (int c = 0; c < DC; ++c) { gq[c] = G(q, c, t); gd += gq[c] * dt_[c]; gda += gq[c] * dA[c]; }
              const double rC = (H(q, t) - gd) - xi + s;
              const double dsa = -rC + gda + dxia, dla = -l - wq * dsa, c2 = dsa * dla;
              if (mode == 1) { pr[4] = fmax(pr[4], -dsa / s); pr[5] = fmax(pr[5], -dla / l); pr[6] += dsa * l; pr[7] += s * dla; pr[8] += c2; }
              const double tau = (mode == 0) ? wq * rC : (sigmu - c2 + l * rC) / s;
              rhs_fin += tau;
#pragma unroll
              for (int c = 0; c < DC; ++c) gT[c] += tau * gq[c];
            }
            if (mode == 0 || mode == 2) {
              // reduced gradient / Hessian in d_t[0:DC]:  -sum tau g + gW rhs_xi / Wsum ;  sum W g g' - gW gW' / Wsum
#pragma unroll
              for (int c = 0; c < DC; ++c) {
                qvl[c] += -gT[c] + gW[c] * rhs_fin * rW;
                if (mode == 0) {
#pragma unroll
                  for (int e2 = 0; e2 < DC; ++e2) Ql[c][e2] += Hc[c][e2] - gW[c] * gW[e2] * rW;
                }
              }
            }
            if (mode >= 3) {
              const double dxi = (rhs_fin - sumWgdz) * rW;
              const double ds0a = -r0 + dxia, dl0a = -l0v - W0 * ds0a, c0 = ds0a * dl0a;
              const double ds0 = -r0 + dxi, dl0 = -l0v + (sigmu - c0) / s0v - W0 * ds0;
              if (mode == 3) { pr[4] = fmax(pr[4], -ds0 / s0v); pr[5] = fmax(pr[5], -dl0 / l0v); }
              else { s0[t] = s0v + al_p * ds0; lam0[t] = l0v + al_d * dl0; xiv[t] = xi + al_p * dxi; }
              for (int q = 0; q < nq; ++q) {
                const double s = sC[(size_t)q * T + t], l = lamC[(size_t)q * T + t], wq = l / s;
                double gd = 0.0, gda = 0.0, gdz = 0.0;
#pragma unroll
                for (int c = 0; c < DC; ++c) { const double g = G(q, c, t); gd += g * dt_[c]; gda += g * dA[c]; gdz += g * dZ[c]; }
                const double rC = (H(q, t) - gd) - xi + s;
                const double dsa = -rC + gda + dxia, dla = -l - wq * dsa, c2 = dsa * dla;
                const double ds = -rC + gdz + dxi, dl = -l + (sigmu - c2) / s - wq * ds;
                if (mode == 3) { pr[4] = fmax(pr[4], -ds / s); pr[5] = fmax(pr[5], -dl / l); }
                else { sC[(size_t)q * T + t] = s + al_p * ds; lamC[(size_t)q * T + t] = l + al_d * dl; }
              }
            }
          }
        }
        if (mode == 0 || mode == 2) {
          // objective gradient and (mode 0) Hessian, equality residual, stationarity
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            qvl[i] += LIN(t, i) + rho * (dt_[i] - SBAR(t, i));
            if (mode == 0) { Ql[i][i] += rho; pr[3] += LIN(t, i) * dt_[i] + 0.5 * rho * (dt_[i] - SBAR(t, i)) * (dt_[i] - SBAR(t, i)); }
          }
          if (stage) {
#pragma unroll
            for (int j = 0; j < M; ++j) {
              rvl[j] += 2.0 * cw * (u[t * M + j] + wt[j]);
              if (mode == 0) pr[3] += cw * (u[t * M + j] + wt[j]) * (u[t * M + j] + wt[j]);
            }
          }
#pragma unroll
          for (int i = 0; i < N; ++i) qv[(size_t)t * N + i] = qvl[i];
#pragma unroll
          for (int j = 0; j < M; ++j) rv[(size_t)t * M + j] = rvl[j];
          if (mode == 0) {
#pragma unroll
            for (int i = 0; i < DC; ++i)
#pragma unroll
              for (int j = 0; j < DC; ++j) Qp[(size_t)t * DC * DC + i * DC + j] = Ql[i][j];
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
              for (int j = 0; j < M; ++j) Rm[(size_t)t * M * M + i * M + j] = Rl[i][j];
            if (stage) {
              // equality residual res = d_{t+1} - (A d_t + B w_t + c_t), c_t = A x_t + B u_t - x_{t+1};  ctil = -res
#pragma unroll
              for (int i = 0; i < N; ++i) {
                double s = -x[(t + 1) * N + i];
#pragma unroll
                for (int l = 0; l < N; ++l) s += Am[i * N + l] * (x[t * N + l] + dt_[l]);
#pragma unroll
                for (int l = 0; l < M; ++l) s += Bm[i * M + l] * (u[t * M + l] + wt[l]);
                const double res = d[(t + 1) * N + i] - s;
                ct[(size_t)t * N + i] = -res;
                pr[1] = fmax(pr[1], fabs(res));
              }
            }
          }
        }
        if (mode == 4) {
#pragma unroll
          for (int i = 0; i < N; ++i) { d[t * N + i] = dt_[i] + al_p * dZ[i]; pi[t * N + i] += al_d * (pin[t * N + i] - pi[t * N + i]); }
          if (stage) {
#pragma unroll
            for (int j = 0; j < M; ++j) w[t * M + j] = wt[j] + al_p * wZ[j];
          }
        }
      }
      // ---- epilogues
      if (mode == 0) {
        const int ops[10] = {0, 2, 2, 0, 0, 0, 0, 0, 0, 0};
        block_reduce<10>(pr, ops, red);
        if (tid == 0) {
          const double comp = red[0], rp = red[1], obj = red[3];
          // gap and stagnation are measured against the SMOOTH part of the objective: when the collision slack (cost 1e4
          // per unit) dominates, a tolerance relative to the full value would leave w undetermined to ~1e-2
          const double slack = red[9], smooth = obj - slack, scale = fmax(fabs(smooth), 1.0);
          const bool stagn = fabs(smooth - gl[8]) <= 1e-9 * scale && fabs(slack - gl[9]) <= 1e-10 * fmax(fabs(slack), 1.0);
          gl[8] = smooth; gl[9] = slack;
          gl[4] = comp; gl[5] = comp / n_rows;
          // optimal: complementarity gap and primal feasibility at tolerance AND a stagnated objective (on saturated
          // problems the multipliers lose accuracy long before the primal point does, so dual feasibility is not the test;
          // tests certify the result with an exact LP bracket)
          int flag = 0;
          if (!(comp == comp) || !(rp == rp)) flag = 2;
          else if (it > 0 && comp <= 1e-9 * scale && rp <= 1e-9 && stagn) flag = 1;
          else if (it > 0 && comp <= 1e-14 * scale && rp <= 1e-9) flag = 1;     // past what fp64 can use
          gl[6] = obj; gl[0] = (double)flag;
          // remember the best primal-feasible iterate: past convergence the Newton systems become unusable (W = lam/s
          // overflows) and later iterates can only get worse
          const bool take = (flag != 2) && rp <= 1e-9 && (gl[10] == 0.0 || obj < gl[11]);
          gl[12] = take ? 1.0 : 0.0;
          if (take) { gl[10] = 1.0; gl[11] = obj; gl[13] = comp; }
        }
        __syncthreads();
        if (gl[12] != 0.0) {
          for (int t = tid; t < T; t += nthr) {
#pragma unroll
            for (int i = 0; i < N; ++i) dbest[t * N + i] = d[t * N + i];
#pragma unroll
            for (int j = 0; j < M; ++j) wbest[t * M + j] = w[t * M + j];
          }
        }
        obj_prev = gl[6];
        if ((int)gl[0] == 1) { status = SCVX_ST_OPTIMAL; break; }
        if ((int)gl[0] == 2) { status = SCVX_ST_NUMERICAL; break; }
        if (tid < 32) riccati(true, dda, dwa);
        __syncthreads();
      } else if (mode == 1) {
        const int ops[9] = {0, 0, 0, 0, 2, 2, 0, 0, 0};
        block_reduce<9>(pr, ops, red);
        if (tid == 0) {
          const double ap = (red[4] > 1.0) ? 1.0 / red[4] : 1.0, ad = (red[5] > 1.0) ? 1.0 / red[5] : 1.0;
          const double comp = gl[4];
          double sg = (comp + ap * red[6] + ad * red[7] + ap * ad * red[8]) / comp;
          sg = fmin(fmax(sg, 0.0), 1.0);
          gl[3] = sg * sg * sg * gl[5];
        }
        __syncthreads();
      } else if (mode == 2) {
        __syncthreads();
        if (tid < 32) riccati(false, dd, dw);
        __syncthreads();
      } else if (mode == 3) {
        double bad = 0.0;
        for (int t = tid; t < T; t += nthr)
#pragma unroll
          for (int i = 0; i < N; ++i) bad = fmax(bad, isfinite(dd[t * N + i]) ? 0.0 : 1.0);
        pr[6] = bad;
        const int ops[9] = {0, 0, 0, 0, 2, 2, 2, 0, 0};
        block_reduce<9>(pr, ops, red);
        if (tid == 0) {
          const double ap = (red[4] > 1.0) ? 1.0 / red[4] : 1.0, ad = (red[5] > 1.0) ? 1.0 / red[5] : 1.0;
          gl[1] = fmin(1.0, 0.995 * ap); gl[2] = fmin(1.0, 0.995 * ad);
          gl[0] = (red[6] > 0.0 || !isfinite(red[4]) || !isfinite(red[5])) ? 2.0 : 0.0;
        }
        __syncthreads();
        if ((int)gl[0] == 2) break;
      } else {
        __syncthreads();
      }
    }
    if ((int)gl[0] == 2) { status = SCVX_ST_NUMERICAL; break; }
    if (status == SCVX_ST_OPTIMAL) break;
  }

  // ---- not converged by the tests above: fall back to the best primal-feasible iterate -------------------------------
  __syncthreads();
  if (status != SCVX_ST_OPTIMAL && gl[10] != 0.0) {
    for (int t = tid; t < T; t += nthr) {
#pragma unroll
      for (int i = 0; i < N; ++i) d[t * N + i] = dbest[t * N + i];
#pragma unroll
      for (int j = 0; j < M; ++j) w[t * M + j] = wbest[t * M + j];
    }
    if (gl[13] <= 1e-7 * fmax(fabs(gl[11]), 1.0)) status = SCVX_ST_OPTIMAL;     // its gap was already at tolerance
    __syncthreads();
  }
  // ---- outputs ------------------------------------------------------------------------------------------------
  double pr[2] = {0.0, 0.0};
  for (int t = tid; t < T; t += nthr) {
#pragma unroll
    for (int i = 0; i < N; ++i) a.d[((size_t)robot * T + t) * N + i] = d[t * N + i];
#pragma unroll
    for (int j = 0; j < M; ++j) a.w[((size_t)robot * T + t) * M + j] = (t < T - 1) ? w[t * M + j] : 0.0;
    double S = 0.0;
    if (nq && t < T - 1) {
      for (int q = 0; q < nq; ++q) {
        double gd = 0.0;
#pragma unroll
        for (int c = 0; c < DC; ++c) gd += G(q, c, t) * d[t * N + c];
        S = fmax(S, H(q, t) - gd);
      }
    }
    if (a.S) a.S[(size_t)robot * T + t] = S;
    pr[0] += cS * S;
#pragma unroll
    for (int i = 0; i < 2; ++i) pr[0] += LIN(t, i) * d[t * N + i] + 0.5 * rho * (d[t * N + i] - SBAR(t, i)) * (d[t * N + i] - SBAR(t, i));
    if (t < T - 1)
#pragma unroll
      for (int j = 0; j < M; ++j) pr[0] += cw * (u[t * M + j] + w[t * M + j]) * (u[t * M + j] + w[t * M + j]);
  }
  const int ops[2] = {0, 0};
  block_reduce<2>(pr, ops, red);
  if (tid == 0) { a.objective[robot] = red[0]; a.status[robot] = status; a.iters[robot] = it; }
}

// ---- per-(robot, t) consensus QP, EXACT for small nq: KKT enumeration --------------------------------------------------
//   min_{sb in R^2, S}  -r.sb + rho/2 |s - sb|^2 + c_S S   s.t.  h_q - g_q.sb - S <= 0 (q < nq),  -S <= 0
// The optimum has at most 3 active rows.  For every candidate active set (<= 3 collision rows, S pinned to 0 or free) the
// KKT equations are linear; the candidate that satisfies primal and dual feasibility is the (unique) solution.  With the
// slack cost 1e6 against rho = 1 an interior-point iteration resolves the point only to ~1e-2 when two opposing rows force
// S > 0; the enumeration is exact to round-off and costs O(nq^3) tiny solves -- used for nq <= 8 (the script has nq = 3).
__device__ __forceinline__ bool solve_small(double (*A)[7], int n, double* x) {
  for (int c = 0; c < n; ++c) {
    int pv = c; double best = fabs(A[c][c]);
    for (int r = c + 1; r < n; ++r) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); pv = r; }
    if (best < 1e-12) return false;
    if (pv != c) for (int j = 0; j <= n; ++j) { const double t = A[c][j]; A[c][j] = A[pv][j]; A[pv][j] = t; }
    for (int r = c + 1; r < n; ++r) { const double f = A[r][c] / A[c][c]; for (int j = c; j <= n; ++j) A[r][j] -= f * A[c][j]; }
  }
  for (int i = n - 1; i >= 0; --i) { double v = A[i][n]; for (int j = i + 1; j < n; ++j) v -= A[i][j] * x[j]; x[i] = v / A[i][i]; }
  return true;
}

__global__ void __launch_bounds__(128)
sbar_enum_kernel(int n_robots, int T, int nq, double rho, double c_S, const double* __restrict__ s_pos,
                 const double* __restrict__ r_dual, const double* __restrict__ col_h, const double* __restrict__ col_g,
                 double* __restrict__ sbar, double* __restrict__ S_out) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_robots * T) return;
  const int robot = gid / T, t = gid - robot * T;
  const double cx = s_pos[(size_t)gid * 2] + r_dual[(size_t)gid * 2] / rho, cy = s_pos[(size_t)gid * 2 + 1] + r_dual[(size_t)gid * 2 + 1] / rho;
  const double* h = col_h + (size_t)robot * nq * T;
  const double* g = col_g + (size_t)robot * nq * 2 * T;
  auto Hq = [&](int q) { return h[(size_t)q * T + t]; };
  auto Gx = [&](int q) { return g[((size_t)q * 2) * T + t]; };
  auto Gy = [&](int q) { return g[((size_t)q * 2 + 1) * T + t]; };
  double best_obj = 1e300, bx = cx, by = cy, bS = 0.0;
  bool found = false;
  const double tol = 1e-9;
  // active sets: k rows (k = 0..3), indices a < b < c; S pinned (free_S = 0) or free (free_S = 1, needs k >= 1)
  for (int k = 0; k <= 3 && k <= nq; ++k)
    for (int a = 0; a < (k >= 1 ? nq : 1); ++a)
      for (int b = (k >= 2 ? a + 1 : 0); b < (k >= 2 ? nq : 1); ++b)
        for (int c = (k >= 3 ? b + 1 : 0); c < (k >= 3 ? nq : 1); ++c)
          for (int free_S = 0; free_S <= (k >= 1 ? 1 : 0); ++free_S) {
            const int idx[3] = {a, b, c};
            const int n = 2 + free_S + k;                 // unknowns: sbx, sby, [S], lam_1..k
            double A[7][7];
            for (int i = 0; i < n; ++i) for (int j = 0; j <= n; ++j) A[i][j] = 0.0;
            // stationarity in sb: rho (sb - c) - sum lam g = 0
            A[0][0] = rho; A[0][n] = rho * cx; A[1][1] = rho; A[1][n] = rho * cy;
            for (int j = 0; j < k; ++j) { A[0][2 + free_S + j] = -Gx(idx[j]); A[1][2 + free_S + j] = -Gy(idx[j]); }
            int row = 2;
            if (free_S) { for (int j = 0; j < k; ++j) A[row][3 + j] = 1.0; A[row][n] = c_S; ++row; }   // sum lam = c_S
            for (int j = 0; j < k; ++j) {                                                               // g.sb + S = h
              A[row][0] = Gx(idx[j]); A[row][1] = Gy(idx[j]); if (free_S) A[row][2] = 1.0; A[row][n] = Hq(idx[j]); ++row;
            }
            double x[7];
            if (!solve_small(A, n, x)) continue;
            const double sx_ = x[0], sy_ = x[1], S = free_S ? x[2] : 0.0;
            bool ok = S >= -tol;
            double lsum = 0.0;
            for (int j = 0; j < k; ++j) { const double l = x[2 + free_S + j]; ok = ok && (l >= -tol * c_S); lsum += l; }
            if (!free_S) ok = ok && (lsum <= c_S * (1.0 + tol));          // multiplier of S >= 0 is c_S - sum lam >= 0
            for (int q = 0; q < nq && ok; ++q) ok = (Hq(q) - (Gx(q) * sx_ + Gy(q) * sy_) - S <= tol * fmax(1.0, fabs(Hq(q))));
            if (!ok) continue;
            const double obj = 0.5 * rho * ((sx_ - cx) * (sx_ - cx) + (sy_ - cy) * (sy_ - cy)) + c_S * fmax(S, 0.0);
            if (obj < best_obj) { best_obj = obj; bx = sx_; by = sy_; bS = fmax(S, 0.0); found = true; }
          }
  (void)found;
  sbar[(size_t)gid * 2] = bx; sbar[(size_t)gid * 2 + 1] = by;
  if (S_out) S_out[gid] = bS;
}

// ---- per-(robot, t) consensus QP of ADMM_decentralized.py:106-139 ---------------------------------------------------
//   min_{sb in R^2, S >= 0}  r.(s - sb) + rho/2 |s - sb|^2 + c_S S   s.t.  h_q - g_q . sb <= S  (q = 0..nq-1)
// One thread per (robot, t): a tiny primal-dual interior-point method with S eliminated (2x2 Newton systems).
__global__ void __launch_bounds__(128)
sbar_qp_kernel(int n_robots, int T, int nq, double rho, double c_S, const double* __restrict__ s_pos,
               const double* __restrict__ r_dual, const double* __restrict__ col_h, const double* __restrict__ col_g,
               double* __restrict__ sbar, double* __restrict__ S_out, double* __restrict__ ws) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_robots * T) return;
  const int robot = gid / T, t = gid - robot * T;
  const double sx = s_pos[(size_t)gid * 2], sy = s_pos[(size_t)gid * 2 + 1];
  const double rx = r_dual[(size_t)gid * 2], ry = r_dual[(size_t)gid * 2 + 1];
  const double* h = col_h + (size_t)robot * nq * T;          // [nq][T]
  const double* g = col_g + (size_t)robot * nq * 2 * T;      // [nq][2][T]
  double* lam = ws + (size_t)gid * 2 * nq;                   // [nq] multipliers, [nq] slacks   (thread-private, strided)
  double* sl = lam + nq;
  // unconstrained minimiser: sb = s + r/rho
  double bx = sx + rx / rho, by = sy + ry / rho;
  const double mu0 = fmax(1.0, c_S);
  // start: xi from the dual-feasible central equation (bisection), rows h - g.sb - xi + s = 0
  double vmax = -1e300;
  for (int q = 0; q < nq; ++q) vmax = fmax(vmax, h[(size_t)q * T + t] - (g[((size_t)q * 2) * T + t] * bx + g[((size_t)q * 2 + 1) * T + t] * by));
  const double L = fmax(vmax, 0.0);
  double lo_ = L, hi_ = L + (nq + 1) * mu0 / c_S;
  for (int itb = 0; itb < 60; ++itb) {
    const double mid = 0.5 * (lo_ + hi_);
    double f = mu0 / mid - c_S;
    for (int q = 0; q < nq; ++q) f += mu0 / (mid - (h[(size_t)q * T + t] - (g[((size_t)q * 2) * T + t] * bx + g[((size_t)q * 2 + 1) * T + t] * by)));
    if (f > 0.0) lo_ = mid; else hi_ = mid;
  }
  double xi = hi_, s0 = xi, l0 = mu0 / xi;
  for (int q = 0; q < nq; ++q) {
    const double v = h[(size_t)q * T + t] - (g[((size_t)q * 2) * T + t] * bx + g[((size_t)q * 2 + 1) * T + t] * by);
    sl[q] = xi - v; lam[q] = mu0 / sl[q];
  }
  const int nr = nq + 1;
  for (int it = 0; it < 80; ++it) {
    // residuals / sums
    double comp = s0 * l0, rpm = fabs(-xi + s0);
    double Wsum = l0 / s0, H00 = rho, H01 = 0.0, H11 = rho, gWx = 0.0, gWy = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double gx = g[((size_t)q * 2) * T + t], gy = g[((size_t)q * 2 + 1) * T + t];
      const double rC = h[(size_t)q * T + t] - (gx * bx + gy * by) - xi + sl[q], wq = lam[q] / sl[q];
      comp += sl[q] * lam[q]; rpm = fmax(rpm, fabs(rC));
      Wsum += wq; gWx += wq * gx; gWy += wq * gy; H00 += wq * gx * gx; H01 += wq * gx * gy; H11 += wq * gy * gy;
    }
    // stationarity with the current multipliers: -r - rho (s - sb) - sum lam_q g_q = 0,  c_S - sum lam_q - lam_0 = 0
    double gx_ = -rx - rho * (sx - bx), gy_ = -ry - rho * (sy - by), gxi = c_S - l0;
    for (int q = 0; q < nq; ++q) {
      gx_ -= lam[q] * g[((size_t)q * 2) * T + t]; gy_ -= lam[q] * g[((size_t)q * 2 + 1) * T + t]; gxi -= lam[q];
    }
    const double rd = fmax(fmax(fabs(gx_), fabs(gy_)), fabs(gxi) * 1e-6);      // the slack row is scaled by its cost 1e6
    // absolute tolerances: the slack term (cost 1e6 per unit) must not swamp the quadratic that determines sbar
    if (it > 0 && comp <= 1e-10 && rpm <= 1e-10 && rd <= 1e-8) break;
    if (comp <= 1e-15) break;                       // past what fp64 can use: further steps only degrade the point
    const double rW = 1.0 / Wsum, mu = comp / nr;
    H00 -= gWx * gWx * rW; H01 -= gWx * gWy * rW; H11 -= gWy * gWy * rW;
    const double det = H00 * H11 - H01 * H01;
    double dbx = 0.0, dby = 0.0, dxi = 0.0, dxia = 0.0, dbxa = 0.0, dbya = 0.0, sigmu = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      // gradient with taus (pass 0: affine, pass 1: corrector using the affine step)
      const double r0 = -xi + s0, W0 = l0 / s0;
      double c0 = 0.0;
      if (pass == 1) { const double ds0a = -r0 + dxia, dl0a = -l0 - W0 * ds0a; c0 = ds0a * dl0a; }
      double rhs_xi = -c_S + ((pass == 0) ? W0 * r0 : (sigmu - c0 + l0 * r0) / s0);
      double gTx = 0.0, gTy = 0.0;
      for (int q = 0; q < nq; ++q) {
        const double gx = g[((size_t)q * 2) * T + t], gy = g[((size_t)q * 2 + 1) * T + t];
        const double rC = h[(size_t)q * T + t] - (gx * bx + gy * by) - xi + sl[q], wq = lam[q] / sl[q];
        double tau;
        if (pass == 0) tau = wq * rC;
        else { const double dsa = -rC + gx * dbxa + gy * dbya + dxia, dla = -lam[q] - wq * dsa; tau = (sigmu - dsa * dla + lam[q] * rC) / sl[q]; }
        rhs_xi += tau; gTx += tau * gx; gTy += tau * gy;
      }
      // gradient of the smooth part wrt sb: -r - rho (s - sb)
      const double qx = -rx - rho * (sx - bx) - gTx + gWx * rhs_xi * rW;
      const double qy = -ry - rho * (sy - by) - gTy + gWy * rhs_xi * rW;
      const double ddx = -(H11 * qx - H01 * qy) / det, ddy = -(-H01 * qx + H00 * qy) / det;
      const double dx_i = (rhs_xi - (gWx * ddx + gWy * ddy)) * rW;
      if (pass == 0) {
        dbxa = ddx; dbya = ddy; dxia = dx_i;
        // affine step lengths and mu_aff
        double qp = 0.0, qd = 0.0, s2l = 0.0, sdl = 0.0, ddl = 0.0;
        { const double ds = -r0 + dxia, dl = -l0 - W0 * ds; qp = fmax(qp, -ds / s0); qd = fmax(qd, -dl / l0); s2l += ds * l0; sdl += s0 * dl; ddl += ds * dl; }
        for (int q = 0; q < nq; ++q) {
          const double gx = g[((size_t)q * 2) * T + t], gy = g[((size_t)q * 2 + 1) * T + t];
          const double rC = h[(size_t)q * T + t] - (gx * bx + gy * by) - xi + sl[q], wq = lam[q] / sl[q];
          const double ds = -rC + gx * dbxa + gy * dbya + dxia, dl = -lam[q] - wq * ds;
          qp = fmax(qp, -ds / sl[q]); qd = fmax(qd, -dl / lam[q]); s2l += ds * lam[q]; sdl += sl[q] * dl; ddl += ds * dl;
        }
        const double ap = (qp > 1.0) ? 1.0 / qp : 1.0, ad = (qd > 1.0) ? 1.0 / qd : 1.0;
        double sg = (comp + ap * s2l + ad * sdl + ap * ad * ddl) / comp;
        sg = fmin(fmax(sg, 0.0), 1.0);
        sigmu = sg * sg * sg * mu;
      } else { dbx = ddx; dby = ddy; dxi = dx_i; }
    }
    // final step lengths and update
    double qp = 0.0, qd = 0.0;
    const double r0 = -xi + s0, W0 = l0 / s0;
    const double ds0a = -r0 + dxia, dl0a = -l0 - W0 * ds0a;
    const double ds0 = -r0 + dxi, dl0 = -l0 + (sigmu - ds0a * dl0a) / s0 - W0 * ds0;
    qp = fmax(qp, -ds0 / s0); qd = fmax(qd, -dl0 / l0);
    for (int q = 0; q < nq; ++q) {
      const double gx = g[((size_t)q * 2) * T + t], gy = g[((size_t)q * 2 + 1) * T + t];
      const double rC = h[(size_t)q * T + t] - (gx * bx + gy * by) - xi + sl[q], wq = lam[q] / sl[q];
      const double dsa = -rC + gx * dbxa + gy * dbya + dxia, dla = -lam[q] - wq * dsa;
      const double ds = -rC + gx * dbx + gy * dby + dxi, dl = -lam[q] + (sigmu - dsa * dla) / sl[q] - wq * ds;
      qp = fmax(qp, -ds / sl[q]); qd = fmax(qd, -dl / lam[q]);
    }
    const double ap = fmin(1.0, 0.995 * ((qp > 1.0) ? 1.0 / qp : 1.0)), ad = fmin(1.0, 0.995 * ((qd > 1.0) ? 1.0 / qd : 1.0));
    if (!isfinite(dbx) || !isfinite(dby) || !isfinite(dxi)) break;
    for (int q = 0; q < nq; ++q) {
      const double gx = g[((size_t)q * 2) * T + t], gy = g[((size_t)q * 2 + 1) * T + t];
      const double rC = h[(size_t)q * T + t] - (gx * bx + gy * by) - xi + sl[q], wq = lam[q] / sl[q];
      const double dsa = -rC + gx * dbxa + gy * dbya + dxia, dla = -lam[q] - wq * dsa;
      const double ds = -rC + gx * dbx + gy * dby + dxi, dl = -lam[q] + (sigmu - dsa * dla) / sl[q] - wq * ds;
      sl[q] += ap * ds; lam[q] += ad * dl;
    }
    s0 += ap * ds0; l0 += ad * dl0;
    bx += ap * dbx; by += ap * dby; xi += ap * dxi;
  }
  sbar[(size_t)gid * 2] = bx; sbar[(size_t)gid * 2 + 1] = by;
  double S = 0.0;
  for (int q = 0; q < nq; ++q) S = fmax(S, h[(size_t)q * T + t] - (g[((size_t)q * 2) * T + t] * bx + g[((size_t)q * 2 + 1) * T + t] * by));
  if (S_out) S_out[gid] = S;
}

template <int N, int M>
size_t lti_smem_bytes(int T) {
  return ((size_t)T * LtiDims<N, M>::PER_T + 32 + 9 * 12 + N * N + N * M + LtiDims<N, M>::SCR) * sizeof(double);
}

template <int N, int M>
int launch_lti(const scvx_lti_args& a, cudaStream_t st) {
  const size_t smem = lti_smem_bytes<N, M>(a.T);
  if (smem > 227 * 1024) {
    snprintf(g_last_error, sizeof(g_last_error), "T=%d needs %zu B of shared memory per robot (> 227 KB)", a.T, smem);
    return SCVX_E_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(lti_qp_kernel<N, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  lti_qp_kernel<N, M><<<a.n_robots, 128, smem, st>>>(a);
  SCVX_CHECK_LAUNCH("scvx_lti_qp_batched");
  return SCVX_OK;
}

}  // namespace scvx

using namespace scvx;

extern "C" unsigned long long scvx_lti_qp_workspace_bytes(int n_robots, int T, int n, int nq) {
  if (n_robots < 0 || T < 3 || nq < 0) return 0ull;
  if (n == 4) return (unsigned long long)n_robots * lti_ws_doubles<4, 2>(T, nq) * sizeof(double);
  if (n == 6) return (unsigned long long)n_robots * lti_ws_doubles<6, 3>(T, nq) * sizeof(double);
  return 0ull;
}

extern "C" int scvx_lti_qp_batched(const scvx_lti_args* a, void* stream) {
  if (!a) return bad_arg("args");
  if (a->n_robots < 0 || a->T < 3 || a->nq < 0) return bad_arg("n_robots/T/nq");
  if (a->n_robots == 0) return SCVX_OK;
  if (!((a->n == 4 && a->m == 2) || (a->n == 6 && a->m == 3))) return bad_arg("(n, m) must be (4, 2) or (6, 3)");
  if (!a->Ad || !a->Bd || !a->x || !a->u || !a->x_des || !a->d || !a->w || !a->objective || !a->status || !a->iters)
    return bad_arg("null pointer");
  if (a->nq > 0 && (!a->col_h || !a->col_g || !(a->c_S > 0.0))) return bad_arg("collision tables / c_S");
  if (!(a->tr > 0.0) || !(a->c_w > 0.0)) return bad_arg("tr and c_w must be positive");
  const unsigned long long need = scvx_lti_qp_workspace_bytes(a->n_robots, a->T, a->n, a->nq);
  if (!a->workspace || a->workspace_bytes < need) {
    snprintf(g_last_error, sizeof(g_last_error), "workspace too small: need %llu bytes", need);
    return SCVX_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return (a->n == 4) ? launch_lti<4, 2>(*a, st) : launch_lti<6, 3>(*a, st);
}

extern "C" int scvx_sbar_qp_batched(int n_robots, int T, int nq, double rho, double c_S, const double* s_pos,
                                    const double* r_dual, const double* col_h, const double* col_g, double* sbar,
                                    double* S_out, void* workspace, unsigned long long workspace_bytes, void* stream) {
  if (n_robots < 0 || T < 1 || nq < 1) return bad_arg("n_robots/T/nq");
  if (n_robots == 0) return SCVX_OK;
  if (!s_pos || !r_dual || !col_h || !col_g || !sbar) return bad_arg("null pointer");
  if (!(rho > 0.0) || !(c_S > 0.0)) return bad_arg("rho and c_S must be positive");
  const unsigned long long need = (unsigned long long)n_robots * T * 2 * nq * sizeof(double);
  if (!workspace || workspace_bytes < need) {
    snprintf(g_last_error, sizeof(g_last_error), "workspace too small: need %llu bytes", need);
    return SCVX_E_WORKSPACE;
  }
  const int total = n_robots * T;
  if (nq <= 8)
    sbar_enum_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_robots, T, nq, rho, c_S, s_pos, r_dual, col_h, col_g, sbar, S_out);
  else
    sbar_qp_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_robots, T, nq, rho, c_S, s_pos, r_dual, col_h, col_g, sbar,
                                                                           S_out, (double*)workspace);
  SCVX_CHECK_LAUNCH("scvx_sbar_qp_batched");
  return SCVX_OK;
}
