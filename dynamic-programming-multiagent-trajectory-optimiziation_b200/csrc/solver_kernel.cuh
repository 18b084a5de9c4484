// solver_kernel.cuh -- stage 3 (template text; instantiated per model in solver_unicycle.cu / solver_si.cu): the convex sub-problem of SCvx, batched, fp64, sm_100a.
//
// Replaces SCProblem.solve (SCvx/optimization/sc_problem.py:15-105; model rows unicycle_model.py:85-115,
// single_integrator_model.py:79-128) and AgentSolver.setup+solve (agent_solver.py:43-117): where the
// reference builds a cvxpy graph and calls ECOS, this kernel runs a structure-exploiting primal-dual
// interior-point method (Mehrotra predictor-corrector), ONE THREAD BLOCK PER AGENT.
//
// Formulation (oracle/ipm_struct.py is the line-by-line CPU twin of this file):
//   z = (w_0..w_{K-1}; sigma, t_nu, t_x, t_u),  w_k = (x_k, u_k);  w_0, w_{K-1} fixed by the boundary conditions
//   min  c_s [ w_nu t_nu + w_sigma sigma + sum_h w_h max(0, b_hk - a_hk.p_k) + rho/2 |P|^2 + <lin, P> ]
//   s.t. e.nu_k(z) <= t_nu, e.(x_k - xref_k) <= t_x, e.(u_k - uref_k) <= t_u   for ALL sign patterns e   (L1 epigraphs
//        without auxiliary variables: |v|_1 <= t  <=>  e.v <= t for every e in {+-1}^n),
//        t_x + t_u +- (sigma - sigma_ref) <= r, sigma >= 0, position box, input box (unicycle) or
//        1/2(|u_k|^2 - v_max^2) <= 0 (single integrator).
//   Each hinge term carries a private slack xi (rows -a.p - xi <= -b, -xi <= 0) that is eliminated analytically
//   from the Newton system, so the KKT matrix is block tridiagonal (n_s x n_s blocks, n_s = n_x + n_u) with a
//   4-column border (sigma, t_nu, t_x, t_u) whatever the number of obstacles / neighbours.
//
// Per IPM iteration: row passes are data-parallel over stages (thread k owns stage k and interval k); the
// block-tridiagonal system is factorised and solved by BLOCK CYCLIC REDUCTION (odd-even elimination, strides
// 1, 2, 4, ...: log2(K) levels, every level data-parallel over (node, column) work items on all threads) with the
// whole factor resident in shared memory; reductions use warp shuffles.  Row state (s, lambda) lives in a global
// workspace laid out [row][k] so that every access is coalesced over k.
// (The first version walked the stages sequentially on warp 0: profiles/r01a shows it latency-bound with 3 of 4
// warps parked at a barrier; cyclic reduction removes the O(K) dependent chain.  profiles/r01g: the four step passes are
// compile-time instantiations, reciprocals are branch-free, hinge rows are fetched four at a time, and the blocks of a launch
// run longest-first -- scvx_solve_args.block_order.)
// Best-response (Nash game) terms: diagonal / linear / consecutive-difference quadratics and sigma == sigma_ref, see
// scvx_solve_args.quad_diag .. fix_sigma in include/scvx_b200.h.
// Round 2, last session: the iteration's rules beyond textbook Mehrotra -- second-order term weighed by the affine step
// length (LPs), step fraction rising to 1 - 1e-6 as the centring target vanishes, warm barrier start (scvx_solve_args.mu0) with a cold
// start where the start point violates a hinge row and a compact retry pass (RETRY instantiation) for the solves that still strand --
// and the template parameters KT / MT that make the sizes of the BASELINE shapes compile-time constants (solver_*_fixed.cu).
#pragma once
#include <cooperative_groups.h>

#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "reduce.cuh"

namespace scvx {

// Optional per-phase cycle counters of block 0 (build with -DSCVX_PHASE_TIMING; read with scvx_debug_phase_cycles).
#ifdef SCVX_PHASE_TIMING
static __device__ unsigned long long g_phase_cycles[32];   // one copy per translation unit (per model)
#define PHASE_INIT() long long ph_t0 = clock64()
#define PHASE(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long ph_t1 = clock64(); g_phase_cycles[i] += (unsigned long long)(ph_t1 - ph_t0); ph_t0 = ph_t1; } } while (0)
#else
#define PHASE_INIT() do { } while (0)
#define PHASE(i) do { } while (0)
#endif

constexpr int SOLVER_MAX_THREADS = 256;
// Slacks are NOT stored: every plain row is linear and the start is strictly feasible, so s = h - G z is recomputed from
// the iterate in every pass (halves the row-state traffic and keeps s consistent with z to round-off).
constexpr double TINY_S = 1e-14;
constexpr int HINGE_CHUNK = 4;     // hinge rows fetched together (see the row passes)

template <class M>
struct Dims {
  static constexpr int NX = M::NX, NU = M::NU, D = M::D, NS = NX + NU;
  static constexpr bool BALL = (NU == 3);                 // single integrator: ||u||_2 <= v_max
  static constexpr int NEX = 1 << NX, NEU = 1 << NU;
  static constexpr int NV = BALL ? 1 : 4;                 // input rows per stage
  static constexpr int NPLAIN = NEX + NEX + NEU + 2 * D + NV;
  static constexpr int R_NU = 0, R_X = NEX, R_U = 2 * NEX, R_P = 2 * NEX + NEU, R_V = R_P + 2 * D;
  static constexpr int NJ = NX * NX + 2 * NX * NU + 2 * NX;    // jacobian doubles per interval (27 / 33, odd)
  static constexpr int NSP = NS | 1;                      // padded stage stride (odd)
  static constexpr int SD = (NS * NS) | 1;                // padded block stride (odd)
  static constexpr int SR = (NS * 4) | 1;                 // padded border stride (odd)
  static constexpr int STG = SD;                          // per-interval staging (16 used) ALIASES factor slot C
  static constexpr int ST2 = 3;                           // corrector staging (e'tau of the interval, sigma-mu coefficient; the rest rides in dW)
  static constexpr int PER_STAGE = 3 * NSP + NJ + 3 * SD + SR + ST2;
  static constexpr int PER_STAGE_NOJAC = 3 * NSP + 3 * SD + SR + ST2;
  static constexpr int SMALL = 64 + 9 * 24;               // globals + reduction scratch (one hinge group: <= 8 warps)
  static constexpr int HSW = (D * D + 2 * D) | 1;         // per-stage partial results of a helper hinge group (odd stride)
  // hinge groups G > 1: reduction scratch for 4 G warps, and the helpers' partial results
  static constexpr int small_of(int G) { return 64 + (G == 1 ? 9 : 4 * G + 1) * 24; }
};

__device__ __forceinline__ double sgn(int e, int i) { return ((e >> i) & 1) ? -1.0 : 1.0; }

// Branch-free reciprocal / reciprocal square root of a positive, normal double: hardware seed (MUFU.RCP64H / RSQ64H, ~20
// bits) + Newton steps to ~1 ulp.  The CUDA intrinsics add a special-case branch (BSSY/CALL) per use; every argument on
// this path is a clamped slack, a sum of positive weights or a guarded pivot, so none is needed.
__device__ __forceinline__ double rcp_fast(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  x = fma(x, fma(e, e, e), x);        // cubic step: error e^3
  e = fma(-a, x, 1.0);
  return fma(x, e, x);
}
__device__ __forceinline__ double rsqrt_fast(double a) {
  double x;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double h = 0.5 * a, e = fma(-h * x, x, 0.5);      // e = (1 - a x^2) / 2
  x = fma(x, fma(1.5 * e, e, e), x);                // x (1 + e + 3/2 e^2): cubic step
  e = fma(-h * x, x, 0.5);
  return fma(x, e, x);
}

// One hinge row pair of one stage.  The row state comes from L2 / HBM (the L1 left beside 2 x 113 KB of shared memory cannot
// hold it) and the hinge loop has a run-time trip count: the passes fetch HINGE_CHUNK hinges at a time so that their round
// trips overlap (a one-ahead software pipeline measured the same on 8 hinges per stage and 1.6x slower on 255).
template <int D>
struct HingeData {
  double a[D], b, xi, l1, l2;
  bool on;
};

// Row state of one agent in the global workspace: multipliers of the plain rows, the stored slack of the (nonlinear) ball row,
// and (slack, two multipliers) of every hinge pair, plus the PENDING STEP of the hinge pairs.  The step-length pass (S) keeps the
// plain rows' step in registers across its block reduction and updates their multipliers in place once the step length is known;
// a stage's hinge pairs are too many for that (up to hundreds with inter-agent rows), so their step is left in `d*` and the
// next residual pass (R) applies it on the fly.  (A first fused version double-buffered the whole state: 3x the footprint, the
// 1024-agent working set fell out of L2 and the passes stalled on DRAM -- profiles/r02_ipm_kernel_ncu.md.)
struct RowState {
  double *lP;                          // [NPLAIN][K]  multipliers of the plain rows
  double *xi, *l1, *l2;                // [NH][K]      hinge slack and the two multipliers of each hinge pair
  double *sB;                          // [K]          stored slack of the ball row, single integrator only
};
struct AgentPtrs { RowState s; double *dxi, *dl1, *dl2; };

// Store that the compiler does not treat as a possible alias of ordinary loads (no "memory" clobber).  Every use below stores
// to an address that the same thread has already loaded in the same pass (data dependence orders the two) and that no thread
// loads again before the next block barrier; with plain stores every later load of the stage would be ordered behind them.
__device__ __forceinline__ void st_na(double* p, double v) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v)); }

// Bulk prefetch of a contiguous global range into L2 (TMA unit, one instruction per range, no registers or shared memory held):
// address and size must be multiples of 16 bytes.
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
constexpr int HINGE_PREFETCH = 8;   // hinge pairs fetched ahead into L2 by one thread of each hinge group

// ---- per-stage linear forms ----------------------------------------------------------------------
// nu-like combination for interval k: Jn w_{k+1} + Jp w_k + Js g_sigma (- zbar if AFFINE)
template <class Dm, bool AFFINE>
__device__ __forceinline__ void nu_form(const double* jac, const double* wk, const double* wk1, double gsig, double* out) {
  constexpr int NX = Dm::NX, NU = Dm::NU;
  const double* A = jac;
  const double* B = jac + NX * NX;
  const double* C = B + NX * NU;
  const double* S = C + NX * NU;
  const double* Z = S + NX;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double v = wk1[i] - S[i] * gsig;
    if (AFFINE) v -= Z[i];
#pragma unroll
    for (int j = 0; j < NX; ++j) v -= A[j * NX + i] * wk[j];
#pragma unroll
    for (int j = 0; j < NU; ++j) v -= B[j * NX + i] * wk[NX + j] + C[j * NX + i] * wk1[NX + j];
    out[i] = v;
  }
}
// Jp' v (NS) and Jn' v (NS) for a 3-vector v;  Jp = [-A, -B], Jn = [I, -C]
template <class Dm>
__device__ __forceinline__ void JpT(const double* jac, const double* v, double* out) {
  constexpr int NX = Dm::NX, NU = Dm::NU;
  const double* A = jac;
  const double* B = jac + NX * NX;
#pragma unroll
  for (int j = 0; j < NX; ++j) {
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) a -= A[j * NX + i] * v[i];
    out[j] = a;
  }
#pragma unroll
  for (int j = 0; j < NU; ++j) {
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) a -= B[j * NX + i] * v[i];
    out[NX + j] = a;
  }
}
template <class Dm>
__device__ __forceinline__ void JnT(const double* jac, const double* v, double* out) {
  constexpr int NX = Dm::NX, NU = Dm::NU;
  const double* C = jac + NX * NX + NX * NU;
#pragma unroll
  for (int j = 0; j < NX; ++j) out[j] = v[j];
#pragma unroll
  for (int j = 0; j < NU; ++j) {
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) a -= C[j * NX + i] * v[i];
    out[NX + j] = a;
  }
}
// dense Jp / Jn into registers [NX][NS]
template <class Dm>
__device__ __forceinline__ void load_Jp(const double* jac, double (*J)[Dm::NS]) {
  constexpr int NX = Dm::NX, NU = Dm::NU;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
#pragma unroll
    for (int j = 0; j < NX; ++j) J[i][j] = -jac[j * NX + i];
#pragma unroll
    for (int j = 0; j < NU; ++j) J[i][NX + j] = -jac[NX * NX + j * NX + i];
  }
}
template <class Dm>
__device__ __forceinline__ void load_Jn(const double* jac, double (*J)[Dm::NS]) {
  constexpr int NX = Dm::NX, NU = Dm::NU;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
#pragma unroll
    for (int j = 0; j < NX; ++j) J[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < NU; ++j) J[i][NX + j] = -jac[NX * NX + NX * NU + j * NX + i];
  }
}

// Everything the kernel needs about one agent's problem, resident in registers / constant per thread.
struct Scal {
  double sig_ref, r_tr, pos_lo, pos_hi, v_max, w_max;
  double c_sig, c_tnu, cs, qrho, hw_obs, hw_col;
};

// ---- block cyclic reduction ------------------------------------------------------------------------
// Free stages are the nodes p = 1..n (n = K-2) of an SPD block-tridiagonal system.  At stride s = 1, 2, 4, ... the
// nodes p = s, 3s, 5s, ... are eliminated; per eliminated node j we keep
//   slot A: Li_j  (inverse of the Cholesky factor of the pivot block; D_j^-1 v is applied as Li'(Li v): applying the
//                  inverse FACTOR twice keeps the accuracy of triangular solves -- an explicit D^-1 does not, it costs
//                  up to 2x the IPM iterations on ill-conditioned late iterates, see oracle/ipm_struct.py)
//   slot B: P_j = D_j^-1 H[j, j-s]      slot C: Q_j = D_j^-1 H[j, j+s]
// Work items are (node, column) pairs spread over all threads of the block.

// Cholesky + triangular inverse of an NS x NS SPD block given by its lower triangle (registers, fully unrolled).
// A pivot that has lost all significance is frozen (reciprocal 0): that row/column drops out of the step.
template <int NS>
__device__ __forceinline__ void chol_inverse(const double (*Din)[NS], double (*Li)[NS]) {
  double L[NS][NS], dinv[NS];
  // no diagonal regularisation: a relative shift of 1e-13 x trace was measured to floor the dual residual at ~1e-6
  // relative (objective errors up to 2e-6); frozen pivots are the safeguard instead.
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const double d0 = Din[j][j];
    double d = d0;
#pragma unroll
    for (int c = 0; c < j; ++c) d -= L[j][c] * L[j][c];
    const bool okp = d > 1e-12 * d0;
    const double rs = okp ? rsqrt_fast(d) : 0.0;
    dinv[j] = rs;
    L[j][j] = okp ? d * rs : 0.0;
#pragma unroll
    for (int i = j + 1; i < NS; ++i) {
      double v = Din[i][j];
#pragma unroll
      for (int c = 0; c < j; ++c) v -= L[i][c] * L[j][c];
      L[i][j] = v * rs;
    }
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    Li[j][j] = dinv[j];
#pragma unroll
    for (int i = j + 1; i < NS; ++i) {
      double v = 0.0;
#pragma unroll
      for (int c = j; c < i; ++c) v -= L[i][c] * Li[c][j];
      Li[i][j] = v * dinv[i];
    }
  }
}
// y = Li' (Li x)  with Li lower triangular
template <int NS>
__device__ __forceinline__ void apply_dinv(const double (*Li)[NS], const double* x, double* y) {
  double t[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j <= i; ++j) acc += Li[i][j] * x[j];
    t[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = i; j < NS; ++j) acc += Li[j][i] * t[j];
    y[i] = acc;
  }
}

template <class Dm>
__device__ __forceinline__ void cr_factor(double* A, double* B, double* C, int n, int tid, int nthr) {
  constexpr int NS = Dm::NS, SD = Dm::SD;
  for (int s = 1, ls = 0; s <= n; s <<= 1, ++ls) {       // (ls = log2 s: n / s as a shift -- the stride is a loop variable, a division by it
    const int cnt = ((n >> ls) + 1) >> 1;        //  costs an integer-division sequence per level and sweep) odd nodes j = s (2m+1) <= n
    const int npr = nthr / NS;                   // nodes per round: the NS columns of a node always share a round
    // One compute section per round: thread (j, c) derives from shared memory -- untouched until the barrier -- the
    // pivot's inverse factor, column c of P_j and Q_j, the push-left update of D_{j-s} and the push-right update of
    // D_{j+s} with the new left coupling of node j+s.  The stores follow in two sections: the two pushes into an even
    // node come from different threads and must not interleave.
    for (int base = 0; base < cnt; base += npr) {
      const int m = base + tid / NS;
      const bool on = (tid < npr * NS) && (m < cnt);
      int j = 0, c = 0, a = 0, b = 0;
      double Li[NS][NS], pc[NS], qc[NS], upd[NS], a1[NS], a2[NS];
      if (on) {
        c = tid - (tid / NS) * NS;
        j = s * (2 * m + 1); a = j - s; b = j + s;
        const double* dj = A + (size_t)j * SD;
        const double* lj = B + (size_t)j * SD;
        {
          double Dl[NS][NS];
#pragma unroll
          for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int q = 0; q <= i; ++q) Dl[i][q] = dj[i * NS + q];
          chol_inverse<NS>(Dl, Li);
        }
        double x[NS];
        if (a >= 1) {
#pragma unroll
          for (int i = 0; i < NS; ++i) x[i] = lj[i * NS + c];
          apply_dinv<NS>(Li, x, pc);
#pragma unroll
          for (int r = 0; r < NS; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) acc += lj[i * NS + r] * pc[i];
            upd[r] = acc;
          }
        } else {
#pragma unroll
          for (int i = 0; i < NS; ++i) { pc[i] = 0.0; upd[i] = 0.0; }
        }
        if (b <= n) {
          const double* lb = B + (size_t)b * SD;
#pragma unroll
          for (int i = 0; i < NS; ++i) x[i] = lb[c * NS + i];      // column c of H[j, b] = row c of H[b, j]
          apply_dinv<NS>(Li, x, qc);
#pragma unroll
          for (int r = 0; r < NS; ++r) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) { const double l = lb[r * NS + i]; s1 += l * qc[i]; s2 += l * pc[i]; }
            a1[r] = s1; a2[r] = s2;
          }
        } else {
#pragma unroll
          for (int i = 0; i < NS; ++i) { qc[i] = 0.0; a1[i] = 0.0; a2[i] = 0.0; }
        }
      }
      __syncthreads();
      if (on) {
        if (c == 0) {
          double* dj = A + (size_t)j * SD;
#pragma unroll
          for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int q = 0; q < NS; ++q) dj[i * NS + q] = (q <= i) ? Li[i][q] : 0.0;
        }
        double* pj = B + (size_t)j * SD;
        double* qj = C + (size_t)j * SD;
#pragma unroll
        for (int i = 0; i < NS; ++i) { pj[i * NS + c] = pc[i]; qj[i * NS + c] = qc[i]; }
        if (a >= 1) {
          double* da = A + (size_t)a * SD;
#pragma unroll
          for (int r = 0; r < NS; ++r) da[r * NS + c] -= upd[r];
        }
        if (b <= n) {
          double* lb = B + (size_t)b * SD;
#pragma unroll
          for (int r = 0; r < NS; ++r) lb[r * NS + c] = -a2[r];
        }
      }
      __syncthreads();
      if (on && b <= n) {
        double* db = A + (size_t)b * SD;
#pragma unroll
        for (int r = 0; r < NS; ++r) db[r * NS + c] -= a1[r];
      }
      __syncthreads();
    }
  }
}

// rhs columns: c < 4 -> border column c (Rb), c == 4 -> the Newton rhs (dW).  NC = 5: all of them, NC = 1: only dW.
template <class Dm, int NC>
__device__ __forceinline__ double* rhs_col(double* Rb, double* dW, int k, int c) {
  return (NC == 1 || c == 4) ? (dW + (size_t)k * Dm::NSP) : (Rb + (size_t)k * Dm::SR + c * Dm::NS);
}

// NC == 5: work items are (node, column); NC == 1 (the corrector's single right-hand side): items are (node, ROW), so
// that the five rows of a node's update run on five threads instead of in one dependent chain.
template <class Dm, int NC>
__device__ __forceinline__ void cr_forward(const double* B, const double* C, double* Rb, double* dW, int n, int tid, int nthr) {
  constexpr int NS = Dm::NS, SD = Dm::SD;
  for (int s = 1, ls = 0; 2 * s <= n; s <<= 1, ++ls) {
    const int nodes = n >> (ls + 1);               // n / (2 s): even nodes a = 2 s (m+1)
    if (NC == 1) {
      for (int it = tid; it < nodes * NS; it += nthr) {
        const int m = it / NS, r = it - m * NS;
        const int a = 2 * s * (m + 1), jl = a - s, jr = a + s;
        double* va = dW + (size_t)a * Dm::NSP;
        const double* vl = dW + (size_t)jl * Dm::NSP;
        const double* ql = C + (size_t)jl * SD;
        double acc = va[r], acc2 = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) acc -= ql[i * NS + r] * vl[i];
        if (jr <= n) {
          const double* vr = dW + (size_t)jr * Dm::NSP;
          const double* pr = B + (size_t)jr * SD;
#pragma unroll
          for (int i = 0; i < NS; ++i) acc2 += pr[i * NS + r] * vr[i];
        }
        va[r] = acc - acc2;
      }
    } else {
      for (int it = tid; it < nodes * NC; it += nthr) {
        const int m = it / NC, c = it - m * NC;
        const int a = 2 * s * (m + 1), jl = a - s, jr = a + s;
        double* va = rhs_col<Dm, NC>(Rb, dW, a, c);
        const double* vl = rhs_col<Dm, NC>(Rb, dW, jl, c);
        const double* ql = C + (size_t)jl * SD;
        double v[NS];
#pragma unroll
        for (int r = 0; r < NS; ++r) {
          double acc = va[r];
#pragma unroll
          for (int i = 0; i < NS; ++i) acc -= ql[i * NS + r] * vl[i];
          v[r] = acc;
        }
        if (jr <= n) {
          const double* vr = rhs_col<Dm, NC>(Rb, dW, jr, c);
          const double* pr = B + (size_t)jr * SD;
#pragma unroll
          for (int r = 0; r < NS; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) acc += pr[i * NS + r] * vr[i];
            v[r] -= acc;
          }
        }
#pragma unroll
        for (int r = 0; r < NS; ++r) va[r] = v[r];
      }
    }
    __syncthreads();
  }
  if (n < 2) __syncthreads();
}

// u_j = D_j^-1 v_j = Li_j' (Li_j v_j) for every node at once (the forward-eliminated right-hand sides are final), so
// that the backward sweep is a plain x_j = u_j - P_j x_{j-s} - Q_j x_{j+s} without the dependent triangular products.
template <class Dm>
__device__ __forceinline__ void cr_apply_dinv_all(const double* A, double* dW, int n, int tid, int nthr) {
  constexpr int NS = Dm::NS, SD = Dm::SD;
  for (int j = 1 + tid; j <= n; j += nthr) {
    double Li[NS][NS], x[NS], y[NS];
    const double* li = A + (size_t)j * SD;
    double* v = dW + (size_t)j * Dm::NSP;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      x[i] = v[i];
#pragma unroll
      for (int q = 0; q < NS; ++q) Li[i][q] = (q <= i) ? li[i * NS + q] : 0.0;
    }
    apply_dinv<NS>(Li, x, y);
#pragma unroll
    for (int i = 0; i < NS; ++i) v[i] = y[i];
  }
  __syncthreads();
}

// Backward sweep on right-hand sides already multiplied by D^-1 (cr_apply_dinv_all / the Schur-complement pass).
template <class Dm, int NC>
__device__ __forceinline__ void cr_backward(const double* B, const double* C, double* Rb, double* dW, int n, int tid, int nthr) {
  constexpr int NS = Dm::NS, SD = Dm::SD;
  int s = 1, ls = 0;
  while (2 * s <= n) { s <<= 1; ++ls; }
  for (; s >= 1; s >>= 1, --ls) {
    const int nodes = ((n >> ls) + 1) >> 1;        // (n / s + 1) / 2: odd nodes j = s (2m+1)
    if (NC == 1) {
      for (int it = tid; it < nodes * NS; it += nthr) {
        const int m = it / NS, i = it - m * NS;
        const int j = s * (2 * m + 1), jl = j - s, jr = j + s;
        double* vj = dW + (size_t)j * Dm::NSP;
        double acc = vj[i], acc2 = 0.0;
        if (jl >= 1) {
          const double* vl = dW + (size_t)jl * Dm::NSP;
          const double* pj = B + (size_t)j * SD + i * NS;
#pragma unroll
          for (int q = 0; q < NS; ++q) acc -= pj[q] * vl[q];
        }
        if (jr <= n) {
          const double* vr = dW + (size_t)jr * Dm::NSP;
          const double* qj = C + (size_t)j * SD + i * NS;
#pragma unroll
          for (int q = 0; q < NS; ++q) acc2 += qj[q] * vr[q];
        }
        vj[i] = acc - acc2;
      }
    } else {
      for (int it = tid; it < nodes * NC; it += nthr) {
        const int m = it / NC, c = it - m * NC;
        const int j = s * (2 * m + 1), jl = j - s, jr = j + s;
        double* vj = rhs_col<Dm, NC>(Rb, dW, j, c);
        double x[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) x[i] = vj[i];
        if (jl >= 1) {
          const double* vl = rhs_col<Dm, NC>(Rb, dW, jl, c);
          const double* pj = B + (size_t)j * SD;
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < NS; ++q) acc += pj[i * NS + q] * vl[q];
            x[i] -= acc;
          }
        }
        if (jr <= n) {
          const double* vr = rhs_col<Dm, NC>(Rb, dW, jr, c);
          const double* qj = C + (size_t)j * SD;
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < NS; ++q) acc += qj[i * NS + q] * vr[q];
            x[i] -= acc;
          }
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) vj[i] = x[i];
      }
    }
    __syncthreads();
  }
}

// 4x4 Schur complement of the border (sigma, t_nu, t_x, t_u): Gaussian elimination with partial pivoting, redone for
// every right-hand side (the matrix is kept, the cost is nil).  An unpivoted Cholesky produced non-finite steps on
// late, nearly singular iterates where LAPACK's pivoted LU (the CPU twin) sails through.
__device__ __forceinline__ void schur_solve(const double* Sm, const double* rhs, double* out) {
  // fully unrolled, register resident: the pivot search is a chain of compare-and-swap of whole rows (static indices only;
  // a dynamically indexed 4x5 array lives in local memory and costs an L2 round trip per access)
  double A[4][5];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) A[i][j] = Sm[i * 4 + j];
    A[i][4] = rhs[i];
  }
  double inv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int r = c + 1; r < 4; ++r) {
      const bool sw = fabs(A[r][c]) > fabs(A[c][c]);
#pragma unroll
      for (int j = c; j < 5; ++j) {
        const double t = A[c][j];
        A[c][j] = sw ? A[r][j] : t;
        A[r][j] = sw ? t : A[r][j];
      }
    }
    const double d = (fabs(A[c][c]) > 1e-300) ? A[c][c] : 1e-300;
    inv[c] = 1.0 / d;
#pragma unroll
    for (int r = c + 1; r < 4; ++r) {
      const double f = A[r][c] * inv[c];
#pragma unroll
      for (int j = c + 1; j < 5; ++j) A[r][j] -= f * A[c][j];
    }
  }
  double x[4];
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    double v = A[i][4];
#pragma unroll
    for (int j = i + 1; j < 4; ++j) v -= A[i][j] * x[j];
    x[i] = v * inv[i];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = x[i];
}

// coefficients of the three global rows  t_x + t_u +- (sigma - sigma_ref) <= r,  -sigma <= 0  on (sigma, t_nu, t_x, t_u)
__device__ __forceinline__ constexpr double gG(int r, int i) {
  return (r == 0) ? ((i == 1) ? 0.0 : 1.0) : (r == 1) ? ((i == 0) ? -1.0 : (i == 1) ? 0.0 : 1.0) : ((i == 0) ? -1.0 : 0.0);
}

// ---- the kernel -----------------------------------------------------------------------------------
// JSM: the interval Jacobians live in shared memory (compile-time, so that their loads are LDS and not generic LD)
// G: hinge groups.  G = 1: one thread per stage does everything (the shape of the independent-agent batches).  G > 1 (K <= 128):
//    the block has G x 128 threads; thread (g, k) walks the g-th share of stage k's hinge rows in every row pass and the G - 1
//    helper threads of a stage hand their partial sums to the stage's owner through shared memory.  With hundreds of inter-agent
//    rows per stage (config 4: 255) the passes are >= 88 % of the solve and bound by one warp per scheduler's dependent-issue
//    latency; the helper groups are extra warps on the same work (they also take part in the cyclic-reduction work items).
// C: thread-block cluster size.  C > 1 (launches that leave SMs idle: fewer agents than SMs / C): the C blocks of a cluster work
//    on ONE agent.  Block 0 of the cluster (the leader) owns the agent's shared-memory state and runs everything; the other
//    blocks run the same code "hollow" -- their owner-thread index lies beyond every stage / work-item loop, so they execute the
//    same barriers and reductions and nothing else -- except the hinge shares: the C x G hinge groups of the cluster split a
//    stage's rows, remote groups read the stage variables from the leader's shared memory (distributed shared memory) and
//    deposit their partial sums there; block reductions that carry hinge statistics are completed across the cluster.
// RETRY: the retry pass over a compact list (a.retry_list: count, then the ids of the agents whose first solve failed), a few
// blocks that walk the list -- a full-grid retry launch was measured to cost 3.5 % of a step just cycling 128 blocks that return
// at once through block slots of 115 KB each; its main-path instantiation (RETRY = false) carries no loop.
// KT: the number of nodes as a compile-time constant (0 = read a.K): every row-state / shared-memory address is then base + k + a
// constant, which removes most of the integer arithmetic of the row passes (37 % of the executed instructions of the run-time-K
// kernel were integer operations and register moves).  Instantiated for K = 100, the horizon of BASELINE configs 2-4.
// MT: with KT, the number of obstacle rows per stage as a compile-time constant for the PLAIN sub-problem of SCProblem (-1 = run time):
//     no inter-agent rows, no quadratic / linear / game terms, sigma free -- the launch checks that the arguments say so.
template <class M, bool JSM, int G, int C, bool RETRY, int KT = 0, int MT = -1>
__global__ void __launch_bounds__(G == 1 ? SOLVER_MAX_THREADS : 128 * G, 1)
ipm_kernel(scvx_solve_args a, double mu0_default, double eps_gap, double eps_feas, size_t jac_ws_offset, double om_floor, double sf_cap) {
  using Dm = Dims<M>;
  constexpr int NX = Dm::NX, NU = Dm::NU, D = Dm::D, NS = Dm::NS, NEX = Dm::NEX, NEU = Dm::NEU;
  constexpr int NSP = Dm::NSP, SD = Dm::SD, SR = Dm::SR, NJ = Dm::NJ, STG = Dm::STG, NPLAIN = Dm::NPLAIN;
  constexpr bool BALL = Dm::BALL;
  namespace cg = cooperative_groups;
  const int rank = (C > 1) ? (int)cg::this_cluster().block_rank() : 0;
  const bool leader = (rank == 0);
  int unit = (int)blockIdx.x / C;                             // one agent per cluster
next_unit:
  if (RETRY && unit >= a.retry_list[0]) return;               // uniform over the block (and the cluster)
  const int K = KT > 0 ? KT : a.K, agent = RETRY ? a.retry_list[1 + unit] : (a.block_order ? a.block_order[unit] : unit);
  const int nthr = (KT > 0 && G == 1) ? ((KT + 31) / 32) * 32 : (int)blockDim.x;      // (the launch uses exactly this block size)
  const int htid = threadIdx.x;                                // index of this thread among its block's hinge workers
  const int tid = leader ? htid : htid + (1 << 20);            // owner-thread index: out of every loop's range on a helper block
  auto CSYNC = [&]() { if (C > 1) cg::this_cluster().sync(); };
  if (a.active && !a.active[agent]) {        // the agent's outer loop has converged: nothing to solve (uniform over the cluster)
    if (threadIdx.x == 0 && leader && !a.retry_failed) a.iters[agent] = 0;
    return;
  }
  if (!RETRY && a.retry_failed && a.status[agent] == SCVX_ST_OPTIMAL) return;      // full-grid retry pass: only the failed agents
  const int Mobs = (MT >= 0) ? MT : a.M, NH = (MT >= 0) ? MT : a.M + a.n_nbr;
  // start value of the barrier parameter: per agent when the caller provides one (scvx_mu0_from_iters), else the default
  const double mu0_agent = (a.mu0 && C == 1) ? a.mu0[agent] : 0.0;      // (cluster launches always start cold)
  double mu0 = (mu0_agent > 0.0) ? mu0_agent : mu0_default;

  extern __shared__ __align__(16) double smem[];
  PHASE_INIT();
  double* W = smem;                     // [K][NSP] current stage variables
  double* dWa = W + (size_t)K * NSP;    // [K][NSP] affine (predictor) direction
  double* dW = dWa + (size_t)K * NSP;   // [K][NSP] rhs / solution of the current solve
  double* Dk = dW + (size_t)K * NSP;    // [K][SD]  slot A: diagonal blocks -> Li (inverse Cholesky factor of the pivot block)
  double* Ek = Dk + (size_t)K * SD;     // [K][SD]  slot B: coupling to the current left neighbour H[k, k-s] -> P_k
  double* Ck = Ek + (size_t)K * SD;     // [K][SD]  slot C: Q_k (aliased by the assembly staging ST before the factorisation)
  double* Rb = Ck + (size_t)K * SD;     // [K][SR]  border columns Bd -> Y = T^-1 Bd     layout [c*NS + i]
  double* ST2 = Rb + (size_t)K * SR;    // [K][3]   corrector staging
  double* gl = ST2 + (size_t)K * Dm::ST2;   // [64] globals
  double* red = gl + 64;                // [9][24] reduction scratch
  double* ST = Ck;                      // per-interval staging of the assembly pass (16 of SD doubles)
  // interval Jacobians [K][NJ]: shared memory when it fits, else a slice of the global workspace
  constexpr int RED = (G == 1 ? 9 : 4 * G + 1) * 24;
  double* JAC = JSM ? (red + RED) : ((double*)a.workspace + jac_ws_offset + (size_t)agent * K * NJ);
  // partial results of the helper hinge groups: slots 0 .. G-2 for the groups of the own block, G-1 .. G+C-3 (leader only) for the
  // combined partials of the other blocks of the cluster; XR: the other blocks' block-reduction results
  constexpr int NSLOT = (G - 1) + (C - 1);
  double* HS = red + RED + (JSM ? (size_t)K * NJ : 0);       // [NSLOT][K][HSW]
  double* XR = HS + (size_t)NSLOT * K * Dm::HSW;              // [C-1][24]
  constexpr int HSW = Dm::HSW;
  // what a helper block reads / writes in the leader's shared memory
  const double* Wl = W; const double* dWal = dWa; const double* dWl = dW; const double* gll = gl;
  double* HSl = HS; double* XRl = XR;
  if (C > 1 && !leader) {
    auto cl = cg::this_cluster();
    Wl = cl.map_shared_rank(W, 0); dWal = cl.map_shared_rank(dWa, 0); dWl = cl.map_shared_rank(dW, 0); gll = cl.map_shared_rank(gl, 0);
    HSl = cl.map_shared_rank(HS, 0); XRl = cl.map_shared_rank(XR, 0);
  }
  // hinge share of this thread: group wgrp of C x G walks hinges [h_lo, h_hi) of stage kt
  const int grp = (G == 1) ? 0 : htid / 128;
  const int kt = (G == 1) ? htid : htid - grp * 128;
  constexpr int Q = C * G;
  const int wgrp = rank * G + grp;
  const int h_per = (NH + Q - 1) / Q;
  const int h_lo = (Q == 1) ? 0 : min(NH, wgrp * h_per), h_hi = (Q == 1) ? NH : min(NH, h_lo + h_per);
  const bool helper = (Q > 1) && wgrp > 0;                   // every hinge group but the stage owner's own
  // block reduction completed across the cluster (reductions that carry statistics of the hinge rows)
  auto cluster_reduce = [&](auto nv_c, double* vals, const int* ops) {
    constexpr int NV = decltype(nv_c)::value;
    block_reduce<NV>(vals, ops, red);
    if (C > 1) {
      if (!leader && htid < NV) XRl[(rank - 1) * 24 + htid] = red[htid];
      cg::this_cluster().sync();
      if (leader && htid < NV) {
        double v = red[htid];
        for (int r = 1; r < C; ++r) {
          const double o = XR[(r - 1) * 24 + htid];
          v = (ops[htid] == 0) ? v + o : (ops[htid] == 1 ? fmin(v, o) : fmax(v, o));
        }
        red[htid] = v;
      }
      __syncthreads();
    }
  };
  // helper blocks: combine the own groups' partials (N doubles per stage) and deposit them in the leader's slot of this block
  auto deposit_remote = [&](auto n_c, double* mine) {
    constexpr int N = decltype(n_c)::value;
    if (C > 1 && !leader) {
      __syncthreads();
      if (grp == 0 && kt > 0 && kt < K - 1) {
        for (int g = 1; g < G; ++g) {
          const double* hs = HS + ((size_t)(g - 1) * K + kt) * HSW;
#pragma unroll
          for (int i = 0; i < N; ++i) mine[i] += hs[i];
        }
        double* out = HSl + ((size_t)(G - 1 + rank - 1) * K + kt) * HSW;
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = mine[i];
      }
    }
  };
  // globals: gl[0..3] = sigma, t_nu, t_x, t_u ; gl[4..7] = dg_aff ; gl[8..11] = dg ; gl[12..14] sG ; gl[15..17] lG ;
  // gl[18..33] = Gg/S 4x4 ; gl[34..37] = bg ; gl[40] flag ; gl[41] alpha_p ; gl[42] alpha_d ; gl[43] sigmu ; gl[44] mu
  // gl[45] comp ; gl[46..49] Y'b scratch

  // ---- per-agent scalars & pointers --------------------------------------------------------------
  Scal sc;
  sc.sig_ref = a.sigma_ref[agent]; sc.r_tr = a.tr_radius[agent];
  sc.pos_lo = a.pos_lo[agent]; sc.pos_hi = a.pos_hi[agent];
  sc.v_max = a.v_max[agent]; sc.w_max = a.w_max ? a.w_max[agent] : 0.0;
  sc.cs = 1.0 / fmax(fmax(a.weight_nu, a.weight_sigma), 1e-300);
  sc.c_sig = a.weight_sigma * sc.cs; sc.c_tnu = a.weight_nu * sc.cs;
  constexpr bool PLAIN = (MT >= 0);
  sc.qrho = (!PLAIN && a.quad_rho) ? a.quad_rho[agent] * sc.cs : 0.0;
  sc.hw_obs = a.weight_slack * sc.cs; sc.hw_col = a.weight_col * sc.cs;
  const double* qlin = (!PLAIN && a.lin_p) ? a.lin_p + (size_t)agent * D * K : nullptr;
  // best-response terms (all absent for the plain problems): diagonal / linear / consecutive-difference quadratics
  const double* qdiag = (!PLAIN && a.quad_diag) ? a.quad_diag + (size_t)agent * NS : nullptr;
  const double* qpair = (!PLAIN && a.quad_pair) ? a.quad_pair + (size_t)agent * NS : nullptr;
  const double* qlinw = (!PLAIN && a.lin_w) ? a.lin_w + (size_t)agent * NS * K : nullptr;
  const bool game = qdiag || qpair || qlinw;
  const bool fix_sig = !PLAIN && a.fix_sigma != 0;
  // gradient and curvature of the smooth extra cost at stage k, component i (scaled); obj: this stage's share of the value
  auto game_terms = [&](int k, int i, const double* w, double& grad, double& curv, double& obj) {
    grad = 0.0; curv = 0.0; obj = 0.0;
    if (qdiag) { const double q = qdiag[i] * sc.cs; grad += q * w[i]; curv += q; obj += 0.5 * q * w[i] * w[i]; }
    if (qlinw) { const double l = qlinw[(size_t)i * K + k] * sc.cs; grad += l; obj += l * w[i]; }
    if (qpair) {
      const double c = qpair[i] * sc.cs;
      if (k > 0) { const double dlt = w[i] - w[i - NSP]; grad += c * dlt; curv += c; }
      if (k < K - 1) { const double dlt = w[i + NSP] - w[i]; grad -= c * dlt; curv += c; obj += 0.5 * c * dlt * dlt; }
    }
  };
  const double* obs_a = a.obs_a ? a.obs_a + (size_t)agent * Mobs * D * K : nullptr;
  const double* obs_b = a.obs_b ? a.obs_b + (size_t)agent * Mobs * K : nullptr;
  const double* col_a = (!PLAIN && a.col_a) ? a.col_a + (size_t)agent * a.n_nbr * D * K : nullptr;
  const double* col_b = (!PLAIN && a.col_b) ? a.col_b + (size_t)agent * a.n_nbr * K : nullptr;
  const unsigned char* col_mask = (!PLAIN && a.col_mask) ? a.col_mask + (size_t)agent * a.n_nbr : nullptr;
  const bool coupled = (sc.qrho > 0.0) || BALL || game;  // common primal/dual step length

  AgentPtrs ws;
  {
    const size_t per = (size_t)K * (NPLAIN + 6 * (size_t)NH + (BALL ? 1 : 0));
    double* base = (double*)a.workspace + (size_t)agent * per;
    ws.s.lP = base; ws.s.xi = ws.s.lP + (size_t)NPLAIN * K; ws.s.l1 = ws.s.xi + (size_t)NH * K; ws.s.l2 = ws.s.l1 + (size_t)NH * K;
    ws.s.sB = ws.s.l2 + (size_t)NH * K;
    ws.dxi = ws.s.sB + (BALL ? K : 0); ws.dl1 = ws.dxi + (size_t)NH * K; ws.dl2 = ws.dl1 + (size_t)NH * K;
    // no step is pending before the first iteration: zero deltas (and step lengths, below)
    for (size_t i = tid; i < 3 * (size_t)NH * K; i += nthr) ws.dxi[i] = 0.0;
  }
  auto hinge_a = [&](int h, int c, int k) -> double {
    return (h < Mobs) ? obs_a[((size_t)h * D + c) * K + k] : col_a[((size_t)(h - Mobs) * D + c) * K + k];
  };
  auto hinge_b = [&](int h, int k) -> double {
    return (h < Mobs) ? obs_b[(size_t)h * K + k] : col_b[(size_t)(h - Mobs) * K + k];
  };
  auto hinge_on = [&](int h) -> bool { return (h < Mobs) || !col_mask || col_mask[h - Mobs]; };
  auto hinge_w = [&](int h) -> double { return (h < Mobs) ? sc.hw_obs : sc.hw_col; };
  auto load_hinge_ab = [&](int h, int k) -> HingeData<D> {
    HingeData<D> r;
    r.on = (h < h_hi) && hinge_on(h);
    r.b = 0.0; r.xi = 1.0; r.l1 = 1.0; r.l2 = 1.0;
#pragma unroll
    for (int c = 0; c < D; ++c) r.a[c] = 0.0;
    if (r.on) {
#pragma unroll
      for (int c = 0; c < D; ++c) r.a[c] = hinge_a(h, c, k);
      r.b = hinge_b(h, k);
    }
    return r;
  };
  auto load_hinge = [&](int h, int k, const RowState& src) -> HingeData<D> {
    HingeData<D> r = load_hinge_ab(h, k);
    if (r.on) {
      const size_t o = (size_t)h * K + k;
      r.xi = src.xi[o]; r.l1 = src.l1[o]; r.l2 = src.l2[o];
    }
    return r;
  };
  // The walk over a stage share's hinge rows is a linear sweep over [h][c][k] tables far larger than L2 (config 4: 1.4-2.6 MB
  // per agent and pass, 366 MB for 256 agents), latency-bound on DRAM with eight warps per SM (long scoreboard 3.0 warps per
  // issue): one thread of every hinge group asks the TMA unit to pull the rows HINGE_PREFETCH pairs ahead into L2 while the
  // group works on the current chunk.  Only the inter-agent rows (h >= Mobs; the obstacle rows are few) and only when the
  // ranges are 16-byte aligned (K even).
  const bool can_prefetch = (K % 2 == 0) && col_a && (kt == 1);       // stage 1: the first free stage of the group
  auto prefetch_hinges = [&](int h0, bool with_step) {
    if (!can_prefetch) return;
    const int ha = max(h0 + HINGE_PREFETCH, Mobs), hb = min(ha + HINGE_CHUNK, h_hi);
    if (hb <= ha) return;
    const unsigned n = (unsigned)(hb - ha), rowb = (unsigned)K * 8u;
    l2_prefetch(col_a + (size_t)(ha - Mobs) * D * K, n * D * rowb);
    l2_prefetch(col_b + (size_t)(ha - Mobs) * K, n * rowb);
    l2_prefetch(ws.s.xi + (size_t)ha * K, n * rowb); l2_prefetch(ws.s.l1 + (size_t)ha * K, n * rowb); l2_prefetch(ws.s.l2 + (size_t)ha * K, n * rowb);
    if (with_step) {
      l2_prefetch(ws.dxi + (size_t)ha * K, n * rowb); l2_prefetch(ws.dl1 + (size_t)ha * K, n * rowb); l2_prefetch(ws.dl2 + (size_t)ha * K, n * rowb);
    }
  };
  // residual pass: the pending step of the hinge pair is applied, in place.  Loads and stores are separate calls so that a chunk
  // issues all of its loads before its first store (ptxas orders a load behind an earlier store it cannot prove disjoint).
  auto load_hinge_apply = [&](int h, int k, double al_p, double al_d) -> HingeData<D> {
    HingeData<D> r = load_hinge_ab(h, k);
    if (r.on) {
      const size_t o = (size_t)h * K + k;
      r.xi = fma(al_p, ws.dxi[o], ws.s.xi[o]); r.l1 = fma(al_d, ws.dl1[o], ws.s.l1[o]); r.l2 = fma(al_d, ws.dl2[o], ws.s.l2[o]);
      if (MT < 0) { st_na(ws.s.xi + o, r.xi); st_na(ws.s.l1 + o, r.l1); st_na(ws.s.l2 + o, r.l2); }      // (see hinge_R_range)
    }
    return r;
  };
  auto store_hinge_state = [&](int h, int k, const HingeData<D>& r) {
    if (r.on) {
      const size_t o = (size_t)h * K + k;
      st_na(ws.s.xi + o, r.xi); st_na(ws.s.l1 + o, r.l1); st_na(ws.s.l2 + o, r.l2);
    }
  };

  // ---- load problem data into shared memory ---------------------------------------------------------
  const double* Xr = a.X_ref + (size_t)agent * NX * K;
  const double* Ur = a.U_ref + (size_t)agent * NU * K;
  // trust-region centre, read from global memory (coalesced over k; L1/L2 resident)
  auto WR = [&](int k, int i) -> double { return (i < NX) ? Xr[(size_t)i * K + k] : Ur[(size_t)(i - NX) * K + k]; };
  {
    const int Km1 = K - 1;
    const double* Ab = a.A_bar + (size_t)agent * NX * NX * Km1;
    const double* Bb = a.B_bar + (size_t)agent * NX * NU * Km1;
    const double* Cb = a.C_bar + (size_t)agent * NX * NU * Km1;
    const double* Sb = a.S_bar + (size_t)agent * NX * Km1;
    const double* Zb = a.z_bar + (size_t)agent * NX * Km1;
    for (int k = tid; k < Km1; k += nthr) {
      double* j = JAC + (size_t)k * NJ;
#pragma unroll
      for (int r = 0; r < NX * NX; ++r) j[r] = Ab[(size_t)r * Km1 + k];
#pragma unroll
      for (int r = 0; r < NX * NU; ++r) { j[NX * NX + r] = Bb[(size_t)r * Km1 + k]; j[NX * NX + NX * NU + r] = Cb[(size_t)r * Km1 + k]; }
#pragma unroll
      for (int r = 0; r < NX; ++r) { j[NX * NX + 2 * NX * NU + r] = Sb[(size_t)r * Km1 + k]; j[NX * NX + 2 * NX * NU + NX + r] = Zb[(size_t)r * Km1 + k]; }
    }
  }
  __syncthreads();

  // ---- initial point (strictly inside the boxes; t's large enough; central complementarity) ----------
  {
    const double dlt = fmin(1e-2 * (sc.pos_hi - sc.pos_lo), sc.r_tr / (16.0 * NS));
    const double dv = fmin(1e-2 * sc.v_max, sc.r_tr / (16.0 * NS));
    const double dw = fmin(2e-2 * sc.w_max, sc.r_tr / (16.0 * NS));
    for (int k = tid; k < K; k += nthr) {
      double* w = W + k * NSP;
      if (k == 0 || k == K - 1) {
        const double* xb = (k == 0 ? a.x_init : a.x_final) + (size_t)agent * NX;
#pragma unroll
        for (int i = 0; i < NX; ++i) w[i] = xb[i];
#pragma unroll
        for (int j = 0; j < NU; ++j) w[NX + j] = 0.0;
      } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) w[i] = WR(k, i);
#pragma unroll
        for (int i = 0; i < D; ++i) w[i] = fmin(fmax(w[i], sc.pos_lo + dlt), sc.pos_hi - dlt);
        if (!BALL) {
          w[NX] = fmin(fmax(w[NX], dv), sc.v_max - dv);
          w[NX + 1] = fmin(fmax(w[NX + 1], -sc.w_max + dw), sc.w_max - dw);
        } else {
          double n2 = 0.0;
#pragma unroll
          for (int j = 0; j < NU; ++j) n2 += w[NX + j] * w[NX + j];
          const double s = fmin(1.0, 0.99 * sc.v_max / fmax(sqrt(n2), 1e-300));
#pragma unroll
          for (int j = 0; j < NU; ++j) w[NX + j] *= s;
        }
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) { dWa[k * NSP + i] = 0.0; dW[k * NSP + i] = 0.0; }
    }
  }
  __syncthreads();
  // A start below the default is for WARM problems only.  Where the start point violates a hinge row -- the trajectory runs through
  // an obstacle of the new linearisation -- that pair's closed-form start carries a multiplier of the full penalty weight which no
  // small plain-row multiplier balances: the dual residual starts ~1e5 x mu0 and the iteration strands on short dual steps (twin:
  // 40-59 iterations against 15-27 from the cold start on the seven stranded sub-problems of tools/dump_stranded.py).  Such
  // problems start cold.
  if (C == 1 && mu0 < mu0_default) {
    double vm[1] = {-1e300};
    if (kt > 0 && kt < K - 1) {
      double w[D];
#pragma unroll
      for (int c = 0; c < D; ++c) w[c] = W[kt * NSP + c];
      for (int h = h_lo; h < h_hi; ++h) {
        if (!hinge_on(h)) continue;
        double ap = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) ap += hinge_a(h, c, kt) * w[c];
        vm[0] = fmax(vm[0], hinge_b(h, kt) - ap);
      }
    }
    const int opv[1] = {2};
    block_reduce<1>(vm, opv, red);
    if (red[0] > 1e-6) mu0 = mu0_default;
  }
  {
    // t_nu, t_x, t_u from maxima over stages
    const double sig0 = fix_sig ? sc.sig_ref : fmax(sc.sig_ref, fmin(1e-2, sc.r_tr / 16.0));
    double v[3] = {0.0, 0.0, 0.0};
    for (int k = tid; k < K; k += nthr) {
      const double* w = W + k * NSP;
      if (k < K - 1) {
        double nu[NX];
        nu_form<Dm, true>(JAC + (size_t)k * NJ, w, w + NSP, sig0, nu);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) s += fabs(nu[i]);
        v[0] = fmax(v[0], s);
      }
      double sx = 0.0, su = 0.0;
#pragma unroll
      for (int i = 0; i < NX; ++i) sx += fabs(w[i] - WR(k, i));
#pragma unroll
      for (int j = 0; j < NU; ++j) su += fabs(w[NX + j] - WR(k, NX + j));
      v[1] = fmax(v[1], sx); v[2] = fmax(v[2], su);
    }
    const int ops[3] = {2, 2, 2};
    block_reduce<3>(v, ops, red);
    if (tid == 0) {
      gl[0] = sig0; gl[1] = red[0] * 1.1 + 1.0; gl[2] = red[1] + sc.r_tr * 0.25; gl[3] = red[2] + sc.r_tr * 0.25;
      const double sG[3] = {sc.r_tr - gl[2] - gl[3] - (sig0 - sc.sig_ref), sc.r_tr - gl[2] - gl[3] + (sig0 - sc.sig_ref), sig0};
      for (int r = 0; r < 3; ++r) { gl[12 + r] = fmax(sG[r], 1e-8); gl[15 + r] = mu0 / gl[12 + r]; }
      gl[41] = 0.0; gl[42] = 0.0;            // no step pending for the first residual pass
    }
    __syncthreads();
  }
  // row state init
  for (int k = tid; k < K; k += nthr) {
    const double* w = W + k * NSP;
    const double sig = gl[0], tnu = gl[1], tx = gl[2], tu = gl[3];
    const bool fr = (k > 0 && k < K - 1);
    if (k < K - 1) {
      double nu[NX];
      nu_form<Dm, true>(JAC + (size_t)k * NJ, w, w + NSP, sig, nu);
#pragma unroll
      for (int e = 0; e < NEX; ++e) {
        double f = -tnu;
#pragma unroll
        for (int i = 0; i < NX; ++i) f += sgn(e, i) * nu[i];
        const double s = fmax(-f, 1e-8);
        ws.s.lP[(size_t)(Dm::R_NU + e) * K + k] = mu0 / s;
      }
    }
#pragma unroll
    for (int e = 0; e < NEX; ++e) {
      double f = -tx;
#pragma unroll
      for (int i = 0; i < NX; ++i) f += sgn(e, i) * (w[i] - WR(k, i));
      const double s = fmax(-f, 1e-8);
      ws.s.lP[(size_t)(Dm::R_X + e) * K + k] = mu0 / s;
    }
#pragma unroll
    for (int e = 0; e < NEU; ++e) {
      double f = -tu;
#pragma unroll
      for (int j = 0; j < NU; ++j) f += sgn(e, j) * (w[NX + j] - WR(k, NX + j));
      const double s = fmax(-f, 1e-8);
      ws.s.lP[(size_t)(Dm::R_U + e) * K + k] = mu0 / s;
    }
    if (fr) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double s = fmax(sc.pos_hi - w[i], 1e-8);
        ws.s.lP[(size_t)(Dm::R_P + i) * K + k] = mu0 / s;
        s = fmax(w[i] - sc.pos_lo, 1e-8);
        ws.s.lP[(size_t)(Dm::R_P + D + i) * K + k] = mu0 / s;
      }
      if (!BALL) {
        const double sv[4] = {sc.v_max - w[NX], w[NX], sc.w_max - w[NX + 1], sc.w_max + w[NX + 1]};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double s = fmax(sv[r], 1e-8);
          ws.s.lP[(size_t)(Dm::R_V + r) * K + k] = mu0 / s;
        }
      } else {
        double n2 = 0.0;
#pragma unroll
        for (int j = 0; j < NU; ++j) n2 += w[NX + j] * w[NX + j];
        const double s = fmax(0.5 * (sc.v_max * sc.v_max - n2), 1e-8);
        ws.s.sB[k] = s; ws.s.lP[(size_t)Dm::R_V * K + k] = mu0 / s;
      }
    }
  }
  // hinge pairs: dual-feasible central start in closed form; every group initialises its own share
  CSYNC();                                  // the leader's stage variables are in place
  if (kt > 0 && kt < K - 1) {
    const int k = kt;
    double w[D];
#pragma unroll
    for (int c = 0; c < D; ++c) w[c] = Wl[k * NSP + c];
    for (int h = h_lo; h < h_hi; ++h) {
      if (!hinge_on(h)) continue;
      double ap = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) ap += hinge_a(h, c, k) * w[c];
      const double viol = hinge_b(h, k) - ap, hw = hinge_w(h);
      const double disc = sqrt(hw * viol * hw * viol + 4.0 * mu0 * mu0);
      const double num = (viol >= 0.0) ? hw * viol + disc : 4.0 * mu0 * mu0 / fmax(disc - hw * viol, 1e-300);
      const double xi = (num + 2.0 * mu0) / (2.0 * hw);
      const size_t o = (size_t)h * K + k;
      ws.s.xi[o] = xi; ws.s.l1[o] = mu0 / (xi - viol); ws.s.l2[o] = mu0 / xi;
    }
  }
  __syncthreads();

  // number of complementarity pairs
  int n_active_h = 0;
  for (int h = 0; h < NH; ++h) n_active_h += hinge_on(h) ? 1 : 0;
  const double n_rows = (double)(K - 1) * NEX + (double)K * NEX + (double)K * NEU + 3.0 +
                        (double)(K - 2) * (2 * D + Dm::NV) + 2.0 * n_active_h * (double)(K - 2);

  PHASE(0);
  const double inv_n_rows = 1.0 / n_rows;
  int status = SCVX_ST_MAXITER, it = 0;
  // a start below the default that has not converged after SMALL_START_CAP iterations gives up (the caller's retry pass solves it cold)
  constexpr int SMALL_START_CAP = 40;
  const int max_iter = (mu0 < mu0_default) ? min(a.max_iter > 0 ? a.max_iter : 120, SMALL_START_CAP) : (a.max_iter > 0 ? a.max_iter : 120);

  // =================================================================================================
  // Three row passes per iteration.  R: the pending step is applied to the row state, then residuals + Newton matrix staging +
  //                                  predictor rhs staging + stationarity staging
  //                               P: affine step statistics (alpha_p, alpha_d, three sums for mu_aff) + corrector rhs staging
  //                               S: final step lengths; the row-state step is left for the next R pass
  // Per-stage results that other threads need go to ST (interval part) and dW / Dk / Rb (own stage).
  // =================================================================================================
  for (it = 0; it < max_iter; ++it) {
    double part[24];
    // ------------------------------------------------------------------------------------ RESIDUAL PASS (+ pending step)
#pragma unroll
    for (int i = 0; i < 24; ++i) part[i] = 0.0;
    // part: 0 comp, 1 rp_inf, 2 rd_inf(xi), 3 Gg00, 4 Gg01, 5 Gg11, 6 Gg22, 7 Gg33, 8..11 bg(tau), 12..15 rdg(lambda), 16 obj_hinge, 17 obj_quad
    const double sig = gl[0], tnu = gl[1], tx = gl[2], tu = gl[3];
    // The step of the PREVIOUS iteration is applied to the hinge pairs here, on the fly (the plain rows were updated by pass S).
    CSYNC();                                // stage variables and step lengths of the previous iteration are final
    const double al_p0 = gll[41], al_d0 = gll[42];
    auto LPA = [&](size_t o) -> double { return ws.s.lP[o]; };
    // One hinge pair of the residual pass: complementarity / stationarity statistics and the pair's contribution to the stage's
    // Newton block (Dacc[c][c2], c, c2 < D), predictor rhs (bt_) and stationarity accumulator (bl_).
    auto hinge_R = [&](int h, const HingeData<D>& hd, const double* w, auto& Dacc, double* bt_, double* bl_) {
      double av[D], ap = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) { av[c] = hd.a[c]; ap += av[c] * w[c]; }
      const double xi = hd.xi, l1 = hd.l1, l2 = hd.l2, hw = hinge_w(h);
      const double viol = hd.b - ap;
      const double s1 = fmax(xi - viol, TINY_S), s2 = fmax(xi, TINY_S);
      const double w1 = l1 * rcp_fast(s1), w2 = l2 * rcp_fast(s2), rw = rcp_fast(w1 + w2), weff = w1 * w2 * rw;
      const double th = w1 * hw * rw;                 // both primal residuals are zero by construction: rhs_xi = -hw
      part[0] += s1 * l1 + s2 * l2;
      part[2] = fmax(part[2], fabs(hw - l1 - l2));
      part[16] += hw * xi;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        bt_[c] -= av[c] * th; bl_[c] -= av[c] * l1;
#pragma unroll
        for (int c2 = 0; c2 < D; ++c2) Dacc[c][c2] += weff * av[c] * av[c2];
      }
    };
    // hinge rows in chunks of HINGE_CHUNK: all loads of a chunk are issued before its first use (memory-level parallelism is
    // what a latency-bound walk over hundreds of neighbour rows needs)
    auto hinge_R_range = [&](int k, const double* w, auto& Dacc, double* bt_, double* bl_) {
      for (int h0 = h_lo; h0 < h_hi; h0 += HINGE_CHUNK) {
        prefetch_hinges(h0, true);
        HingeData<D> hb[HINGE_CHUNK];
#pragma unroll
        for (int c = 0; c < HINGE_CHUNK; ++c) hb[c] = load_hinge_apply(h0 + c, k, al_p0, al_d0);
        // few hinge rows (the compile-time-M kernels): all stores of the chunk after all of its loads (-5 % on the bench step); with
        // hundreds of rows per stage (config 4) a pair's stores right behind its loads are better (same-box A/B: 41.3 vs 44.4 ms per round)
        if (MT >= 0) {
#pragma unroll
          for (int c = 0; c < HINGE_CHUNK; ++c) store_hinge_state(h0 + c, k, hb[c]);
        }
#pragma unroll
        for (int c = 0; c < HINGE_CHUNK; ++c)
          if (hb[c].on) hinge_R(h0 + c, hb[c], w, Dacc, bt_, bl_);
      }
    };
    {
      // helper groups: their share of stage kt's hinge rows; partial sums go to the stage's owner through (distributed) shared memory
      double ph[D * D + 2 * D];
#pragma unroll
      for (int i = 0; i < D * D + 2 * D; ++i) ph[i] = 0.0;
      if (helper && kt > 0 && kt < K - 1) {
        double Dh[D][D], bth[D], blh[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
          bth[c] = 0.0; blh[c] = 0.0;
#pragma unroll
          for (int c2 = 0; c2 < D; ++c2) Dh[c][c2] = 0.0;
        }
        double wl[D];                       // the stage's position, once (for a helper block it lives in the leader's shared memory)
#pragma unroll
        for (int c = 0; c < D; ++c) wl[c] = Wl[kt * NSP + c];
        hinge_R_range(kt, wl, Dh, bth, blh);
#pragma unroll
        for (int c = 0; c < D; ++c) {
          ph[D * D + c] = bth[c]; ph[D * D + D + c] = blh[c];
#pragma unroll
          for (int c2 = 0; c2 < D; ++c2) ph[c * D + c2] = Dh[c][c2];
        }
        if (grp > 0) {
          double* hs = HS + ((size_t)(grp - 1) * K + kt) * HSW;
#pragma unroll
          for (int i = 0; i < D * D + 2 * D; ++i) hs[i] = ph[i];
        }
      }
      deposit_remote(std::integral_constant<int, D * D + 2 * D>{}, ph);
    }
    for (int k = tid; k < K; k += nthr) {
      const double* w = W + k * NSP;
      const bool fr = (k > 0 && k < K - 1);
      double* st = ST + (size_t)k * STG;
      // ---- nu rows of interval k
      if (k < K - 1) {
        const double* jac = JAC + (size_t)k * NJ;
        double nu[NX];
        nu_form<Dm, true>(jac, w, w + NSP, sig, nu);
        double Mm[6] = {0, 0, 0, 0, 0, 0}, mv[NX] = {0, 0, 0}, sw = 0.0, et[NX] = {0, 0, 0}, stau = 0.0, el[NX] = {0, 0, 0}, slam = 0.0;
#pragma unroll
        for (int e = 0; e < NEX; ++e) {
          const size_t o = (size_t)(Dm::R_NU + e) * K + k;
          const double l = LPA(o);
          double f = -tnu;
#pragma unroll
          for (int i = 0; i < NX; ++i) f += sgn(e, i) * nu[i];
          const double s = fmax(-f, TINY_S), rp = 0.0, wgt = l * rcp_fast(s), tau = 0.0;
          part[0] += s * l; part[1] = fmax(part[1], fabs(rp));
          sw += wgt; stau += tau; slam += l;
#pragma unroll
          for (int i = 0; i < NX; ++i) { mv[i] += wgt * sgn(e, i); et[i] += tau * sgn(e, i); el[i] += l * sgn(e, i); }
          Mm[0] += wgt; Mm[1] += wgt * sgn(e, 0) * sgn(e, 1); Mm[2] += wgt * sgn(e, 0) * sgn(e, 2);
          Mm[3] += wgt; Mm[4] += wgt * sgn(e, 1) * sgn(e, 2); Mm[5] += wgt;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) st[i] = Mm[i];
#pragma unroll
        for (int i = 0; i < NX; ++i) { st[6 + i] = mv[i]; st[10 + i] = et[i]; st[13 + i] = el[i]; }
        st[9] = sw;
        // global-column contributions of this interval
        const double* S = jac + NX * NX + 2 * NX * NU;
        double Js[NX], MJs[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) Js[i] = -S[i];
        MJs[0] = Mm[0] * Js[0] + Mm[1] * Js[1] + Mm[2] * Js[2];
        MJs[1] = Mm[1] * Js[0] + Mm[3] * Js[1] + Mm[4] * Js[2];
        MJs[2] = Mm[2] * Js[0] + Mm[4] * Js[1] + Mm[5] * Js[2];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          part[3] += Js[i] * MJs[i]; part[4] -= mv[i] * Js[i];
          part[8] += et[i] * Js[i]; part[12] += el[i] * Js[i];
        }
        part[5] += sw; part[9] -= stau; part[13] -= slam;
      }
      // ---- own-stage accumulators: diag block (sym, full storage), border cols, rhs (tau), stationarity (lambda)
      double Dl[NS][NS], bt[NS], bl[NS], bx[NX], bu[NU];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        bt[i] = 0.0; bl[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) Dl[i][j] = 0.0;
      }
      // x-trust rows
      {
        double dx[NX], mvx[NX] = {0, 0, 0}, swx = 0.0, stx = 0.0, slx = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) dx[i] = w[i] - WR(k, i);
#pragma unroll
        for (int e = 0; e < NEX; ++e) {
          const size_t o = (size_t)(Dm::R_X + e) * K + k;
          const double l = LPA(o);
          double f = -tx;
#pragma unroll
          for (int i = 0; i < NX; ++i) f += sgn(e, i) * dx[i];
          const double s = fmax(-f, TINY_S), rp = 0.0, wgt = l * rcp_fast(s), tau = 0.0;
          part[0] += s * l; part[1] = fmax(part[1], fabs(rp));
          swx += wgt; stx += tau; slx += l;
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            mvx[i] += wgt * sgn(e, i); bt[i] += tau * sgn(e, i); bl[i] += l * sgn(e, i);
#pragma unroll
            for (int j = 0; j < NX; ++j) Dl[i][j] += wgt * sgn(e, i) * sgn(e, j);
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) bx[i] = -mvx[i];
        part[6] += swx; part[10] -= stx; part[14] -= slx;
      }
      // u-trust rows
      {
        double du[NU], mvu[NU], swu = 0.0, stu = 0.0, slu = 0.0;
#pragma unroll
        for (int j = 0; j < NU; ++j) { du[j] = w[NX + j] - WR(k, NX + j); mvu[j] = 0.0; }
#pragma unroll
        for (int e = 0; e < NEU; ++e) {
          const size_t o = (size_t)(Dm::R_U + e) * K + k;
          const double l = LPA(o);
          double f = -tu;
#pragma unroll
          for (int j = 0; j < NU; ++j) f += sgn(e, j) * du[j];
          const double s = fmax(-f, TINY_S), rp = 0.0, wgt = l * rcp_fast(s), tau = 0.0;
          part[0] += s * l; part[1] = fmax(part[1], fabs(rp));
          swu += wgt; stu += tau; slu += l;
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            mvu[i] += wgt * sgn(e, i); bt[NX + i] += tau * sgn(e, i); bl[NX + i] += l * sgn(e, i);
#pragma unroll
            for (int j = 0; j < NU; ++j) Dl[NX + i][NX + j] += wgt * sgn(e, i) * sgn(e, j);
          }
        }
#pragma unroll
        for (int j = 0; j < NU; ++j) bu[j] = -mvu[j];
        part[7] += swu; part[11] -= stu; part[15] -= slu;
      }
      if (fr) {
        // position box
#pragma unroll
        for (int i = 0; i < D; ++i) {
          size_t o = (size_t)(Dm::R_P + i) * K + k;
          double l = LPA(o), s = fmax(sc.pos_hi - w[i], TINY_S), rp = 0.0, wgt = l * rcp_fast(s);
          part[0] += s * l; part[1] = fmax(part[1], fabs(rp));
          Dl[i][i] += wgt; bt[i] += wgt * rp; bl[i] += l;
          o = (size_t)(Dm::R_P + D + i) * K + k;
          l = LPA(o); s = fmax(w[i] - sc.pos_lo, TINY_S); rp = 0.0; wgt = l * rcp_fast(s);
          part[0] += s * l; part[1] = fmax(part[1], fabs(rp));
          Dl[i][i] += wgt; bt[i] -= wgt * rp; bl[i] -= l;
        }
        if (!BALL) {
          const double gz[4] = {w[NX] - sc.v_max, -w[NX], w[NX + 1] - sc.w_max, -w[NX + 1] - sc.w_max};
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const size_t o = (size_t)(Dm::R_V + r) * K + k;
            const double l = LPA(o), s = fmax(-gz[r], TINY_S), rp = 0.0, wgt = l * rcp_fast(s);
            const int c = NX + (r >> 1);
            const double sg = (r & 1) ? -1.0 : 1.0;
            part[0] += s * l; part[1] = fmax(part[1], fabs(rp));
            Dl[c][c] += wgt; bt[c] += sg * wgt * rp; bl[c] += sg * l;
          }
        } else {
          const size_t o = (size_t)Dm::R_V * K + k;
          const double l = LPA(o);
          double n2 = 0.0;
#pragma unroll
          for (int j = 0; j < NU; ++j) n2 += w[NX + j] * w[NX + j];
          // the ball row is quadratic: its slack stays an independent (stored) variable, rp = c(u) + s
          const double s = ws.s.sB[k];
          const double rp = 0.5 * (n2 - sc.v_max * sc.v_max) + s, wgt = l * rcp_fast(s);
          part[0] += s * l; part[1] = fmax(part[1], fabs(rp));
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            bt[NX + i] += wgt * rp * w[NX + i]; bl[NX + i] += l * w[NX + i];
            Dl[NX + i][NX + i] += l;
#pragma unroll
            for (int j = 0; j < NU; ++j) Dl[NX + i][NX + j] += wgt * w[NX + i] * w[NX + j];
          }
        }
        // hinge rows: this thread's share (all of them when G == 1)
        hinge_R_range(k, w, Dl, bt, bl);
      }
      // quadratic / linear position terms (enter both the rhs and the stationarity residual)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double ql = qlin ? qlin[(size_t)c * K + k] * sc.cs : 0.0;
        const double g = sc.qrho * w[c] + ql;
        bt[c] += g; bl[c] += g;
        Dl[c][c] += sc.qrho;
        part[17] += 0.5 * sc.qrho * w[c] * w[c] + ql * w[c];
      }
      if (game) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          double g, cv, ob;
          game_terms(k, i, w, g, cv, ob);
          bt[i] += g; bl[i] += g; Dl[i][i] += cv; part[17] += ob;
        }
      }
      // write own-stage pieces (interval pieces are added after the barrier)
      double* dk = Dk + (size_t)k * SD;
#pragma unroll
      for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) dk[i * NS + j] = Dl[i][j];
      double* rb = Rb + (size_t)k * SR;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        rb[0 * NS + i] = 0.0; rb[1 * NS + i] = 0.0;
        rb[2 * NS + i] = (i < NX) ? bx[i] : 0.0;
        rb[3 * NS + i] = (i >= NX) ? bu[i - NX] : 0.0;
        dW[k * NSP + i] = bt[i];
        dWa[k * NSP + i] = bl[i];     // stationarity accumulator (dWa is free at this point)
      }
    }
    if (C > 1) CSYNC(); else __syncthreads();
    PHASE(1);
    // ---- second half of the assembly: interval pieces (k-1 via Jn, k via Jp)
    double rdmax = 0.0;
    for (int k = tid; k < K; k += nthr) {
      const bool fr = (k > 0 && k < K - 1);
      double* dk = Dk + (size_t)k * SD;
      double* ek = Ek + (size_t)(k + 1 < K ? k + 1 : k) * SD;    // slot B of stage k+1 holds H[k+1, k]
      double* rb = Rb + (size_t)k * SR;
      if (!fr) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
#pragma unroll
          for (int j = 0; j < NS; ++j) { dk[i * NS + j] = (i == j) ? 1.0 : 0.0; ek[i * NS + j] = 0.0; }
#pragma unroll
          for (int c = 0; c < 4; ++c) rb[c * NS + i] = 0.0;
          dW[k * NSP + i] = 0.0;
        }
        continue;
      }
      double acc_t[NS], acc_l[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) { acc_t[i] = dW[k * NSP + i]; acc_l[i] = dWa[k * NSP + i]; }
      if (NSLOT > 0) {
        for (int g = 0; g < NSLOT; ++g) {
          const double* hs = HS + ((size_t)g * K + k) * HSW;
#pragma unroll
          for (int c = 0; c < D; ++c) {
            acc_t[c] += hs[D * D + c]; acc_l[c] += hs[D * D + D + c];
#pragma unroll
            for (int c2 = 0; c2 < D; ++c2) dk[c * NS + c2] += hs[c * D + c2];
          }
        }
      }
      // interval k (this stage is the "previous" node): Jp
      {
        const double* jac = JAC + (size_t)k * NJ;
        const double* st = ST + (size_t)k * STG;
        double J[NX][NS];
        load_Jp<Dm>(jac, J);
        const double M00 = st[0], M01 = st[1], M02 = st[2], M11 = st[3], M12 = st[4], M22 = st[5];
        double T[NX][NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          T[0][j] = M00 * J[0][j] + M01 * J[1][j] + M02 * J[2][j];
          T[1][j] = M01 * J[0][j] + M11 * J[1][j] + M12 * J[2][j];
          T[2][j] = M02 * J[0][j] + M12 * J[1][j] + M22 * J[2][j];
        }
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
          for (int j = 0; j < NS; ++j) dk[i * NS + j] += J[0][i] * T[0][j] + J[1][i] * T[1][j] + J[2][i] * T[2][j];
        // E_k = Jn_k' M_k Jp_k (rows: stage k+1, cols: stage k); zero when stage k+1 is fixed
        if (k + 1 < K - 1) {
          double Jn[NX][NS];
          load_Jn<Dm>(jac, Jn);
#pragma unroll
          for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int j = 0; j < NS; ++j) ek[i * NS + j] = Jn[0][i] * T[0][j] + Jn[1][i] * T[1][j] + Jn[2][i] * T[2][j];
          if (qpair) {
#pragma unroll
            for (int i = 0; i < NS; ++i) ek[i * NS + i] -= qpair[i] * sc.cs;
          }
        } else {
#pragma unroll
          for (int i = 0; i < NS * NS; ++i) ek[i] = 0.0;
        }
        const double* S = jac + NX * NX + 2 * NX * NU;
        double MJs[NX];
        MJs[0] = -(M00 * S[0] + M01 * S[1] + M02 * S[2]);
        MJs[1] = -(M01 * S[0] + M11 * S[1] + M12 * S[2]);
        MJs[2] = -(M02 * S[0] + M12 * S[1] + M22 * S[2]);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          rb[0 * NS + i] += J[0][i] * MJs[0] + J[1][i] * MJs[1] + J[2][i] * MJs[2];
          rb[1 * NS + i] -= J[0][i] * st[6] + J[1][i] * st[7] + J[2][i] * st[8];
          acc_t[i] += J[0][i] * st[10] + J[1][i] * st[11] + J[2][i] * st[12];
          acc_l[i] += J[0][i] * st[13] + J[1][i] * st[14] + J[2][i] * st[15];
        }
      }
      // interval k-1 (this stage is the "next" node): Jn
      {
        const double* jac = JAC + (size_t)(k - 1) * NJ;
        const double* st = ST + (size_t)(k - 1) * STG;
        double J[NX][NS];
        load_Jn<Dm>(jac, J);
        const double M00 = st[0], M01 = st[1], M02 = st[2], M11 = st[3], M12 = st[4], M22 = st[5];
        double T[NX][NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          T[0][j] = M00 * J[0][j] + M01 * J[1][j] + M02 * J[2][j];
          T[1][j] = M01 * J[0][j] + M11 * J[1][j] + M12 * J[2][j];
          T[2][j] = M02 * J[0][j] + M12 * J[1][j] + M22 * J[2][j];
        }
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
          for (int j = 0; j < NS; ++j) dk[i * NS + j] += J[0][i] * T[0][j] + J[1][i] * T[1][j] + J[2][i] * T[2][j];
        const double* S = jac + NX * NX + 2 * NX * NU;
        double MJs[NX];
        MJs[0] = -(M00 * S[0] + M01 * S[1] + M02 * S[2]);
        MJs[1] = -(M01 * S[0] + M11 * S[1] + M12 * S[2]);
        MJs[2] = -(M02 * S[0] + M12 * S[1] + M22 * S[2]);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          rb[0 * NS + i] += J[0][i] * MJs[0] + J[1][i] * MJs[1] + J[2][i] * MJs[2];
          rb[1 * NS + i] -= J[0][i] * st[6] + J[1][i] * st[7] + J[2][i] * st[8];
          acc_t[i] += J[0][i] * st[10] + J[1][i] * st[11] + J[2][i] * st[12];
          acc_l[i] += J[0][i] * st[13] + J[1][i] * st[14] + J[2][i] * st[15];
        }
      }
      if (k == 1) {
        // E_0 couples the fixed stage 0: zero (stage 0 thread wrote zeros already)
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        dW[k * NSP + i] = -acc_t[i];
        rdmax = fmax(rdmax, fabs(acc_l[i]));
        if (fix_sig) rb[i] = 0.0;
      }
    }
    // E_k for k = 0 is zero (stage 0 fixed) -- written by the !fr branch.  Reduce the partials.
    part[18] = rdmax;
    {
      const int ops[19] = {0, 2, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2};
      cluster_reduce(std::integral_constant<int, 19>{}, part, ops);
    }
    if (tid == 0) {
      // global rows
      const double gz[3] = {tx + tu + (sig - sc.sig_ref) - sc.r_tr, tx + tu - (sig - sc.sig_ref) - sc.r_tr, -sig};
      double Gg[4][4], bg[4], rdg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Gg[i][j] = 0.0;
      double comp = red[0], rp_inf = red[1];
#pragma unroll
      for (int c = 0; c < 4; ++c) { bg[c] = red[8 + c]; rdg[c] = red[12 + c]; }
      Gg[0][0] = red[3]; Gg[0][1] = Gg[1][0] = red[4]; Gg[1][1] = red[5]; Gg[2][2] = red[6]; Gg[3][3] = red[7];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double s = fmax(-gz[r], TINY_S), l = gl[15 + r], rp = 0.0, wgt = l * rcp_fast(s);
        gl[12 + r] = s;                         // recomputed slack, valid for the rest of this iteration
        comp += s * l; rp_inf = fmax(rp_inf, fabs(rp));
        gl[50 + r] = rp;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          bg[i] += gG(r, i) * wgt * rp; rdg[i] += gG(r, i) * l;
#pragma unroll
          for (int j = 0; j < 4; ++j) Gg[i][j] += wgt * gG(r, i) * gG(r, j);
        }
      }
      bg[0] += sc.c_sig; bg[1] += sc.c_tnu; rdg[0] += sc.c_sig; rdg[1] += sc.c_tnu;
      if (fix_sig) {       // sigma == sigma_ref: unit row/column, zero right-hand side, no stationarity condition
#pragma unroll
        for (int i = 0; i < 4; ++i) { Gg[0][i] = 0.0; Gg[i][0] = 0.0; }
        Gg[0][0] = 1.0; bg[0] = 0.0; rdg[0] = 0.0;
      }
      double rd_inf = fmax(red[2], red[18]);
#pragma unroll
      for (int c = 0; c < 4; ++c) { rd_inf = fmax(rd_inf, fabs(rdg[c])); gl[34 + c] = -bg[c]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gl[18 + i * 4 + j] = Gg[i][j];
      const double obj = sc.c_sig * sig + sc.c_tnu * tnu + red[16] + red[17];
      const double mu = comp * inv_n_rows;
      gl[44] = mu; gl[45] = comp;
      int flag = 0;
      if (!(mu == mu) || !(rp_inf == rp_inf) || !(rd_inf == rd_inf) || mu > 1e300) flag = 2;
      else {
        // dual residual is judged relative to the largest (scaled) cost coefficient; once the gap is far past its
        // target the residual sits at its round-off floor and a looser bound applies
        double cmax = fmax(sc.c_tnu, sc.c_sig);
        if (Mobs > 0) cmax = fmax(cmax, sc.hw_obs);
        if (a.n_nbr > 0) cmax = fmax(cmax, sc.hw_col);
        const double gscale = fmax(fabs(obj), 1e-3);
        const bool gap_ok = comp <= eps_gap * gscale, deep = comp <= 1e-4 * eps_gap * gscale;
        if (gap_ok && rp_inf <= eps_feas && rd_inf <= (deep ? 1e-5 : 1e-7) * (1.0 + cmax)) flag = 1;
        // stall rule (see oracle/ipm_struct.py): optimum tiny in scaled cost units -> the relative gap target is below
        // what fp64 delivers; accept once the gap is <= 1e-7 (1 + |obj|) and has stopped shrinking / rd blows up.
        else if (it > 0 && comp <= 1e-7 * (1.0 + fabs(obj)) && rp_inf <= eps_feas &&
                 (comp > 0.5 * gl[62] || (rd_inf > 10.0 * gl[63] && rd_inf > 1e-8 * (1.0 + cmax)))) flag = 1;
        gl[62] = comp; gl[63] = rd_inf;
      }
      gl[40] = (double)flag;
    }
    __syncthreads();
    CSYNC();
    {
      const int flag = (int)gll[40];
      if (flag == 1) { status = SCVX_ST_OPTIMAL; break; }
      if (flag == 2) { status = SCVX_ST_NUMERICAL; break; }
    }

    PHASE(2);
    // ------------------------------------------------------------------- factor + predictor solve (all threads)
    cr_factor<Dm>(Dk, Ek, Ck, K - 2, tid, nthr);
    PHASE(3);
    cr_forward<Dm, 5>(Ek, Ck, Rb, dW, K - 2, tid, nthr);
    PHASE(4);
    // Schur complement of the border:  S = Gg - sum_j (Li_j V_j)'(Li_j V_j),  rg = bg - sum_j (Li_j V_j)'(Li_j v_j)
    {
      double sp[14];
#pragma unroll
      for (int i = 0; i < 14; ++i) sp[i] = 0.0;
      for (int k = 1 + tid; k <= K - 2; k += nthr) {
        const double* li = Dk + (size_t)k * SD;
        double Z[5][NS];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          double* v = (c < 4) ? (Rb + (size_t)k * SR + c * NS) : (dW + k * NSP);
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j <= i; ++j) acc += li[i * NS + j] * v[j];
            Z[c][i] = acc;
          }
          // the backward sweep wants D^-1 v = Li' (Li v): finish it here, in place, while Li and Z are at hand
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = i; j < NS; ++j) acc += li[j * NS + i] * Z[c][j];
            v[i] = acc;
          }
        }
        int q = 0;
#pragma unroll
        for (int ca = 0; ca < 4; ++ca)
#pragma unroll
          for (int cb = ca; cb < 5; ++cb) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) acc += Z[ca][i] * Z[cb][i];
            sp[q++] += acc;
          }
      }
      const int ops[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      block_reduce<14>(sp, ops, red);
      if (tid == 0) {
        // unpack: pairs (ca, cb>=ca) in the order generated above; cb == 4 is the rhs column
        int q = 0;
#pragma unroll
        for (int ca = 0; ca < 4; ++ca)
#pragma unroll
          for (int cb = ca; cb < 5; ++cb) {
            const double v = red[q++];
            if (cb < 4) { gl[18 + ca * 4 + cb] -= v; if (cb != ca) gl[18 + cb * 4 + ca] -= v; }
            else gl[34 + ca] -= v;
          }
        schur_solve(gl + 18, gl + 34, gl + 4);
      }
      __syncthreads();
    }
    PHASE(5);
    cr_backward<Dm, 5>(Ek, Ck, Rb, dW, K - 2, tid, nthr);
    PHASE(6);
    // dWa = v - Y dg_aff
    {
      const double g0 = gl[4], g1 = gl[5], g2 = gl[6], g3 = gl[7];
      for (int k = tid; k < K; k += nthr) {
        const bool fr = (k > 0 && k < K - 1);
        const double* rb = Rb + (size_t)k * SR;
#pragma unroll
        for (int i = 0; i < NS; ++i)
          dWa[k * NSP + i] = fr ? dW[k * NSP + i] - (rb[i] * g0 + rb[NS + i] * g1 + rb[2 * NS + i] * g2 + rb[3 * NS + i] * g3) : 0.0;
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------------------------ PASS P / PASS S
    // A generic visitor over all rows of stage k, instantiated twice (straight-line code, no per-row mode tests):
    //   PASS 1 (P): statistics of the affine (predictor) step -- step-length ratios and the three sums for mu_aff -- AND the
    //               corrector's right-hand side.  The row term of that right-hand side is tau = (sigma mu - ds_a dl_a + l r_p) / s,
    //               linear in the centring target sigma mu that only the block-wide reduction of this very pass determines:
    //               so the pass accumulates the coefficient of sigma mu (1/s) and the rest separately and combines them after
    //               the reduction.  (The first version walked the rows a second time for this.)
    //   PASS 3 (S): step-length ratios of the final direction; the plain rows' multiplier step stays in registers and is applied
    //               in place once the step length is known, the hinge pairs' step goes to the d* buffers for the next R pass.
    PHASE(7);
    auto run_pass = [&](auto pass_c) -> bool {
      constexpr int PASS = decltype(pass_c)::value;
      constexpr bool P = (PASS == 1);
      CSYNC();                              // the direction this pass walks (dWa / dW) and sigma mu are final in the leader
      const double sigmu = gll[43];                                        // valid in PASS 3 (written by the tail of PASS 1)
      const double om = P ? 1.0 : gll[56];                                 // weight of the second-order term (PASS 3; see the tail of PASS 1)
      double pr[13];
#pragma unroll
      for (int i = 0; i < 13; ++i) pr[i] = 0.0;
      // pr: 0 max primal ratio, 1 max dual ratio, [P] 2 sum ds*l, 3 sum s*dl, 4 sum ds*dl, 5..7 border rhs pieces (t_nu, t_x, t_u)
      //     multiplying sigma mu, 8..10 the same, constant part, 11 / 12 the sigma-column piece (times sigma mu / constant);
      //     [S] 2 non-finite flag
      const double ga0 = gl[4], ga1 = gl[5], ga2 = gl[6], ga3 = gl[7];     // affine global step
      const double gd0 = gl[8], gd1 = gl[9], gd2 = gl[10], gd3 = gl[11];   // final global step (PASS 3)
      // Step-length ratios without divisions: the primal ratio -ds/s reuses the row's reciprocal 1/s; the dual ratio
      // -dl/l is tracked as a fraction (qdn/qdd) compared by cross-multiplication.  alpha = min(1, 1/max ratio).
      double qp = 0.0, qdn = 0.0, qdd = 1.0;
      auto upd_d = [&](double dl, double l) { if (-dl * qdd > qdn * l) { qdn = -dl; qdd = l; } };
      // per-row kernel.  P: returns the two pieces of tau (ta * sigma mu + tb).  S: stores the multiplier step to `pd`
      // (and the slack step of the nonlinear ball row to `psd`).
      auto row = [&](double l, double gz_h, double gdza, double gdz, double& ta, double& tb, double& dl_out, const double* ps = nullptr,
                     double* ds_out = nullptr, double* t0_out = nullptr) {
        const double s = ps ? *ps : fmax(-gz_h, TINY_S);          // stored slack only for the nonlinear ball row
        const double rs = rcp_fast(s);
        const double rp = ps ? gz_h + s : 0.0, wgt = l * rs;
        const double dsa = -rp - gdza, dla = -l - wgt * dsa;
        const double c2 = dsa * dla;
        if (P) {
          qp = fmax(qp, -dsa * rs); upd_d(dla, l);
          pr[2] += dsa * l; pr[3] += s * dla; pr[4] += c2;
          ta = rs; tb = -c2 * rs;                                  // tau = ta * sigma mu + om * tb (+ t0: only the ball row has r_p != 0)
          if (t0_out) *t0_out = l * rp * rs;
          return;
        }
        const double ds = -rp - gdz, dl = -l + (sigmu - om * c2) * rs - wgt * ds;
        qp = fmax(qp, -ds * rs); upd_d(dl, l);
        dl_out = dl;
        if (ds_out) *ds_out = ds;
      };
      double dlr[P ? 1 : NPLAIN], dsb = 0.0;     // S: the plain rows' multiplier step (and the ball row's slack step), in registers
#pragma unroll
      for (int r = 0; r < (P ? 1 : NPLAIN); ++r) dlr[r] = 0.0;
      auto DLR = [&](int r) -> double& { return dlr[P ? 0 : r]; };

      // own-stage rhs pieces of the corrector (P), kept in registers across the reduction: bta * sigma mu + om * btb + bt0, where
      // om weighs the second-order term (known, like sigma mu, only after the reduction)
      double bta[NS], btb[NS], bt0[P ? NS : 1];
#pragma unroll
      for (int i = 0; i < NS; ++i) { bta[i] = 0.0; btb[i] = 0.0; bt0[P ? i : 0] = 0.0; }
      auto BT0 = [&](int i) -> double& { return bt0[P ? i : 0]; };
      // one hinge pair: P -- affine-step statistics and the pair's three pieces of the corrector rhs (bta_ * sigma mu + om * btb_ + bt0_);
      //                 S -- step-length ratios of the final direction, the pair's step goes to the d* buffers
      auto hinge_PS = [&](int h, const HingeData<D>& hd, int k, const double* w, const double* da, const double* dz, double* bta_,
                          double* btb_, double* bt0_) {
        double av[D], ap = 0.0, ada = 0.0, adz = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) { av[c] = hd.a[c]; ap += av[c] * w[c]; ada += av[c] * da[c]; adz += av[c] * dz[c]; }
        const size_t o = (size_t)h * K + k;
        const double xi = hd.xi, l1 = hd.l1, l2 = hd.l2, hw = hinge_w(h);
        const double viol = hd.b - ap;
        const double s1 = fmax(xi - viol, TINY_S), s2 = fmax(xi, TINY_S);
        const double rs1 = rcp_fast(s1), rs2 = rcp_fast(s2);
        const double w1 = l1 * rs1, w2 = l2 * rs2, rw = rcp_fast(w1 + w2);
        // affine step of this hinge pair (sigma mu = 0, no second-order term; both primal residuals are zero by construction)
        const double dxia = (-hw - w1 * ada) * rw;
        const double ds1a = ada + dxia, ds2a = dxia;
        const double dl1a = -l1 - w1 * ds1a, dl2a = -l2 - w2 * ds2a;
        const double c1 = ds1a * dl1a, c2 = ds2a * dl2a;
        if (P) {
          qp = fmax(qp, fmax(-ds1a * rs1, -ds2a * rs2)); upd_d(dl1a, l1); upd_d(dl2a, l2);
          pr[2] += ds1a * l1 + ds2a * l2; pr[3] += s1 * dl1a + s2 * dl2a; pr[4] += c1 + c2;
          // t1 = (sigma mu - c1) / s1, t2 = (sigma mu - c2) / s2; rhs_xi = -hw + t1 + t2; th = t1 - w1 rhs_xi / (w1 + w2)
          const double t1b = -c1 * rs1, t2b = -c2 * rs2;
          const double tha = rs1 - w1 * (rs1 + rs2) * rw, thb = t1b - w1 * (t1b + t2b) * rw, th0 = w1 * hw * rw;
#pragma unroll
          for (int c = 0; c < D; ++c) { bta_[c] -= av[c] * tha; btb_[c] -= av[c] * thb; bt0_[c] -= av[c] * th0; }
          return;
        }
        const double t1 = (sigmu - om * c1) * rs1, t2 = (sigmu - om * c2) * rs2;
        const double rhs_xi = -hw + t1 + t2;
        const double dxi = (rhs_xi - w1 * adz) * rw;
        const double ds1 = adz + dxi, ds2 = dxi;
        const double dl1 = -l1 + t1 - w1 * ds1, dl2 = -l2 + t2 - w2 * ds2;
        qp = fmax(qp, fmax(-ds1 * rs1, -ds2 * rs2)); upd_d(dl1, l1); upd_d(dl2, l2);
        st_na(ws.dxi + o, dxi); st_na(ws.dl1 + o, dl1); st_na(ws.dl2 + o, dl2);
      };
      auto hinge_PS_range = [&](int k, const double* w, const double* da, const double* dz, double* bta_, double* btb_, double* bt0_) {
        for (int h0 = h_lo; h0 < h_hi; h0 += HINGE_CHUNK) {
          prefetch_hinges(h0, false);
          HingeData<D> hb[HINGE_CHUNK];
#pragma unroll
          for (int c = 0; c < HINGE_CHUNK; ++c) hb[c] = load_hinge(h0 + c, k, ws.s);
#pragma unroll
          for (int c = 0; c < HINGE_CHUNK; ++c)
            if (hb[c].on) hinge_PS(h0 + c, hb[c], k, w, da, dz, bta_, btb_, bt0_);
        }
      };
      {
        // helper groups: their share of stage kt's hinge rows (P: partial rhs pieces to the stage's owner)
        double ph[3 * D];
#pragma unroll
        for (int i = 0; i < 3 * D; ++i) ph[i] = 0.0;
        if (helper && kt > 0 && kt < K - 1) {
          double wl[D], dal[D], dzl[D];     // position and the two directions' position parts, once
#pragma unroll
          for (int c = 0; c < D; ++c) { wl[c] = Wl[kt * NSP + c]; dal[c] = dWal[kt * NSP + c]; dzl[c] = P ? 0.0 : dWl[kt * NSP + c]; }
          hinge_PS_range(kt, wl, dal, dzl, ph, ph + D, ph + 2 * D);
          if (P && grp > 0) {
            double* hs = HS + ((size_t)(grp - 1) * K + kt) * HSW;
#pragma unroll
            for (int i = 0; i < 3 * D; ++i) hs[i] = ph[i];
          }
        }
        if (P) deposit_remote(std::integral_constant<int, 3 * D>{}, ph);
      }
      for (int k = tid; k < K; k += nthr) {
        const double* w = W + k * NSP;
        const double* da = dWa + k * NSP;
        const double* dz = dW + k * NSP;       // valid in PASS 3 (final direction)
        const bool fr = (k > 0 && k < K - 1);
        if (k < K - 1) {
          const double* jac = JAC + (size_t)k * NJ;
          double nu[NX], nua[NX], nud[NX] = {0, 0, 0};
          nu_form<Dm, true>(jac, w, w + NSP, sig, nu);
          nu_form<Dm, false>(jac, da, da + NSP, ga0, nua);
          if (!P) nu_form<Dm, false>(jac, dz, dz + NSP, gd0, nud);
          double eta[NX] = {0, 0, 0}, etb[NX] = {0, 0, 0}, sta = 0.0, stb = 0.0;
#pragma unroll
          for (int e = 0; e < NEX; ++e) {
            const size_t o = (size_t)(Dm::R_NU + e) * K + k;
            double f = -tnu, fa = -ga1, fd = -gd1;
#pragma unroll
            for (int i = 0; i < NX; ++i) { f += sgn(e, i) * nu[i]; fa += sgn(e, i) * nua[i]; fd += sgn(e, i) * nud[i]; }
            double ta = 0.0, tb = 0.0;
            row(ws.s.lP[o], f, fa, fd, ta, tb, DLR(Dm::R_NU + e));
            if (P) {
              sta += ta; stb += tb;
#pragma unroll
              for (int i = 0; i < NX; ++i) { eta[i] += ta * sgn(e, i); etb[i] += tb * sgn(e, i); }
            }
          }
          if (P) {
            const double* S = jac + NX * NX + 2 * NX * NU;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
              ST2[(size_t)k * Dm::ST2 + i] = eta[i]; dW[k * NSP + i] = etb[i];      // dW is free until the combine below
              pr[11] -= eta[i] * S[i]; pr[12] -= etb[i] * S[i];
            }
            pr[5] -= sta; pr[8] -= stb;
          }
        }
        {
          double dx[NX];
#pragma unroll
          for (int i = 0; i < NX; ++i) dx[i] = w[i] - WR(k, i);
          double sxa = 0.0, sxb = 0.0;
#pragma unroll
          for (int e = 0; e < NEX; ++e) {
            const size_t o = (size_t)(Dm::R_X + e) * K + k;
            double f = -tx, fa = -ga2, fd = -gd2;
#pragma unroll
            for (int i = 0; i < NX; ++i) { f += sgn(e, i) * dx[i]; fa += sgn(e, i) * da[i]; fd += sgn(e, i) * dz[i]; }
            double ta = 0.0, tb = 0.0;
            row(ws.s.lP[o], f, fa, P ? 0.0 : fd, ta, tb, DLR(Dm::R_X + e));
            if (P) {
              sxa += ta; sxb += tb;
#pragma unroll
              for (int i = 0; i < NX; ++i) { bta[i] += ta * sgn(e, i); btb[i] += tb * sgn(e, i); }
            }
          }
          pr[6] -= sxa; pr[9] -= sxb;
          double sua = 0.0, sub = 0.0;
#pragma unroll
          for (int e = 0; e < NEU; ++e) {
            const size_t o = (size_t)(Dm::R_U + e) * K + k;
            double f = -tu, fa = -ga3, fd = -gd3;
#pragma unroll
            for (int j = 0; j < NU; ++j) {
              f += sgn(e, j) * (w[NX + j] - WR(k, NX + j)); fa += sgn(e, j) * da[NX + j]; fd += sgn(e, j) * dz[NX + j];
            }
            double ta = 0.0, tb = 0.0;
            row(ws.s.lP[o], f, fa, P ? 0.0 : fd, ta, tb, DLR(Dm::R_U + e));
            if (P) {
              sua += ta; sub += tb;
#pragma unroll
              for (int j = 0; j < NU; ++j) { bta[NX + j] += ta * sgn(e, j); btb[NX + j] += tb * sgn(e, j); }
            }
          }
          pr[7] -= sua; pr[10] -= sub;
        }
        if (fr) {
#pragma unroll
          for (int i = 0; i < D; ++i) {
            double ta = 0.0, tb = 0.0;
            size_t o = (size_t)(Dm::R_P + i) * K + k;
            row(ws.s.lP[o], w[i] - sc.pos_hi, da[i], P ? 0.0 : dz[i], ta, tb, DLR(Dm::R_P + i));
            if (P) { bta[i] += ta; btb[i] += tb; }
            o = (size_t)(Dm::R_P + D + i) * K + k;
            row(ws.s.lP[o], sc.pos_lo - w[i], -da[i], P ? 0.0 : -dz[i], ta, tb, DLR(Dm::R_P + D + i));
            if (P) { bta[i] -= ta; btb[i] -= tb; }
          }
          if (!BALL) {
            const double gz[4] = {w[NX] - sc.v_max, -w[NX], w[NX + 1] - sc.w_max, -w[NX + 1] - sc.w_max};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const size_t o = (size_t)(Dm::R_V + r) * K + k;
              const int c = NX + (r >> 1);
              const double sg = (r & 1) ? -1.0 : 1.0;
              double ta = 0.0, tb = 0.0;
              row(ws.s.lP[o], gz[r], sg * da[c], P ? 0.0 : sg * dz[c], ta, tb, DLR(Dm::R_V + r));
              if (P) { bta[c] += sg * ta; btb[c] += sg * tb; }
            }
          } else {
            const size_t o = (size_t)Dm::R_V * K + k;
            double n2 = 0.0, uda = 0.0, udz = 0.0;
#pragma unroll
            for (int j = 0; j < NU; ++j) { n2 += w[NX + j] * w[NX + j]; uda += w[NX + j] * da[NX + j]; udz += w[NX + j] * dz[NX + j]; }
            double ta = 0.0, tb = 0.0;
            double t0 = 0.0;
            row(ws.s.lP[o], 0.5 * (n2 - sc.v_max * sc.v_max), uda, P ? 0.0 : udz, ta, tb, DLR(Dm::R_V), ws.s.sB + k, &dsb, &t0);
            if (P) {
#pragma unroll
              for (int j = 0; j < NU; ++j) { bta[NX + j] += ta * w[NX + j]; btb[NX + j] += tb * w[NX + j]; BT0(NX + j) += t0 * w[NX + j]; }
            }
          }
          hinge_PS_range(k, w, da, dz, bta, btb, bt0);
        }
        if (P) {
          // gradient of the smooth cost terms: independent of sigma mu and of the second-order weight
#pragma unroll
          for (int c = 0; c < D; ++c) {
            const double ql = qlin ? qlin[(size_t)c * K + k] * sc.cs : 0.0;
            BT0(c) += sc.qrho * w[c] + ql;
          }
          if (game) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              double g, cv, ob;
              game_terms(k, i, w, g, cv, ob);
              BT0(i) += g;
            }
          }
        }
      }
      PHASE(8 + 4 * (PASS - 1) / 2);
      // ---- pass epilogues
      pr[0] = qp; pr[1] = qdn * rcp_fast(qdd);            // max ratios of this thread's rows
      if (P) {
        const int ops[13] = {2, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        cluster_reduce(std::integral_constant<int, 13>{}, pr, ops);
        if (tid == 0) {
          // global rows
          // (step lengths are kept as their reciprocals -- max ratios -- until the very end: two divisions in all)
          double qpm = fmax(red[0], 1.0), qdm = fmax(red[1], 1.0);
          double s2l = red[2], sdl = red[3], dd = red[4];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const double s = gl[12 + r], l = gl[15 + r], rp = gl[50 + r], rs = rcp_fast(s), wgt = l * rs;
            const double gdza = gG(r, 0) * ga0 + gG(r, 2) * ga2 + gG(r, 3) * ga3;
            const double dsa = -rp - gdza, dla = -l - wgt * dsa;
            qpm = fmax(qpm, -dsa * rs); qdm = fmax(qdm, -dla * rcp_fast(fmax(l, 1e-290)));
            s2l += dsa * l; sdl += s * dla; dd += dsa * dla;
            gl[53 + r] = dsa * dla;
          }
          double ap = 1.0 / qpm, ad = 1.0 / qdm;
          if (coupled) { ap = ad = fmin(ap, ad); }
          const double comp = gl[45];
          const double comp_aff = comp + ap * s2l + ad * sdl + ap * ad * dd;
          double sg = comp_aff * rcp_fast(comp);
          sg = fmin(fmax(sg, 0.0), 1.0);
          const double smu = sg * sg * sg * gl[44];
          gl[43] = smu;
          // Mehrotra's second-order term ds_a dl_a presumes a full affine step; after a short one it points the corrector at products
          // the iterate cannot reach.  It is weighed by om = min(alpha_p, alpha_d) of the affine step (oracle/ipm_struct.py: 23 % fewer
          // iterations over 256 sub-problems of the bench scenes together with the step fraction of pass S and mu0 = 1e-3).
          // (problems with one common step length -- quadratic cost or the ball row -- keep the full term: weighing it was measured
          // to cost 15 % more iterations on config 4)
          const double om = coupled ? 1.0 : fmax(fmin(ap, ad), om_floor);
          gl[56] = om;
          // border part of the corrector's right-hand side
          double bg[4] = {fma(smu, red[11], om * red[12]), fma(smu, red[5], om * red[8]), fma(smu, red[6], om * red[9]),
                          fma(smu, red[7], om * red[10])};
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const double s = gl[12 + r], l = gl[15 + r], rp = gl[50 + r];
            gl[53 + r] *= om;
            const double tau = (smu - gl[53 + r] + l * rp) * rcp_fast(s);
#pragma unroll
            for (int i = 0; i < 4; ++i) bg[i] += gG(r, i) * tau;
          }
          bg[0] += sc.c_sig; bg[1] += sc.c_tnu;
          if (fix_sig) bg[0] = 0.0;
#pragma unroll
          for (int c = 0; c < 4; ++c) gl[34 + c] = -bg[c];
        }
        __syncthreads();
        // own-stage + interval pieces of the corrector's right-hand side, now that sigma mu is known
        {
          const double smu = gl[43], om = gl[56];
          for (int k = tid; k < K; k += nthr) {
            const bool fr = (k > 0 && k < K - 1);
            if (!fr) continue;
            double ek[NX], ekm[NX], t[NS], t2[NS];
            const double* s0 = ST2 + (size_t)k * Dm::ST2;
            const double* s1 = ST2 + (size_t)(k - 1) * Dm::ST2;
#pragma unroll
            for (int i = 0; i < NX; ++i) { ek[i] = fma(smu, s0[i], om * dW[k * NSP + i]); ekm[i] = fma(smu, s1[i], om * dW[(k - 1) * NSP + i]); }
            JpT<Dm>(JAC + (size_t)k * NJ, ek, t);
            JnT<Dm>(JAC + (size_t)(k - 1) * NJ, ekm, t2);
            if (NSLOT > 0) {
              for (int g = 0; g < NSLOT; ++g) {
                const double* hs = HS + ((size_t)g * K + k) * HSW;
#pragma unroll
                for (int c = 0; c < D; ++c) { bta[c] += hs[c]; btb[c] += hs[D + c]; BT0(c) += hs[2 * D + c]; }
              }
            }
#pragma unroll
            for (int i = 0; i < NS; ++i) btb[i] = -(fma(smu, bta[i], fma(om, btb[i], BT0(i))) + t[i] + t2[i]);
          }
          __syncthreads();            // every thread has read its neighbour's staging out of dW
          for (int k = tid; k < K; k += nthr) {
            const bool fr = (k > 0 && k < K - 1);
#pragma unroll
            for (int i = 0; i < NS; ++i) dW[k * NSP + i] = fr ? btb[i] : 0.0;
          }
        }
        __syncthreads();
        // corrector solve (all threads): rg = bg - Y'b ; v = T^-1 b ; dg = S^-1 rg ; dW = v - Y dg
        {
          double yb[4] = {0.0, 0.0, 0.0, 0.0};
          for (int k = 1 + tid; k <= K - 2; k += nthr) {
            const double* rb = Rb + (size_t)k * SR;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int i = 0; i < NS; ++i) yb[c] += rb[c * NS + i] * dW[k * NSP + i];
          }
          const int ops4[4] = {0, 0, 0, 0};
          block_reduce<4>(yb, ops4, red);
          if (tid == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) gl[34 + c] -= red[c];
            schur_solve(gl + 18, gl + 34, gl + 8);
          }
          // (the barrier inside the first level of cr_forward orders these writes before any read of gl[8..11])
        }
        PHASE(18);
        cr_forward<Dm, 1>(Ek, Ck, Rb, dW, K - 2, tid, nthr);
        PHASE(16);
        cr_apply_dinv_all<Dm>(Dk, dW, K - 2, tid, nthr);
        cr_backward<Dm, 1>(Ek, Ck, Rb, dW, K - 2, tid, nthr);
        PHASE(17);
        {
          const double g0 = gl[8], g1 = gl[9], g2 = gl[10], g3 = gl[11];
          for (int k = tid; k < K; k += nthr) {
            const bool fr = (k > 0 && k < K - 1);
            const double* rb = Rb + (size_t)k * SR;
#pragma unroll
            for (int i = 0; i < NS; ++i)
              dW[k * NSP + i] = fr ? dW[k * NSP + i] - (rb[i] * g0 + rb[NS + i] * g1 + rb[2 * NS + i] * g2 + rb[3 * NS + i] * g3) : 0.0;
          }
        }
        __syncthreads();
      } else {
        // a non-finite direction must never be applied: the iterate stays at the last good (primal feasible) point
        double bad = 0.0;
        for (int k = tid; k < K; k += nthr)
#pragma unroll
          for (int i = 0; i < NS; ++i) bad = fmax(bad, isfinite(dW[k * NSP + i]) ? 0.0 : 1.0);
        pr[2] = bad;
        const int ops[3] = {2, 2, 2};
        cluster_reduce(std::integral_constant<int, 3>{}, pr, ops);
        if (tid == 0) {
          double qpm = fmax(red[0], 1.0), qdm = fmax(red[1], 1.0);
          const bool nan_step = red[2] > 0.0 || !isfinite(red[0]) || !isfinite(red[1]) || !isfinite(gd0 + gd1 + gd2 + gd3);
          gl[40] = nan_step ? 2.0 : 0.0;
          double dsg[3], dlg[3];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const double s = gl[12 + r], l = gl[15 + r], rp = gl[50 + r], rs = rcp_fast(s), wgt = l * rs;
            const double gdz = gG(r, 0) * gd0 + gG(r, 2) * gd2 + gG(r, 3) * gd3;
            const double ds = -rp - gdz, dl = -l + (sigmu - gl[53 + r]) * rs - wgt * ds;
            qpm = fmax(qpm, -ds * rs); qdm = fmax(qdm, -dl * rcp_fast(fmax(l, 1e-290)));
            dsg[r] = ds; dlg[r] = dl;
          }
          double ap = 1.0 / qpm, ad = 1.0 / qdm;
          if (coupled) { ap = ad = fmin(ap, ad); }
          // fraction of the way to the boundary: 0.999 while the centring target is large, up to 1 - 1e-6 as it vanishes (the
          // payload-free form of Mehrotra's step-to-boundary rule: same iteration counts on the twin as the full rule)
          const double sf = fmin(sf_cap, fmax(0.999, 1.0 - 1000.0 * sigmu));
          ap = fmin(1.0, sf * ap); ad = fmin(1.0, sf * ad);
          gl[41] = ap; gl[42] = ad;
          if (!nan_step) {
            // the globals move now; the stage variables below; the row state in the next residual pass
#pragma unroll
            for (int c = 0; c < 4; ++c) gl[c] += ap * gl[8 + c];
#pragma unroll
            for (int r = 0; r < 3; ++r) { gl[12 + r] += ap * dsg[r]; gl[15 + r] += ad * dlg[r]; }
          }
        }
        __syncthreads();
        CSYNC();
        if ((int)gll[40] == 2) return false;    // uniform: leave before anything is applied
        {
          const double ap = gl[41], ad = gl[42];
          for (int k = tid; k < K; k += nthr) {
            const bool fr = (k > 0 && k < K - 1);
#pragma unroll
            for (int i = 0; i < NS; ++i) W[k * NSP + i] += ap * dW[k * NSP + i];
            // multipliers of this stage's plain rows, in place (the same rows the passes visit).  All loads of a chunk are issued
            // before its first store: written row by row, ptxas cannot prove that the store of one row and the load of the next do
            // not alias and serialises 20 of the 28 L2 round trips (r02w source page: ~450 stall samples on each of them, 5.3 % of
            // the kernel's samples on this one statement).
            constexpr int LCH = 14;
            if (BALL) {      // (the single-integrator kernels, short of registers, keep the row-by-row form)
#pragma unroll
              for (int r = 0; r < NPLAIN; ++r) {
                const bool visited = (r < NEX) ? (k < K - 1) : ((r < Dm::R_P) ? true : fr);
                if (visited) {
                  const size_t o = (size_t)r * K + k;
                  st_na(ws.s.lP + o, fma(ad, DLR(r), ws.s.lP[o]));
                }
              }
            }
#pragma unroll
            for (int r0 = 0; r0 < (BALL ? 0 : NPLAIN); r0 += LCH) {
              double cur[LCH];
#pragma unroll
              for (int j = 0; j < LCH; ++j) {
                const int r = r0 + j;
                const bool visited = (r < NPLAIN) && ((r < NEX) ? (k < K - 1) : ((r < Dm::R_P) ? true : fr));
                cur[j] = visited ? ws.s.lP[(size_t)r * K + k] : 0.0;
              }
#pragma unroll
              for (int j = 0; j < LCH; ++j) {
                const int r = r0 + j;
                const bool visited = (r < NPLAIN) && ((r < NEX) ? (k < K - 1) : ((r < Dm::R_P) ? true : fr));
                if (visited) st_na(ws.s.lP + (size_t)r * K + k, fma(ad, DLR(r < NPLAIN ? r : 0), cur[j]));
              }
            }
            if (BALL && fr) st_na(ws.s.sB + k, fma(ap, dsb, ws.s.sB[k]));
          }
        }
        __syncthreads();
      }
      PHASE(9 + 4 * (PASS - 1) / 2);
      return true;
    };
    if (run_pass(std::integral_constant<int, 1>{})) run_pass(std::integral_constant<int, 3>{});
    if ((int)gll[40] == 2) { status = SCVX_ST_NUMERICAL; break; }  // non-finite step refused (iterate untouched)
  }

  // ---- epilogue: outputs in the reference's layouts --------------------------------------------------------
  CSYNC();                                  // (the last step-length pass moved the stage variables after its own cluster barrier)
  {
    const double sig = gl[0];
    double* Xo = a.X + (size_t)agent * NX * K;
    double* Uo = a.U + (size_t)agent * NU * K;
    double* No = a.nu + (size_t)agent * NX * (K - 1);
    double pr[4] = {0.0, 0.0, 0.0, 0.0};   // max_k |nu_k|_1, sum obstacle slack, sum collision slack, quad+lin
    for (int k = tid; k < K; k += nthr) {
      const double* w = W + k * NSP;
#pragma unroll
      for (int i = 0; i < NX; ++i) Xo[(size_t)i * K + k] = w[i];
#pragma unroll
      for (int j = 0; j < NU; ++j) Uo[(size_t)j * K + k] = w[NX + j];
      if (k < K - 1) {
        double nu[NX], s = 0.0;
        nu_form<Dm, true>(JAC + (size_t)k * NJ, w, w + NSP, sig, nu);
#pragma unroll
        for (int i = 0; i < NX; ++i) { No[(size_t)i * (K - 1) + k] = nu[i]; s += fabs(nu[i]); }
        pr[0] = fmax(pr[0], s);
      }
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double ql = qlin ? qlin[(size_t)c * K + k] : 0.0;
        pr[3] += 0.5 * (a.quad_rho ? a.quad_rho[agent] : 0.0) * w[c] * w[c] + ql * w[c];
      }
      if (game) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          double g, cv, ob;
          game_terms(k, i, w, g, cv, ob);
          pr[3] += ob / sc.cs;
        }
      }
    }
    // hinge values at every node (fixed end nodes included), every group its share
    if (kt < K) {
      const int k = kt;
      double w[D];
#pragma unroll
      for (int c = 0; c < D; ++c) w[c] = Wl[k * NSP + c];
      for (int h = h_lo; h < h_hi; ++h) {
        double ap = 0.0;
        const bool on = hinge_on(h);
        if (on) {
#pragma unroll
          for (int c = 0; c < D; ++c) ap += hinge_a(h, c, k) * w[c];
        }
        const double v = on ? fmax(0.0, hinge_b(h, k) - ap) : 0.0;
        if (h < Mobs) { a.s_prime[((size_t)agent * Mobs + h) * K + k] = v; pr[1] += v; }
        else {
          if (a.col_slack) a.col_slack[((size_t)agent * a.n_nbr + (h - Mobs)) * K + k] = v;
          pr[2] += v;
        }
      }
    }
    const int ops[4] = {2, 0, 0, 0};
    cluster_reduce(std::integral_constant<int, 4>{}, pr, ops);
    if (tid == 0) {
      a.sigma[agent] = sig;
      a.objective[agent] = a.weight_nu * red[0] + a.weight_slack * red[1] + a.weight_col * red[2] + a.weight_sigma * sig + red[3];
      a.status[agent] = status;
      a.iters[agent] = it;
    }
  }
  CSYNC();                                  // no block of the cluster leaves while another may still read its shared memory
  PHASE(19);
  if (RETRY) {                              // next entry of the failed list
    __syncthreads();
    unit += (int)gridDim.x / C;
    goto next_unit;
  }
}

// Compact list of the agents a retry pass has to solve: list[0] = count, list[1..] = their ids (any order).
static __global__ void __launch_bounds__(256) retry_list_kernel(int n, const int* __restrict__ status, const int* __restrict__ active,
                                                        int* __restrict__ list) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (status[i] != SCVX_ST_OPTIMAL && (!active || active[i])) list[1 + atomicAdd(&cnt, 1)] = i;
  __syncthreads();
  if (threadIdx.x == 0) list[0] = cnt;
}
constexpr int RETRY_BLOCKS = 8;

constexpr size_t SMEM_LIMIT = 227 * 1024;
// Two blocks per SM need <= (228 KB - 2 x 1 KB reserved) / 2 each; keep the Jacobians in shared memory only if that holds
// or if the problem does not fit otherwise anyway.
constexpr size_t SMEM_TWO_PER_SM = 113 * 1024;

template <class M>
size_t solver_smem_bytes(int K, bool jac_in_smem, int G = 1, int C = 1) {
  using Dm = Dims<M>;
  const size_t slots = (size_t)(G - 1) + (size_t)(C - 1);
  return ((size_t)K * (jac_in_smem ? Dm::PER_STAGE : Dm::PER_STAGE_NOJAC) + Dm::small_of(G) + slots * K * Dm::HSW + (size_t)(C - 1) * 24) *
         sizeof(double);
}
template <class M>
bool solver_jac_in_smem(int K) {
  const size_t with = solver_smem_bytes<M>(K, true), without = solver_smem_bytes<M>(K, false);
  if (with <= SMEM_TWO_PER_SM) return true;          // fits twice per SM either way
  if (without <= SMEM_TWO_PER_SM) return false;      // dropping the Jacobians buys the second block
  return with <= SMEM_LIMIT;                         // one block per SM: keep them on chip if possible
}
// Hinge groups: only where the block is alone on its SM anyway (its shared memory is > half an SM's) and a stage has enough
// hinge rows to share out; SCVX_HINGE_GROUPS (1, 2 or 4) overrides the choice for experiments.
template <class M>
int solver_hinge_groups(int K, int NH) {
  if (K > 128) return 1;
  int want = (NH >= 64 && solver_smem_bytes<M>(K, solver_jac_in_smem<M>(K)) > SMEM_TWO_PER_SM) ? 2 : 1;
  if (const char* e = getenv("SCVX_HINGE_GROUPS")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) want = v;
  }
  while (want > 1 && solver_smem_bytes<M>(K, solver_jac_in_smem<M>(K), want) > SMEM_LIMIT) want >>= 1;
  return want;
}
template <class M>
size_t solver_ws_doubles_per_agent(int K, int NH) {
  using Dm = Dims<M>;
  return (size_t)K * (Dm::NPLAIN + 6 * (size_t)NH + (Dm::BALL ? 1 : 0));     // row state + the hinge pairs' pending step
}
template <class M>
size_t solver_ws_total_doubles(int n_agents, int K, int NH) {
  using Dm = Dims<M>;
  size_t tot = solver_ws_doubles_per_agent<M>(K, NH) * (size_t)n_agents;
  if (!solver_jac_in_smem<M>(K)) tot += (size_t)n_agents * K * Dm::NJ;
  return tot;
}

// launch parameters that do not depend on the problem (the rules at the tail of pass P / pass S, with their experiment switches:
// SCVX_OM_FLOOR=1 SCVX_SF_CAP=0.999 is plain Mehrotra)
struct IpmTuning { double mu0_def, om_floor, sf_cap; };
inline IpmTuning ipm_tuning() {
  static const IpmTuning t = {getenv("SCVX_MU0_DEFAULT") ? atof(getenv("SCVX_MU0_DEFAULT")) : 10.0,
                              getenv("SCVX_OM_FLOOR") ? atof(getenv("SCVX_OM_FLOOR")) : 0.0,
                              getenv("SCVX_SF_CAP") ? atof(getenv("SCVX_SF_CAP")) : 0.999999};
  return t;
}
template <class Kern>
int launch_ipm_kernel(Kern kern, int C, const scvx_solve_args& a, cudaStream_t st, size_t smem, int threads, size_t jac_off) {
  const IpmTuning t = ipm_tuning();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  if (C == 1) {
    kern<<<a.n_agents, threads, smem, st>>>(a, t.mu0_def, /*eps_gap=*/1e-8, /*eps_feas=*/1e-9, jac_off, t.om_floor, t.sf_cap);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.n_agents * C);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, a, t.mu0_def, 1e-8, 1e-9, jac_off, t.om_floor, t.sf_cap);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx");
  }
  SCVX_CHECK_LAUNCH("scvx_solve_batched");
  return SCVX_OK;
}

// Instantiations with the sizes of the BASELINE configs as compile-time constants live in translation units of their own; same
// code, results equal to round-off (the compiler contracts other multiply-add pairs once sizes and absent terms are constants)
// (solver_unicycle_fixed.cu, solver_si_fixed.cu: they compile in parallel with the generic ones).  Returns IPM_NOT_FIXED when no
// instantiation matches the arguments; the generic kernel runs then.  SCVX_NO_FIXED=1 switches them off (experiments / tests).
constexpr int IPM_NOT_FIXED = 1;
template <class M>
int launch_ipm_fixed(const scvx_solve_args& a, cudaStream_t st, size_t smem, int threads, size_t jac_off, bool jsm, int G, int C);
template <>
int launch_ipm_fixed<Unicycle>(const scvx_solve_args& a, cudaStream_t st, size_t smem, int threads, size_t jac_off, bool jsm, int G, int C);
template <>
int launch_ipm_fixed<SingleIntegrator>(const scvx_solve_args& a, cudaStream_t st, size_t smem, int threads, size_t jac_off, bool jsm, int G, int C);
// the plain sub-problem of SCProblem: obstacle rows only, no optional term
inline bool ipm_args_plain(const scvx_solve_args& a) {
  return a.n_nbr == 0 && !a.quad_rho && !a.lin_p && !a.quad_diag && !a.lin_w && !a.quad_pair && !a.fix_sigma && !a.col_a && !a.col_b && !a.col_mask;
}

template <class M, bool JSM, int G, int C>
int launch_ipm_g(const scvx_solve_args& a, cudaStream_t st, size_t smem, int threads, size_t jac_off) {
  if constexpr (G == 1 && C == 1) {
    if (a.retry_failed && a.retry_list) {
      // compact retry pass: list the failed agents, then a few blocks walk the list
      const IpmTuning t = ipm_tuning();
      auto rk = ipm_kernel<M, JSM, 1, 1, true>;
      cudaError_t e2 = cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e2 != cudaSuccess) return cuda_fail(e2, "cudaFuncSetAttribute");
      retry_list_kernel<<<1, 256, 0, st>>>(a.n_agents, a.status, a.active, a.retry_list);
      rk<<<a.n_agents < RETRY_BLOCKS ? a.n_agents : RETRY_BLOCKS, threads, smem, st>>>(a, t.mu0_def, 1e-8, 1e-9, jac_off, t.om_floor, t.sf_cap);
      SCVX_CHECK_LAUNCH("scvx_solve_batched (retry)");
      return SCVX_OK;
    }
  }
  const bool no_fixed = getenv("SCVX_NO_FIXED") != nullptr;      // (read per call: the tests switch it)
  if (!no_fixed) {
    const int r = launch_ipm_fixed<M>(a, st, smem, threads, jac_off, JSM, G, C);
    if (r != IPM_NOT_FIXED) return r;
  }
  return launch_ipm_kernel(ipm_kernel<M, JSM, G, C, false>, C, a, st, smem, threads, jac_off);
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <class M>
int launch_ipm(const scvx_solve_args& a, cudaStream_t st) {
  const bool jac_smem = solver_jac_in_smem<M>(a.K);
  int G = solver_hinge_groups<M>(a.K, a.M + a.n_nbr);
  // Cluster size: hinge-heavy problems (G > 1) on a launch that would leave most SMs idle get C blocks per agent -- e.g. one
  // GPU's 32 agents of config 4 on 8 GPUs.  SCVX_CLUSTER (1, 2, 4) overrides.
  int C = 1;
  if (G > 1 && jac_smem) {
    const int sms = sm_count();
    if (a.n_agents * 4 <= sms) C = 4; else if (a.n_agents * 2 <= sms) C = 2;
    if (const char* e = getenv("SCVX_CLUSTER")) {
      const int v = atoi(e);
      if (v == 1 || v == 2 || v == 4) C = v;
    }
    if (C > 1) G = 2;                      // cluster kernels are instantiated for two hinge groups per block
    while (C > 1 && solver_smem_bytes<M>(a.K, jac_smem, G, C) > SMEM_LIMIT) C >>= 1;
  }
  const size_t smem = solver_smem_bytes<M>(a.K, jac_smem, G, C);
  if (smem > SMEM_LIMIT) {
    snprintf(g_last_error, sizeof(g_last_error), "K=%d needs %zu B of shared memory per agent (> 227 KB)", a.K, smem);
    return SCVX_E_UNSUPPORTED;
  }
  int threads = ((a.K + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  if (threads > SOLVER_MAX_THREADS) threads = SOLVER_MAX_THREADS;
  if (a.K > threads) {        // the passes keep per-stage partial results in registers across block reductions: one stage per thread
    snprintf(g_last_error, sizeof(g_last_error), "K=%d exceeds the %d threads of an agent's block", a.K, threads);
    return SCVX_E_UNSUPPORTED;
  }
  if (G > 1) threads = 128 * G;
  const size_t jac_off = solver_ws_doubles_per_agent<M>(a.K, a.M + a.n_nbr) * (size_t)a.n_agents;
  if (C == 4) return launch_ipm_g<M, true, 2, 4>(a, st, smem, threads, jac_off);
  if (C == 2) return launch_ipm_g<M, true, 2, 2>(a, st, smem, threads, jac_off);
  if (G == 4) return jac_smem ? launch_ipm_g<M, true, 4, 1>(a, st, smem, threads, jac_off) : launch_ipm_g<M, false, 4, 1>(a, st, smem, threads, jac_off);
  if (G == 2) return jac_smem ? launch_ipm_g<M, true, 2, 1>(a, st, smem, threads, jac_off) : launch_ipm_g<M, false, 2, 1>(a, st, smem, threads, jac_off);
  return jac_smem ? launch_ipm_g<M, true, 1, 1>(a, st, smem, threads, jac_off) : launch_ipm_g<M, false, 1, 1>(a, st, smem, threads, jac_off);
}


// per-phase cycle counters of this translation unit's kernels (all zero unless built with -DSCVX_PHASE_TIMING)
inline int phase_cycles_of_this_unit(unsigned long long* out32, int reset) {
#ifdef SCVX_PHASE_TIMING
  cudaError_t e = cudaMemcpyFromSymbol(out32, g_phase_cycles, 32 * sizeof(unsigned long long));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyFromSymbol");
  if (reset) {
    unsigned long long z[32] = {0};
    e = cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyToSymbol");
  }
#else
  for (int i = 0; i < 32; ++i) out32[i] = 0ull;
  (void)reset;
#endif
  return SCVX_OK;
}

}  // namespace scvx
