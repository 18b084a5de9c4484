/*
 * scvx_b200.h -- C-ABI of the B200-native SCvx inner loop (libscvx_b200.so).
 *
 * Drop-in boundary for the hot path of
 *   shiivashaakeri/Dynamic-Programming-MultiAgent-Trajectory-Optimiziation
 * (paths below are relative to that repository).  The reference has no FFI of its own -- its
 * "operator API" is a set of Python classes -- so each entry point names the Python method whose
 * arithmetic it replaces; the Python mirror in scvx_b200/ binds these with ctypes
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain C types only; every array pointer is a DEVICE pointer (cudaMalloc / torch CUDA tensor)
 *     unless the parameter name ends in _h (host).  `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream).  Nothing here allocates device memory except the explicit
 *     workspace-size query + caller-provided workspace, and the *_host convenience calls.
 *   - all arithmetic is IEEE fp64.
 *   - batched layout is agent-major, then the reference's own per-agent numpy layout, C-order:
 *       X      [n_agents][n_x][K]          (numpy (n_x, K) per agent)
 *       U      [n_agents][n_u][K]
 *       A_bar  [n_agents][n_x*n_x][K-1]    column k = Phi_k flattened order='F'
 *       B_bar  [n_agents][n_x*n_u][K-1]    C_bar likewise;  S_bar, z_bar [n_agents][n_x][K-1]
 *     so a batch-of-1 slice is bit-compatible with the arrays the reference produces.
 *   - return value: 0 on success, negative SCVX_E_* on error (nothing is written on argument errors).
 *     Per-agent solver outcomes are written to `status` arrays (SCVX_ST_*).
 *   - thread-safety: calls are re-entrant; concurrent calls must use distinct streams/workspaces.
 */
#ifndef SCVX_B200_H
#define SCVX_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SCVX_ABI_VERSION 1

/* model ids (SCvx/models/unicycle_model.py, SCvx/models/single_integrator_model.py) */
#define SCVX_MODEL_UNICYCLE 0           /* n_x=3 n_u=2 d=2 */
#define SCVX_MODEL_SINGLE_INTEGRATOR 1  /* n_x=3 n_u=3 d=3 */
#define SCVX_MODEL_USER_BASE 16         /* ids >= this: models registered at run time, see scvx_user_model_register */

/* error codes */
#define SCVX_OK 0
#define SCVX_E_BADARG (-1)
#define SCVX_E_CUDA (-2)
#define SCVX_E_WORKSPACE (-3)
#define SCVX_E_UNSUPPORTED (-4)

/* per-agent solver status words */
#define SCVX_ST_OPTIMAL 0
#define SCVX_ST_MAXITER 1
#define SCVX_ST_NUMERICAL 2   /* NaN / non-positive pivot that regularisation could not repair */

int scvx_abi_version(void);
/* last CUDA error string seen by this library on the calling thread (never NULL) */
const char* scvx_last_error(void);
/* n_x, n_u, position dimension d of a model id; returns SCVX_E_BADARG for unknown ids */
int scvx_model_dims(int model_id, int* n_x, int* n_u, int* d);

/* ---------------------------------------------------------------------------------------------
 * Model plug-in (BaseModel contract, SCvx/models/base_model.py:16-88; the reference accepts any model whose
 * get_equations() returns f, A = df/dx, B = df/du -- its own models derive them with sympy, unicycle_model.py:54-63).
 * `program` is a CUDA C++ translation unit holding `struct scvx::UserModel { NX, NU, D, eval, f_only }` generated from the
 * user's symbolic dynamics, the stage-1 templates of csrc/foh_kernels.cuh and the three kernels scvx_user_foh /
 * scvx_user_piecewise / scvx_user_full (scvx_b200/codegen.py builds it).  It is compiled for sm_100a with NVRTC
 * (`nvrtc_path`: optional path of libnvrtc.so.12 to try first; the library dlopens it) and *model_id receives an id
 * >= SCVX_MODEL_USER_BASE that the stage-1 entry points accept; stages 2-3 depend on a model's dimensions and constraint kind
 * only and run a registered model through the kernels of its shape ((n_x, n_u, d) = (3, 2, 2) or (3, 3, 3); other shapes have
 * stage 1 only and scvx_solve_batched returns SCVX_E_BADARG for them).  On a compile error: SCVX_E_BADARG and the compiler's
 * log in scvx_user_model_log(). */
int scvx_user_model_register(const char* program, int n_x, int n_u, int d, const char* nvrtc_path, int* model_id);
const char* scvx_user_model_log(void);

/* ---------------------------------------------------------------------------------------------
 * Stage 1 -- first-order-hold discretisation.
 * Replaces FirstOrderHold.calculate_discretization (SCvx/discretization/first_order_hold.py:52-87)
 * and its right-hand side _ode_dVdt (:89-125), for a whole batch of agents at once.
 * One thread integrates one (agent, interval) with classical RK4, n_sub sub-steps over [0, dt],
 * dt = 1/(K-1).  n_sub = 0 selects the sub-step count per interval on the device from
 * sigma*dt*max(1,|u|) so that the result is within ~1e-10 relative of the exact ODE solution.
 * sigma: [n_agents].
 */
int scvx_foh_batched(int model_id, int n_agents, int K, int n_sub,
                     const double* X, const double* U, const double* sigma,
                     double* A_bar, double* B_bar, double* C_bar, double* S_bar, double* z_bar,
                     void* stream);

/* FirstOrderHold.integrate_nonlinear_piecewise (first_order_hold.py:127-140): X_nl[:,0]=X_lin[:,0],
 * X_nl[:,k+1] = flow of xdot=f(x,u(t)) over [0, dt*sigma] from X_lin[:,k].  X_nl: [n_agents][n_x][K]. */
int scvx_integrate_piecewise_batched(int model_id, int n_agents, int K, int n_sub,
                                     const double* X_lin, const double* U, const double* sigma,
                                     double* X_nl, void* stream);
/* FirstOrderHold.integrate_nonlinear_full (first_order_hold.py:142-155): x0 [n_agents][n_x]. */
int scvx_integrate_full_batched(int model_id, int n_agents, int K, int n_sub,
                                const double* x0, const double* U, const double* sigma,
                                double* X_nl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 2 -- constraint linearisation.
 *
 * Obstacle half-spaces (UnicycleModel.get_constraints, SCvx/models/unicycle_model.py:103-114;
 * SingleIntegratorModel twin, single_integrator_model.py:113-126):
 *   a_jk = (p_ref_k - c_j) / (||p_ref_k - c_j||_2 + 1e-6),   b_jk = clearance_j + a_jk . c_j
 *   so that the constraint reads  a_jk . p_k + s'_jk >= b_jk.
 * obs_c [n_agents][M][d], obs_clear [n_agents][M] (= r_j + r_rob (+margin));
 * out: obs_a [n_agents][M][d][K], obs_b [n_agents][M][K].
 */
int scvx_linearize_obstacles_batched(int model_id, int n_agents, int K, int M,
                                     const double* X_ref, const double* obs_c, const double* obs_clear,
                                     double* obs_a, double* obs_b, void* stream);

/* Inter-agent half-spaces (MultiAgentModel.linearize_collision, SCvx/models/multi_agent_model.py:61-79;
 * SI_MultiAgentModel.linearize_inter_agent_collision, SI_multi_agent_model.py:49-74), all ordered pairs
 * between `n_local` agents (global ids i0..i0+n_local-1) and all n_agents neighbours:
 *   a_ijk = (p_i,k - q_j,k)/(||.||_2 + 1e-6),   b_ijk = d_min + a_ijk . q_j,k
 * X_own [n_local][n_x][K] are the linearisation references of the local agents; X_nbr
 * [n_agents][n_x][K] the neighbours' current trajectories.  Slot j==i0+i is written as zeros.
 * out: col_a [n_local][n_agents][d][K], col_b [n_local][n_agents][K].
 */
int scvx_linearize_collision_batched(int model_id, int n_local, int i0, int n_agents, int K, double d_min,
                                     const double* X_own, const double* X_nbr,
                                     double* col_a, double* col_b, void* stream);

/* Neighbour culling for large N (a documented DEVIATION from the reference, whose agents couple with ALL others):
 * scvx_cross_min_dist2: d2 [n_local][n_agents] = min over k of the squared position distance between local agent i
 * (X_own [n_local][n_x][K]) and agent j (X_all [n_agents][n_x][K]).
 * scvx_linearize_collision_indexed: the half-spaces of scvx_linearize_collision_batched for a per-agent LIST of neighbours,
 * nbr_idx [n_local][n_sel] (global agent ids, -1 = empty slot -> zero row): col_a [n_local][n_sel][d][K], col_b [n_local][n_sel][K]. */
int scvx_cross_min_dist2(int model_id, int n_local, int n_agents, int K, const double* X_own, const double* X_all, double* d2,
                         void* stream);
int scvx_linearize_collision_indexed(int model_id, int n_local, int n_sel, int n_agents, int K, double d_min, const double* X_own,
                                     const double* X_all, const int* nbr_idx, double* col_a, double* col_b, void* stream);

/* Slab rows of the Nash best response: GameUnicycleModel.update_slabs (SCvx/models/game_model.py:56-67) and the rows
 * z_jk.(p_ik - Y_jk) >= collision_radius of get_cost_function (game_model.py:118-124), in the half-space layout of
 * scvx_linearize_collision_batched (a = z, b = radius + z.Y).  z_jk = (P_own_ik - X_dir_jk)/||.|| -- exactly zero when the
 * norm is below 1e-6, as in the reference (no 1e-6 in the denominator).  radius [n_local]; P_own [n_local][n_x][K];
 * X_dir, X_off [n_agents][n_x][K] (the neighbours' trajectories that give the direction resp. the offset Y);
 * degenerate [n_local] or NULL: number of vanished normals per agent (the reference's problem is infeasible then). */
int scvx_slab_normals_batched(int model_id, int n_local, int i0, int n_agents, int K, const double* radius,
                              const double* P_own, const double* X_dir, const double* X_off, double* col_a,
                              double* col_b, int* degenerate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 3 -- convex sub-problem, batched.
 * Replaces SCProblem.solve (SCvx/optimization/sc_problem.py:15-105, problem statement :21-83 with the
 * model rows of unicycle_model.py:85-115 / single_integrator_model.py:79-128) and, when n_nbr > 0,
 * AgentSolver.setup+solve (SCvx/optimization/agent_solver.py:43-117).  cvxpy+ECOS is replaced by a
 * structure-exploiting primal-dual interior-point method, one thread block per agent, block-tridiagonal
 * KKT factor resident in shared memory.  See DESIGN.md for the formulation.
 *
 * All per-agent inputs are arrays over agents.  Scalars that the reference keeps global
 * (weights) are passed per call.
 */
typedef struct scvx_solve_args {
  int model_id, n_agents, K;
  int M;             /* obstacles per agent (half-space tables below) */
  int n_nbr;         /* inter-agent half-space slots per agent (0 = plain SCProblem) */
  int max_iter;      /* IPM iteration cap (<=0: default 120; hinge-heavy QPs of config 4 average 40 iterations with a tail beyond 80) */
  int norm1_induced; /* 1: cvxpy matrix 1-norm (max column abs-sum) -- the reference's semantics */
  /* FOH matrices (stage 1 output layout) */
  const double *A_bar, *B_bar, *C_bar, *S_bar, *z_bar;
  /* trust-region centre */
  const double *X_ref, *U_ref, *sigma_ref;        /* [n][n_x][K], [n][n_u][K], [n] */
  const double *tr_radius;                        /* [n] */
  /* boundary conditions and bounds */
  const double *x_init, *x_final;                 /* [n][n_x] */
  const double *pos_lo, *pos_hi;                  /* [n]  (lower_bound + r_rob, upper_bound - r_rob) */
  const double *v_max, *w_max;                    /* [n]  (w_max ignored for the single integrator) */
  /* obstacle half-spaces  a.p + s' >= b  (stage 2 layout) */
  const double *obs_a, *obs_b;                    /* [n][M][d][K], [n][M][K] */
  /* inter-agent half-spaces a.(p - Y) + S >= d_min  <=>  a.p + S >= b  with b = d_min + a.Y */
  const double *col_a, *col_b;                    /* [n][n_nbr][d][K], [n][n_nbr][K] */
  const unsigned char *col_mask;                  /* [n][n_nbr] 1 = slot active (NULL: all active) */
  /* augmented-Lagrangian terms on positions: quad_rho/2 * ||P||^2 + <lin_p, P>  */
  const double *quad_rho;                         /* [n] or NULL */
  const double *lin_p;                            /* [n][d][K] or NULL */
  double weight_nu, weight_slack, weight_sigma, weight_col;
  /* outputs */
  double *X, *U, *nu, *sigma;                     /* [n][n_x][K], [n][n_u][K], [n][n_x][K-1], [n] */
  double *s_prime;                                /* [n][M][K]  obstacle slacks (hinge values) */
  double *col_slack;                              /* [n][n_nbr][K] or NULL */
  double *objective;                              /* [n] objective value incl. slack terms */
  int *status, *iters;                            /* [n] */
  /* scratch */
  void* workspace; unsigned long long workspace_bytes;
  /* Best-response (Nash game) terms -- AgentBestResponse.setup (SCvx/optimization/agent_best_response.py:36-100) with the
   * cost of GameUnicycleModel.get_cost_function (SCvx/models/game_model.py:69-126); all NULL / 0 for the plain problems.
   *   sum_k sum_i [ quad_diag_i/2 w_ik^2 + lin_w_ik w_ik ] + sum_{k<K-1} sum_i quad_pair_i/2 (w_i,k+1 - w_ik)^2,
   *   w_k = (x_k, u_k), i over the n_x + n_u stage components: control effort c ||U||^2 -> quad_diag_u = 2c; inertia
   *   c ||X - X_prev||^2 -> quad_diag_x = 2c, lin_w_x = -2c X_prev; control rate / curvature c sum (dw)^2 -> quad_pair = 2c.
   * fix_sigma != 0 adds the constraint sigma == sigma_ref (agent_best_response.py:76-77): the sigma column is frozen. */
  const double *quad_diag;                        /* [n][n_x+n_u] or NULL */
  const double *lin_w;                            /* [n][n_x+n_u][K] or NULL */
  const double *quad_pair;                        /* [n][n_x+n_u] or NULL */
  int fix_sigma;
  /* Launch order of the agents' thread blocks: block b solves agent block_order[b] (a permutation of 0..n-1), NULL = identity.
   * Agents differ 4x in interior-point iterations; starting the long ones first (scvx_order_by_iters on the previous outer
   * iteration's counts) removes most of the idle tail of the launch.  Results do not depend on the order. */
  const int *block_order;
  /* Per-agent start value of the barrier parameter mu (n_agents doubles), NULL or a non-positive entry = the default 10 (the
   * cold start).  A small start pays when the reference trajectory is the previous solution (a warm sub-problem): with 1e-3 the
   * bench workload's mean interior-point iteration count falls 8.7 -> 5.5 (scvx_mu0_from_iters builds the array from the previous
   * solve's iteration counts).  Two safeguards in the kernel: a problem whose start point violates a hinge row (the trajectory
   * runs through an obstacle of the new linearisation) starts cold whatever mu0 says, and a solve from a small start that has not
   * converged after 40 iterations ends with SCVX_ST_MAXITER -- see retry_failed below.  Ignored by cluster launches. */
  const double *mu0;
  /* Outer-loop activity flags (n_agents ints, NULL = all active): the block of an agent whose flag is 0 -- its outer loop has
   * converged (scvx_outer_update) -- returns at once, iters = 0, every other output keeps its previous value. */
  const int *active;
  /* Retry pass.  A solve that starts below the default barrier parameter (mu0[i] < 10) and has not converged after 40
   * interior-point iterations ends with SCVX_ST_MAXITER (1 in ~20 000 warm solves of the bench scenes strands like this; every one
   * of them converges from the cold start).  Launch the same arguments again with retry_failed = 1 and mu0 = NULL: blocks of agents
   * whose status is SCVX_ST_OPTIMAL return at once and leave every output alone, the others are solved from the cold start. */
  int retry_failed;
  /* Scratch of the retry pass: n_agents + 1 ints (count, then the ids of the failed agents).  With it the retry pass is a list
   * kernel plus a launch of a few blocks that walk the list; NULL (or a launch with hinge groups / clusters) = a full-grid launch
   * whose blocks of optimal agents return at once. */
  int *retry_list;
} scvx_solve_args;

/* order[r] = index of the agent with the r-th LARGEST iters (ties by index): a longest-first launch order for the next solve */
int scvx_order_by_iters(int n_agents, const int* iters, int* order, void* stream);

/* mu0[i] = (0 < iters[i] <= easy_max_iters) ? mu0_easy : mu0_hard -- the per-agent barrier start of the NEXT solve from this
 * solve's iteration counts (scvx_solve_args.mu0).  No reference counterpart (cvxpy/ECOS choose their own start). */
int scvx_mu0_from_iters(int n_agents, const int* iters, int easy_max_iters, double mu0_easy, double mu0_hard, double* mu0,
                        void* stream);
/* bytes of device workspace scvx_solve_batched needs for these sizes */
unsigned long long scvx_solve_workspace_bytes(int model_id, int n_agents, int K, int M, int n_nbr);
int scvx_solve_batched(const scvx_solve_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Consensus round (ADMMCoordinator.solve, SCvx/optimization/admm_coordinator.py:80-96 and
 * admm_utils.py:8-31):  Y+ = (Y + P)/2;  Lambda += rho (P - Y+);
 * pr[j] = ||P_j - Y+_j||_F, du[j] = ||Y+_j - Y_j||_F.   P, Y, Lambda: [n_agents][d][K] (in place).
 */
int scvx_consensus_update(int n_agents, int d, int K, double rho, const double* P,
                          double* Y, double* Lambda, double* pr, double* du, void* stream);
/* the same with the positions read in place from the agents' states X [n_agents][n_x][K] (rows 0..d-1) */
int scvx_consensus_update_x(int n_agents, int n_x, int d, int K, double rho, const double* X, double* Y, double* Lambda,
                            double* pr, double* du, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Everything of an ADMM round between the position exchange and the sub-problem solve, for the local agents
 * i0 .. i0+n_local-1 of a rank, in scvx_solve_args' layout (csrc/admm_round.cu):
 *   col_a[i][q][c][k]  normals of MultiAgentModel.linearize_collision (SCvx/models/multi_agent_model.py:61-79, SI twin
 *                      SI_multi_agent_model.py:49-74) about (X_own_i, X_all_j),
 *   col_b[i][q][k]     = d_min + a . Y_j   (collision rows of AgentSolver, SCvx/optimization/agent_solver.py:79-90),
 *   lin_p[i][c][k]     = sum_j Lambda_j - rho sum_j Y_j,   quad_rho[i] = rho * #active neighbours,
 *   aug_const[i]       = rho/2 sum ||Y_j||^2 - sum <Lambda_j, Y_j>     (agent_solver.py:92-95 collapsed),
 *   mask_out[i][q]     1 = slot active.
 * Slot q is agent q (nbr_idx NULL: all pairs, n_slots == n_agents) or agent nbr_idx[i][q] (-1 = empty); the own slot and slots
 * with mask_in[i][q] == 0 (mask_in may be NULL) are inactive.  X_own [n_local][n_x][K], X_all [n_agents][n_x][K],
 * Y, Lambda [n_agents][d][K]. */
int scvx_admm_round_prep(int model_id, int n_local, int i0, int n_agents, int K, int n_slots, double d_min, double rho,
                         const double* X_own, const double* X_all, const double* Y, const double* Lambda, const int* nbr_idx,
                         const unsigned char* mask_in, double* col_a, double* col_b, unsigned char* mask_out, double* lin_p,
                         double* quad_rho, double* aug_const, void* stream);
/* k_sel nearest neighbours per local agent from scvx_cross_min_dist2's table d2 [n_local][n_agents] (used as scratch: selected
 * entries are struck out), ascending, own column excluded, entries beyond radius2 (> 0) excluded, unused slots -1.
 * No reference counterpart (the reference couples all pairs). */
int scvx_knn_select(int n_local, int i0, int n_agents, int k_sel, double radius2, double* d2, int* nbr_idx, void* stream);
/* mask[i][j] = d2[i][j] <= radius2 (neighbour culling of the all-pairs tables; no reference counterpart) */
int scvx_radius_mask(int n_local, int n_agents, double radius2, const double* d2, unsigned char* mask, void* stream);

/* On-device SCvx bookkeeping for one outer iteration of a batch (SCVXSolver.solve,
 * SCvx/optimization/scvx_solver.py:82-111, :125-133): metrics, convergence flag, trust-region update and
 * iterate acceptance, without a host round trip.  metrics: [n_agents][6] = nu_norm, slack_norm, dx, du, ds, sigma_new.
 * `active` [n_agents] (1 = still iterating) is updated in place; converged agents keep their OLD iterate.
 */
int scvx_outer_update(int model_id, int n_agents, int K, int M, double conv_tol,
                      const double* X_new, const double* U_new, const double* nu_new, const double* sigma_new,
                      const double* s_prime, double* X, double* U, double* sigma, double* tr_radius,
                      int* active, double* metrics, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Distributed_opt scripts -- per-robot perturbation QP of the exact-ZOH double integrator.
 * Replaces the cvxpy -> CLARABEL solves inside x_traj_opt:
 *   Distributed_opt/ADMM_decentralized.py:52-98  (n=4, m=2, c_w=100, rho/lin/sbar set, nq=0)
 *   Distributed_opt/dist_scvx_3d.py:51-111       (n=6, m=3, c_w=1, rho=0, nq = N-1 collision rows, c_S=1e4)
 *   min c_w sum_{t<T-1} |u_t+w_t|^2 + sum_t [ lin_t.dpos_t + rho/2 |dpos_t - sbar_t|^2 ] + c_S sum_t S_t
 *   s.t. d_0=0, d_{T-1} = x_des - x_{T-1}, x_{t+1}+d_{t+1} = A(x_t+d_t) + B(u_t+w_t), |w_t|_1 <= tr,
 *        box on (x_t+d_t)[0:2], h_tq - g_tq.d_t[0:n/2] <= S_t, S_t >= 0          (t = 0..T-2)
 * Interior-point method with a Riccati recursion per Newton step; one thread block per robot.
 */
typedef struct scvx_lti_args {
  int n_robots, T, n, m;         /* (n, m) = (4, 2) or (6, 3) */
  int nq;                        /* collision rows per time step (0: none) */
  int max_iter;                  /* IPM iteration cap (<=0: default 60) */
  const double *Ad, *Bd;         /* [n*n], [n*m] row-major, shared by all robots (descete_f) */
  const double *x, *u;           /* [R][T][n], [R][T][m]  current trajectory (row T-1 of u ignored) */
  const double *x_des;           /* [R][n] */
  double tr, c_w, rho, c_S;
  double box_lo0, box_hi0, box_lo1, box_hi1;
  const double *lin, *sbar;      /* [R][T][2] or NULL */
  const double *col_h, *col_g;   /* [R][nq][T-1], [R][nq][n/2][T-1] or NULL */
  double *d, *w;                 /* out: [R][T][n], [R][T][m] (row T-1 of w is 0) */
  double *S;                     /* out: [R][T] collision slack (hinge value), may be NULL */
  double *objective;             /* out: [R] */
  int *status, *iters;           /* out: [R] */
  void* workspace; unsigned long long workspace_bytes;
} scvx_lti_args;
unsigned long long scvx_lti_qp_workspace_bytes(int n_robots, int T, int n, int nq);
int scvx_lti_qp_batched(const scvx_lti_args* args, void* stream);

/* Per-(robot, t) consensus QP of ADMM_decentralized.py:106-139 (separable over t):
 *   min_{sbar_t in R^2, S_t >= 0} r_t.(s_t - sbar_t) + rho/2 |s_t - sbar_t|^2 + c_S S_t   s.t. h_tq - g_tq.sbar_t <= S_t.
 * s_pos, r_dual, sbar: [R][T][2]; col_h [R][nq][T]; col_g [R][nq][2][T]; S_out [R][T] or NULL;
 * workspace: n_robots*T*2*nq doubles. */
int scvx_sbar_qp_batched(int n_robots, int T, int nq, double rho, double c_S, const double* s_pos, const double* r_dual,
                         const double* col_h, const double* col_g, double* sbar, double* S_out,
                         void* workspace, unsigned long long workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Callers either side of the loop (SURVEY section 8 f, ranks 2 and 4).
 *
 * scvx_warm_start_batched -- replaces initial_guess() of SCvx/utils/initial_guess.py:61-107 (model UNICYCLE: tangent
 * way-points around inflated discs, headings from forward differences, U0 = 0) and of SCvx/utils/IS_initial_guess.py:87-126
 * (model SINGLE_INTEGRATOR: detour way-points around inflated spheres, U0 = forward differences / dt, dt = 1/(K-1)).
 *   p0, p1 [n_agents][3] (unicycle: x, y, theta -- theta is ignored, as in the reference); obs_c [n_agents][M_max][d];
 *   obs_r [n_agents][M_max] (NOT inflated; `clearance` is added here); obs_count [n_agents] or NULL (= M_max each);
 *   X0 [n_agents][3][K]; U0 [n_agents][n_u][K].
 *   status [n_agents]: 0 ok; 1 start/goal inside or on an inflated obstacle that the straight segment crosses (the reference
 *   raises ValueError "Point inside/on circle"); 2 the last segment would get a negative sample count (numpy.linspace raises
 *   ValueError); 3 start and goal closer than 1e-6 (ValueError in compute_detour_waypoints).  X0/U0 of such agents are zeroed.
 */
int scvx_warm_start_batched(int model_id, int n_agents, int K, int M_max, const double* p0, const double* p1,
                            const double* obs_c, const double* obs_r, const int* obs_count, double clearance,
                            double* X0, double* U0, int* status, void* stream);

/* scvx_min_inter_agent_distance -- replaces min_inter_agent_distance (SCvx/utils/analysis.py:10-31):
 *   d_mat [n_agents][n_agents] (symmetric, zero diagonal) = min over k of || X_i[0:n_rows, k] - X_j[0:n_rows, k] ||_2;
 *   d_min [1] = smallest strictly positive entry (+inf when there is none: the reference raises on an empty selection).
 *   X [n_agents][n_x][K]; the reference always takes n_rows = 3 (for the unicycle that includes the heading row).
 * scvx_min_agent_obstacle_distance -- replaces min_agent_obstacle_distance (analysis.py:34-62):
 *   d_mat [n_agents][M] = min over k of || X_i[0:n_rows, k] - c_j || - (robot_radius + r_j); d_min [1] = its minimum
 *   (+inf when n_agents*M == 0).  obs_c [M][n_rows], obs_r [M]. */
int scvx_min_inter_agent_distance(int n_agents, int K, int n_x, int n_rows, const double* X, double* d_mat, double* d_min,
                                  void* stream);
int scvx_min_agent_obstacle_distance(int n_agents, int K, int n_x, int n_rows, int M, const double* X, const double* obs_c,
                                     const double* obs_r, double robot_radius, double* d_mat, double* d_min, void* stream);

/* scvx_intersample_batched -- inter-sample obstacle clearance (SURVEY 8 f rank 3) for the segment flow of make_segment_f
 * (SCvx/utils/intersample_collision.py:100-125): per (agent, segment k, obstacle j) the interior minima t* of
 * h(t) = || x(t)[0:proj_dim] - c_j || - r_j found as find_critical_times does (:29-72: central-difference phi on a grid of
 * num_samples points of [eps, t_range - eps], sign changes, <= 30 bisections to |b - a| < tol, keep 0 < t* < t_range with
 * phi2 > 0), and at each t* the linearisation of linearize_h (:75-97): h0 and the central-difference gradient in x_k
 * (grad_u is identically zero in the reference: the segment flow ignores its control argument).
 *   X [n][n_x][K], U [n][n_u][K], sigma [n] (the reference passes sigma = 1.0), obs_c [n][M][proj_dim], obs_r [n][M];
 *   items are ordered [agent][segment][obstacle]; n_roots [items] (may exceed max_roots: the rest is dropped);
 *   t_star, h0 [items][max_roots]; grad_x [items][max_roots][n_x].
 * scvx_clearance_samples_batched -- h at t = i/resolution, i < resolution, of every segment for one obstacle per agent
 * (the intent of compute_intersample_clearance, SCvx/utils/analysis.py:64-104): h_cont [n][K-1][resolution]. */
int scvx_intersample_batched(int model_id, int n_agents, int K, int M, int proj_dim, const double* X, const double* U,
                             const double* sigma, const double* obs_c, const double* obs_r, double t_range, int num_samples,
                             double eps, double tol, int max_roots, int* n_roots, double* t_star, double* h0, double* grad_x,
                             void* stream);
int scvx_clearance_samples_batched(int model_id, int n_agents, int K, int proj_dim, int resolution, const double* X,
                                   const double* U, const double* sigma, const double* obs_c, const double* total_r,
                                   double* h_cont, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Measurement helpers (bench.py only; not on the product path).
 * scvx_probe_fp64: `blocks` x 256 threads x `iters` x 8 independent DFMA; *flops_h (HOST pointer) receives the flop
 * count of the launch; `out` needs blocks*256 doubles.  scvx_l2_flush: write sweep over a buffer (> L2) between timed steps.
 */
int scvx_probe_fp64(int blocks, int iters, double* out, double* flops_h, void* stream);
int scvx_l2_flush(double* buf, unsigned long long n_doubles, void* stream);
/* scvx_debug_phase_cycles: SM cycles block 0 of scvx_solve_batched spent in each kernel phase (out32: 32 counters on the HOST;
 * all zero unless the library was built with -DSCVX_PHASE_TIMING, see tools/phase_timing.py). */
int scvx_debug_phase_cycles(unsigned long long* out32, int reset);

#ifdef __cplusplus
}
#endif
#endif /* SCVX_B200_H */
