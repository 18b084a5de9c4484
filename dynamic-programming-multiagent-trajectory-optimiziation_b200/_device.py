"""Batched device operations: thin, shape-checked wrappers over the C-ABI (include/scvx_b200.h).

Inputs and outputs are torch CUDA float64 tensors in the agent-major layouts the header documents.
torch is used for device memory and streams only; every number is produced by libscvx_b200.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import SolveArgs, check, load, ptr, stream_ptr

F64 = torch.float64



class _ModelDims(dict):
    """model id -> (n_x, n_u, d); ids of models registered at run time (codegen.register) are looked up in the library."""

    def __missing__(self, model_id):
        n_x, n_u, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(load().scvx_model_dims(int(model_id), ctypes.byref(n_x), ctypes.byref(n_u), ctypes.byref(d)), "scvx_model_dims")
        self[model_id] = (n_x.value, n_u.value, d.value)
        return self[model_id]


MODEL_DIMS = _ModelDims({_lib.MODEL_UNICYCLE: (3, 2, 2), _lib.MODEL_SINGLE_INTEGRATOR: (3, 3, 3)})


def _dev(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == F64):
        raise _lib.ScvxError("expected a float64 CUDA tensor (there is no CPU path)")
    return t.contiguous()


def foh(model_id, X, U, sigma, n_sub=0, out=None):
    """FirstOrderHold.calculate_discretization for a batch (first_order_hold.py:52-87).

    X (n, n_x, K), U (n, n_u, K), sigma (n,) -> A_bar (n, n_x*n_x, K-1), B_bar, C_bar (n, n_x*n_u, K-1),
    S_bar, z_bar (n, n_x, K-1)."""
    n_x, n_u, _ = MODEL_DIMS[model_id]
    X, U, sigma = _dev(X), _dev(U), _dev(sigma)
    n, _, K = X.shape
    assert X.shape == (n, n_x, K) and U.shape == (n, n_u, K) and sigma.shape == (n,)
    if out is None:
        out = tuple(torch.empty((n, r, K - 1), dtype=F64, device=X.device)
                    for r in (n_x * n_x, n_x * n_u, n_x * n_u, n_x, n_x))
    A, B, C, S, z = out
    check(load().scvx_foh_batched(model_id, n, K, n_sub, ptr(X), ptr(U), ptr(sigma), ptr(A), ptr(B), ptr(C),
                                  ptr(S), ptr(z), stream_ptr()), "scvx_foh_batched")
    return out


def integrate_piecewise(model_id, X_lin, U, sigma, n_sub=0):
    n_x, n_u, _ = MODEL_DIMS[model_id]
    X_lin, U, sigma = _dev(X_lin), _dev(U), _dev(sigma)
    n, _, K = X_lin.shape
    out = torch.empty_like(X_lin)
    check(load().scvx_integrate_piecewise_batched(model_id, n, K, n_sub, ptr(X_lin), ptr(U), ptr(sigma), ptr(out),
                                                  stream_ptr()), "scvx_integrate_piecewise_batched")
    return out


def integrate_full(model_id, x0, U, sigma, n_sub=0):
    n_x, n_u, _ = MODEL_DIMS[model_id]
    x0, U, sigma = _dev(x0), _dev(U), _dev(sigma)
    n, _, K = U.shape
    out = torch.empty((n, n_x, K), dtype=F64, device=U.device)
    check(load().scvx_integrate_full_batched(model_id, n, K, n_sub, ptr(x0), ptr(U), ptr(sigma), ptr(out),
                                             stream_ptr()), "scvx_integrate_full_batched")
    return out


def linearize_obstacles(model_id, X_ref, obs_c, obs_clear, out=None):
    """Obstacle half-spaces a.p + s' >= b (unicycle_model.py:103-114).
    X_ref (n, n_x, K), obs_c (n, M, d), obs_clear (n, M) -> obs_a (n, M, d, K), obs_b (n, M, K)."""
    _, _, d = MODEL_DIMS[model_id]
    X_ref, obs_c, obs_clear = _dev(X_ref), _dev(obs_c), _dev(obs_clear)
    n, _, K = X_ref.shape
    M = obs_c.shape[1]
    assert obs_c.shape == (n, M, d) and obs_clear.shape == (n, M)
    if out is None:
        out = (torch.empty((n, M, d, K), dtype=F64, device=X_ref.device),
               torch.empty((n, M, K), dtype=F64, device=X_ref.device))
    a, b = out
    check(load().scvx_linearize_obstacles_batched(model_id, n, K, M, ptr(X_ref), ptr(obs_c), ptr(obs_clear), ptr(a),
                                                  ptr(b), stream_ptr()), "scvx_linearize_obstacles_batched")
    return out


def linearize_collision(model_id, X_own, X_nbr, d_min, i0=0, out=None):
    """Inter-agent half-spaces (multi_agent_model.py:61-79) between the local agents (global ids
    i0..i0+n_local-1, references X_own) and every agent in X_nbr.
    -> col_a (n_local, n_agents, d, K), col_b (n_local, n_agents, K); slot j == i0+i is zero."""
    _, _, d = MODEL_DIMS[model_id]
    X_own, X_nbr = _dev(X_own), _dev(X_nbr)
    nl, _, K = X_own.shape
    na = X_nbr.shape[0]
    if out is None:
        out = (torch.empty((nl, na, d, K), dtype=F64, device=X_own.device),
               torch.empty((nl, na, K), dtype=F64, device=X_own.device))
    a, b = out
    check(load().scvx_linearize_collision_batched(model_id, nl, i0, na, K, float(d_min), ptr(X_own), ptr(X_nbr),
                                                  ptr(a), ptr(b), stream_ptr()), "scvx_linearize_collision_batched")
    return out


def consensus_update(P, Y, Lam, rho):
    """Y <- (Y+P)/2, Lam += rho (P - Y+) in place; returns per-agent (primal, dual) residual norms
    (admm_coordinator.py:80-96, admm_utils.py:8-31).  P, Y, Lam: (n, d, K)."""
    P, Y, Lam = _dev(P), _dev(Y), _dev(Lam)
    n, d, K = P.shape
    pr = torch.empty(n, dtype=F64, device=P.device)
    du = torch.empty(n, dtype=F64, device=P.device)
    check(load().scvx_consensus_update(n, d, K, float(rho), ptr(P), ptr(Y), ptr(Lam), ptr(pr), ptr(du), stream_ptr()),
          "scvx_consensus_update")
    return pr, du


def consensus_update_x(X, d, Y, Lam, rho, out=None):
    """consensus_update with the positions read in place from the agents' states X (n, n_x, K) (rows 0..d-1)."""
    X, Y, Lam = _dev(X), _dev(Y), _dev(Lam)
    n, n_x, K = X.shape
    pr, du = out if out is not None else (torch.empty(n, dtype=F64, device=X.device), torch.empty(n, dtype=F64, device=X.device))
    check(load().scvx_consensus_update_x(n, n_x, d, K, float(rho), ptr(X), ptr(Y), ptr(Lam), ptr(pr), ptr(du), stream_ptr()),
          "scvx_consensus_update_x")
    return pr, du


class AdmmRoundTables:
    """Outputs of scvx_admm_round_prep for (n_local agents, n_slots neighbour slots), allocated once."""

    def __init__(self, model_id, n_local, n_slots, K, device):
        _, _, d = MODEL_DIMS[model_id]
        self.col_a = torch.empty((n_local, n_slots, d, K), dtype=F64, device=device)
        self.col_b = torch.empty((n_local, n_slots, K), dtype=F64, device=device)
        self.mask = torch.empty((n_local, n_slots), dtype=torch.uint8, device=device)
        self.lin_p = torch.empty((n_local, d, K), dtype=F64, device=device)
        self.quad_rho = torch.empty(n_local, dtype=F64, device=device)
        self.aug_const = torch.empty(n_local, dtype=F64, device=device)


def admm_round_prep(model_id, X_own, X_all, Y, Lam, d_min, rho, i0, tables, nbr_idx=None, mask_in=None):
    """Collision half-spaces about (own reference, neighbours' trajectories), their right-hand sides d_min + a.Y_j and the
    collapsed augmented-Lagrangian terms of the local agents, one launch (multi_agent_model.py:61-79, agent_solver.py:79-95)."""
    X_own, X_all, Y, Lam = _dev(X_own), _dev(X_all), _dev(Y), _dev(Lam)
    nl, _, K = X_own.shape
    N = X_all.shape[0]
    n_slots = tables.col_a.shape[1]
    if nbr_idx is not None:
        assert nbr_idx.is_cuda and nbr_idx.dtype == torch.int32 and nbr_idx.shape == (nl, n_slots) and nbr_idx.is_contiguous()
    if mask_in is not None:
        assert mask_in.is_cuda and mask_in.dtype == torch.uint8 and mask_in.shape == (nl, n_slots) and mask_in.is_contiguous()
    check(load().scvx_admm_round_prep(model_id, nl, int(i0), N, K, n_slots, float(d_min), float(rho), ptr(X_own), ptr(X_all), ptr(Y),
                                      ptr(Lam), ptr(nbr_idx), ptr(mask_in), ptr(tables.col_a), ptr(tables.col_b), ptr(tables.mask),
                                      ptr(tables.lin_p), ptr(tables.quad_rho), ptr(tables.aug_const), stream_ptr()),
          "scvx_admm_round_prep")
    return tables


def knn_select(d2, i0, k_sel, radius=None, out=None):
    """k_sel nearest neighbours per row of the squared-distance table d2 (n_local, N) -- d2 is used as scratch."""
    d2 = _dev(d2)
    nl, N = d2.shape
    if out is None:
        out = torch.empty((nl, k_sel), dtype=torch.int32, device=d2.device)
    check(load().scvx_knn_select(nl, int(i0), N, int(k_sel), float(radius) ** 2 if radius is not None else 0.0, ptr(d2), ptr(out),
                                 stream_ptr()), "scvx_knn_select")
    return out


def radius_mask(d2, radius, out=None):
    d2 = _dev(d2)
    nl, N = d2.shape
    if out is None:
        out = torch.empty((nl, N), dtype=torch.uint8, device=d2.device)
    check(load().scvx_radius_mask(nl, N, float(radius) ** 2, ptr(d2), ptr(out), stream_ptr()), "scvx_radius_mask")
    return out


def outer_update(model_id, M, conv_tol, X_new, U_new, nu_new, sigma_new, s_prime, X, U, sigma, tr_radius, active,
                 metrics):
    """One on-device outer-loop bookkeeping step (scvx_solver.py:82-111, :125-133); updates X, U, sigma,
    tr_radius, active in place and writes metrics (n, 6)."""
    n, _, K = X.shape
    check(load().scvx_outer_update(model_id, n, K, M, float(conv_tol), ptr(X_new), ptr(U_new), ptr(nu_new),
                                   ptr(sigma_new), ptr(s_prime) if M > 0 else None, ptr(X), ptr(U), ptr(sigma),
                                   ptr(tr_radius), ptr(active), ptr(metrics), stream_ptr()), "scvx_outer_update")


class SubproblemWorkspace:
    """Device scratch + output tensors for scvx_solve_batched, sized once for (n, K, M, n_nbr)."""

    def __init__(self, model_id, n, K, M, n_nbr, device, X_out=None):
        n_x, n_u, d = MODEL_DIMS[model_id]
        self.model_id, self.n, self.K, self.M, self.n_nbr = model_id, n, K, M, n_nbr
        nbytes = int(load().scvx_solve_workspace_bytes(model_id, n, K, M, n_nbr))
        self.scratch = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=device)
        self.nbytes = nbytes
        # X_out: the solver writes its states straight into a caller-owned buffer (BatchedADMM: the all-gather's send buffer)
        self.X = torch.empty((n, n_x, K), dtype=F64, device=device) if X_out is None else X_out
        assert self.X.shape == (n, n_x, K) and self.X.is_contiguous()
        self.U = torch.empty((n, n_u, K), dtype=F64, device=device)
        self.nu = torch.empty((n, n_x, K - 1), dtype=F64, device=device)
        self.sigma = torch.empty(n, dtype=F64, device=device)
        self.s_prime = torch.empty((n, M, K), dtype=F64, device=device)
        self.col_slack = torch.empty((n, n_nbr, K), dtype=F64, device=device) if n_nbr else None
        self.objective = torch.empty(n, dtype=F64, device=device)
        self.status = torch.empty(n, dtype=torch.int32, device=device)
        self.iters = torch.empty(n, dtype=torch.int32, device=device)
        self.retry_list = torch.zeros(n + 1, dtype=torch.int32, device=device)      # scratch of the retry pass


def solve_subproblem(ws: SubproblemWorkspace, mats, X_ref, U_ref, sigma_ref, tr_radius, x_init, x_final, pos_lo,
                     pos_hi, v_max, w_max, obs_a, obs_b, weight_nu, weight_slack, weight_sigma,
                     col_a=None, col_b=None, col_mask=None, quad_rho=None, lin_p=None, weight_col=1e5,
                     max_iter=0, norm1_induced=True, quad_diag=None, lin_w=None, quad_pair=None, fix_sigma=False,
                     block_order=None, mu0=None, active=None, retry_failed=False):
    """SCProblem.solve / AgentSolver.solve for a batch (sc_problem.py:15-105, agent_solver.py:43-117).
    Results land in the workspace's output tensors."""
    a = SolveArgs()
    a.model_id, a.n_agents, a.K, a.M, a.n_nbr = ws.model_id, ws.n, ws.K, ws.M, ws.n_nbr
    a.max_iter, a.norm1_induced = int(max_iter), 1 if norm1_induced else 0
    keep = []

    def P(t):
        if t is None:
            return None
        t = t.contiguous()
        keep.append(t)
        return ptr(t)

    A_bar, B_bar, C_bar, S_bar, z_bar = mats
    a.A_bar, a.B_bar, a.C_bar, a.S_bar, a.z_bar = P(A_bar), P(B_bar), P(C_bar), P(S_bar), P(z_bar)
    a.X_ref, a.U_ref, a.sigma_ref, a.tr_radius = P(X_ref), P(U_ref), P(sigma_ref), P(tr_radius)
    a.x_init, a.x_final, a.pos_lo, a.pos_hi, a.v_max, a.w_max = P(x_init), P(x_final), P(pos_lo), P(pos_hi), P(v_max), P(w_max)
    a.obs_a, a.obs_b = (P(obs_a), P(obs_b)) if ws.M else (None, None)
    a.col_a, a.col_b, a.col_mask = P(col_a), P(col_b), P(col_mask)
    a.quad_rho, a.lin_p = P(quad_rho), P(lin_p)
    a.quad_diag, a.lin_w, a.quad_pair, a.fix_sigma = P(quad_diag), P(lin_w), P(quad_pair), 1 if fix_sigma else 0
    a.block_order = P(block_order)
    a.mu0 = P(mu0)
    a.active = P(active)
    a.retry_failed = 1 if retry_failed else 0
    a.retry_list = ptr(ws.retry_list) if retry_failed else None
    a.weight_nu, a.weight_slack, a.weight_sigma, a.weight_col = float(weight_nu), float(weight_slack), float(weight_sigma), float(weight_col)
    a.X, a.U, a.nu, a.sigma = ptr(ws.X), ptr(ws.U), ptr(ws.nu), ptr(ws.sigma)
    a.s_prime = ptr(ws.s_prime) if ws.M else None
    a.col_slack = ptr(ws.col_slack) if ws.n_nbr else None
    a.objective, a.status, a.iters = ptr(ws.objective), ptr(ws.status), ptr(ws.iters)
    a.workspace, a.workspace_bytes = ptr(ws.scratch), ws.nbytes
    check(load().scvx_solve_batched(ctypes.byref(a), stream_ptr()), "scvx_solve_batched")
    return ws


def warm_start(model_id, p0, p1, obs_c, obs_r, clearance, K, obs_count=None):
    """Batched initial_guess (SCvx/utils/initial_guess.py:61-107, IS_initial_guess.py:87-126).

    p0, p1 (n, 3); obs_c (n, M_max, d); obs_r (n, M_max); obs_count (n,) int32 or None -> X0 (n, 3, K), U0 (n, n_u, K),
    status (n,) int32 (see include/scvx_b200.h)."""
    _, n_u, d = MODEL_DIMS[model_id]
    p0, p1 = _dev(p0), _dev(p1)
    n = p0.shape[0]
    assert p0.shape == (n, 3) and p1.shape == (n, 3)
    M = 0 if obs_c is None else obs_c.shape[1]
    if M:
        obs_c, obs_r = _dev(obs_c), _dev(obs_r)
        assert obs_c.shape == (n, M, d) and obs_r.shape == (n, M)
    if obs_count is not None:
        assert obs_count.is_cuda and obs_count.dtype == torch.int32 and obs_count.shape == (n,)
        obs_count = obs_count.contiguous()
    X0 = torch.empty((n, 3, K), dtype=F64, device=p0.device)
    U0 = torch.empty((n, n_u, K), dtype=F64, device=p0.device)
    status = torch.empty((n,), dtype=torch.int32, device=p0.device)
    check(load().scvx_warm_start_batched(model_id, n, K, M, ptr(p0), ptr(p1), ptr(obs_c) if M else None,
                                         ptr(obs_r) if M else None, ptr(obs_count), float(clearance), ptr(X0), ptr(U0),
                                         ptr(status), stream_ptr()), "scvx_warm_start_batched")
    return X0, U0, status


def min_inter_agent_distance(X, n_rows=3):
    """analysis.py:10-31 on a stacked (N, n_x, K) tensor -> (d_min (1,), d_mat (N, N))."""
    X = _dev(X)
    N, n_x, K = X.shape
    d_mat = torch.empty((N, N), dtype=F64, device=X.device)
    d_min = torch.empty((1,), dtype=F64, device=X.device)
    check(load().scvx_min_inter_agent_distance(N, K, n_x, n_rows, ptr(X), ptr(d_mat), ptr(d_min), stream_ptr()),
          "scvx_min_inter_agent_distance")
    return d_min, d_mat


def min_agent_obstacle_distance(X, obs_c, obs_r, robot_radius, n_rows=3):
    """analysis.py:34-62 on a stacked (N, n_x, K) tensor; obs_c (M, n_rows), obs_r (M,) -> (d_min (1,), d_mat (N, M))."""
    X = _dev(X)
    N, n_x, K = X.shape
    M = obs_c.shape[0]
    d_mat = torch.full((N, M), float("inf"), dtype=F64, device=X.device)
    d_min = torch.empty((1,), dtype=F64, device=X.device)
    if M:
        obs_c, obs_r = _dev(obs_c), _dev(obs_r)
        assert obs_c.shape == (M, n_rows) and obs_r.shape == (M,)
    check(load().scvx_min_agent_obstacle_distance(N, K, n_x, n_rows, M, ptr(X), ptr(obs_c) if M else None,
                                                  ptr(obs_r) if M else None, float(robot_radius), ptr(d_mat), ptr(d_min),
                                                  stream_ptr()), "scvx_min_agent_obstacle_distance")
    return d_min, d_mat


def slab_normals(model_id, P_own, X_dir, X_off, radius, i0=0, out=None):
    """Slab rows of the Nash best response (game_model.py:56-67,118-124) for local agents i0..i0+n_local-1.
    P_own (n_local, n_x, K); X_dir, X_off (n_agents, n_x, K); radius (n_local,) ->
    col_a (n_local, n_agents, d, K), col_b (n_local, n_agents, K), degenerate (n_local,) int32."""
    _, _, d = MODEL_DIMS[model_id]
    P_own, X_dir, X_off, radius = _dev(P_own), _dev(X_dir), _dev(X_off), _dev(radius)
    n_local, _, K = P_own.shape
    n_agents = X_dir.shape[0]
    assert X_off.shape == X_dir.shape and radius.shape == (n_local,)
    if out is None:
        out = (torch.empty((n_local, n_agents, d, K), dtype=F64, device=P_own.device),
               torch.empty((n_local, n_agents, K), dtype=F64, device=P_own.device))
    a, b = out
    deg = torch.empty((n_local,), dtype=torch.int32, device=P_own.device)
    check(load().scvx_slab_normals_batched(model_id, n_local, i0, n_agents, K, ptr(radius), ptr(P_own), ptr(X_dir),
                                           ptr(X_off), ptr(a), ptr(b), ptr(deg), stream_ptr()), "scvx_slab_normals_batched")
    return a, b, deg


def intersample(model_id, X, U, sigma, obs_c, obs_r, proj_dim=None, t_range=1.0, num_samples=100, eps=1e-4, tol=1e-6,
                max_roots=4):
    """find_critical_times + linearize_h (intersample_collision.py:29-97) for every (agent, segment, obstacle).
    X (n, n_x, K), U (n, n_u, K), sigma (n,), obs_c (n, M, m), obs_r (n, M) ->
    n_roots (n, K-1, M) int32, t_star, h0 (n, K-1, M, max_roots), grad_x (n, K-1, M, max_roots, n_x)."""
    n_x, n_u, _ = MODEL_DIMS[model_id]
    X, U, sigma, obs_c, obs_r = _dev(X), _dev(U), _dev(sigma), _dev(obs_c), _dev(obs_r)
    n, _, K = X.shape
    M = obs_c.shape[1]
    m = obs_c.shape[2] if proj_dim is None else proj_dim
    assert obs_c.shape == (n, M, m) and obs_r.shape == (n, M)
    dev = X.device
    n_roots = torch.zeros((n, K - 1, M), dtype=torch.int32, device=dev)
    t_star = torch.full((n, K - 1, M, max_roots), float("nan"), dtype=F64, device=dev)
    h0 = torch.full((n, K - 1, M, max_roots), float("nan"), dtype=F64, device=dev)
    grad_x = torch.full((n, K - 1, M, max_roots, n_x), float("nan"), dtype=F64, device=dev)
    check(load().scvx_intersample_batched(model_id, n, K, M, m, ptr(X), ptr(U), ptr(sigma), ptr(obs_c), ptr(obs_r),
                                          float(t_range), int(num_samples), float(eps), float(tol), int(max_roots),
                                          ptr(n_roots), ptr(t_star), ptr(h0), ptr(grad_x), stream_ptr()),
          "scvx_intersample_batched")
    return n_roots, t_star, h0, grad_x


def clearance_samples(model_id, X, U, sigma, obs_c, total_r, resolution=50):
    """h at t = i/resolution of every segment, one obstacle per agent: (n, K-1, resolution)."""
    X, U, sigma, obs_c, total_r = _dev(X), _dev(U), _dev(sigma), _dev(obs_c), _dev(total_r)
    n, _, K = X.shape
    m = obs_c.shape[1]
    out = torch.empty((n, K - 1, resolution), dtype=F64, device=X.device)
    check(load().scvx_clearance_samples_batched(model_id, n, K, m, int(resolution), ptr(X), ptr(U), ptr(sigma), ptr(obs_c),
                                                ptr(total_r), ptr(out), stream_ptr()), "scvx_clearance_samples_batched")
    return out


def cross_min_dist2(model_id, X_own, X_all, out=None):
    """min over k of the squared position distance between every local agent and every agent: (n_local, n_agents)."""
    X_own, X_all = _dev(X_own), _dev(X_all)
    nl, _, K = X_own.shape
    N = X_all.shape[0]
    d2 = torch.empty((nl, N), dtype=F64, device=X_own.device) if out is None else out
    check(load().scvx_cross_min_dist2(model_id, nl, N, K, ptr(X_own), ptr(X_all), ptr(d2), stream_ptr()), "scvx_cross_min_dist2")
    return d2


def linearize_collision_indexed(model_id, X_own, X_all, nbr_idx, d_min, out=None):
    """Inter-agent half-spaces for per-agent neighbour lists nbr_idx (n_local, n_sel) int32 (-1 = empty):
    col_a (n_local, n_sel, d, K), col_b (n_local, n_sel, K)."""
    _, _, d = MODEL_DIMS[model_id]
    X_own, X_all = _dev(X_own), _dev(X_all)
    nl, _, K = X_own.shape
    assert nbr_idx.is_cuda and nbr_idx.dtype == torch.int32 and nbr_idx.shape[0] == nl
    nbr_idx = nbr_idx.contiguous()
    n_sel = nbr_idx.shape[1]
    if out is None:
        out = (torch.empty((nl, n_sel, d, K), dtype=F64, device=X_own.device), torch.empty((nl, n_sel, K), dtype=F64, device=X_own.device))
    a, b = out
    check(load().scvx_linearize_collision_indexed(model_id, nl, n_sel, X_all.shape[0], K, float(d_min), ptr(X_own), ptr(X_all),
                                                  ptr(nbr_idx), ptr(a), ptr(b), stream_ptr()), "scvx_linearize_collision_indexed")
    return out


def order_by_iters(iters, out=None):
    """Longest-first launch order from a solve's iteration counts (int32 (n,) -> int32 (n,) permutation)."""
    assert iters.is_cuda and iters.dtype == torch.int32
    iters = iters.contiguous()
    if out is None:
        out = torch.empty_like(iters)
    check(load().scvx_order_by_iters(iters.numel(), ptr(iters), ptr(out), stream_ptr()), "scvx_order_by_iters")
    return out


def mu0_from_iters(iters, out=None, easy_max_iters=10, mu0_easy=0.1, mu0_hard=10.0):
    """Per-agent barrier start of the next solve from this solve's iteration counts (int32 (n,) -> float64 (n,))."""
    assert iters.is_cuda and iters.dtype == torch.int32
    iters = iters.contiguous()
    if out is None:
        out = torch.empty(iters.numel(), dtype=F64, device=iters.device)
    check(load().scvx_mu0_from_iters(iters.numel(), ptr(iters), int(easy_max_iters), float(mu0_easy), float(mu0_hard), ptr(out),
                                     stream_ptr()), "scvx_mu0_from_iters")
    return out
