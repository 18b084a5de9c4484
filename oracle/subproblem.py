"""Oracle restatement of the reference's convex sub-problem (stage 3) as an explicit LP / QP.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED against cvxpy+ECOS / CLARABEL: the
packages are absent from the build container AND from the GPU box (tools/probe_solvers.py, log
profiles/r02_probe_solvers_gpu_box.json), upstream pins no versions and holds no golden values for
this stage; exact solver = HiGHS; compare optimal VALUES, not minimisers.

Follows (reference paths):
  SCvx/optimization/sc_problem.py:21-26    variables X, U, nu, sigma>=0
  SCvx/optimization/sc_problem.py:53-68    dynamics rows (cvx.reshape default order 'F')
  SCvx/optimization/sc_problem.py:71-74    trust region  norm(dX,1)+norm(dU,1)+|ds| <= r
  SCvx/optimization/sc_problem.py:77-82    objective  w_nu*norm(nu,1) + w_slack*sum(s') + w_sigma*sigma
  SCvx/models/unicycle_model.py:85-115     boundary / input / box / obstacle rows (unicycle)
  SCvx/models/single_integrator_model.py:79-128   same, ||u_k||_2 <= v_max (SOC) for the single integrator
  SCvx/optimization/agent_solver.py:79-102 inter-agent rows + augmented Lagrangian (ADMM variant)

cvxpy semantics honoured: ``cvx.norm(M, 1)`` on a 2-D expression is the INDUCED matrix 1-norm
(max column abs-sum) -- ``norm1_mode="induced"`` (default).  ``"entrywise"`` is kept as a switch
(SURVEY fact 4).  ``np.linalg.norm(nu, 1)`` in scvx_solver.py:82 is the same induced norm.

The SOC rows of the single integrator are handled by Kelley cutting planes on top of the exact
LP/QP solver (outer approximation tightened until every ||u_k||_2 <= v_max + 1e-9), which converges
to the SOCP optimum.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import highs
from .models import linearize_obstacles

WEIGHT_COLLISION_SLACK = 1e5  # SCvx/optimization/admm_utils.py:5


class Params:
    """All numeric inputs of one sub-problem (what SCProblem.set_parameters / AgentSolver.setup bind)."""

    def __init__(self, model, K, mats, X_ref, U_ref, sigma_ref, tr_radius,
                 weight_nu=1e4, weight_slack=1e6, weight_sigma=100.0, norm1_mode="induced",
                 neighbors=None, rho=1.0, d_min=0.0, weight_col=WEIGHT_COLLISION_SLACK):
        self.model, self.K = model, K
        self.A_bar, self.B_bar, self.C_bar, self.S_bar, self.z_bar = [np.asarray(a, float) for a in mats]
        self.X_ref, self.U_ref, self.sigma_ref = np.asarray(X_ref, float), np.asarray(U_ref, float), float(sigma_ref)
        self.tr_radius = float(tr_radius)
        self.weight_nu, self.weight_slack, self.weight_sigma = float(weight_nu), float(weight_slack), float(weight_sigma)
        self.norm1_mode = norm1_mode
        # neighbours: list of dicts {a: (d,K), Y: (d,K), Lam: (d,K)}  (agent_solver.py:82-95)
        self.neighbors = neighbors or []
        self.rho, self.d_min, self.weight_col = float(rho), float(d_min), float(weight_col)
        self.obs_a, self.obs_rhs, self.obs_c = linearize_obstacles(model, self.X_ref)
        # Nash best response (agent_best_response.py:36-100, game_model.py:69-126); see set_game()
        self.game = None

    def set_game(self, quad_diag=None, lin_w=None, quad_pair=None, hard_rows=(), fix_sigma=True, const=0.0):
        """Extra terms of AgentBestResponse: with w_k = (x_k, u_k),
        sum_k sum_i [quad_diag_i/2 w_ik^2 + lin_w_ik w_ik] + sum_{k<K-1} sum_i quad_pair_i/2 (w_i,k+1 - w_ik)^2 + const,
        hard slab rows a_k.p_k >= b_k (game_model.py:118-124) and sigma == sigma_ref (agent_best_response.py:76-77)."""
        ns = self.model.n_x + self.model.n_u
        z = lambda v, shape: np.zeros(shape) if v is None else np.asarray(v, float).reshape(shape)   # noqa: E731
        self.game = {"qd": z(quad_diag, (ns,)), "lw": z(lin_w, (ns, self.K)), "qp": z(quad_pair, (ns,)),
                     "rows": [(np.asarray(a, float), np.asarray(b, float)) for a, b in hard_rows],
                     "fix_sigma": bool(fix_sigma), "const": float(const)}
        return self

    def game_cost(self, X, U):
        """(value, gradient wrt W = vstack(X, U)) of the smooth extra cost."""
        g = self.game
        W = np.vstack([X, U])
        val = 0.5 * (g["qd"][:, None] * W * W).sum() + (g["lw"] * W).sum() + g["const"]
        grad = g["qd"][:, None] * W + g["lw"]
        dW = W[:, 1:] - W[:, :-1]
        val += 0.5 * (g["qp"][:, None] * dW * dW).sum()
        grad[:, 1:] += g["qp"][:, None] * dW
        grad[:, :-1] -= g["qp"][:, None] * dW
        return float(val), grad


# ---------------------------------------------------------------------------------------------
def defects(p: Params, X, U, sigma):
    """nu_k = x_{k+1} - (A_k x_k + B_k u_k + C_k u_{k+1} + S_k sigma + z_k)  (sc_problem.py:60-68)."""
    n_x, n_u, K = p.model.n_x, p.model.n_u, p.K
    nu = np.zeros((n_x, K - 1))
    for k in range(K - 1):
        A = p.A_bar[:, k].reshape((n_x, n_x), order="F")
        B = p.B_bar[:, k].reshape((n_x, n_u), order="F")
        C = p.C_bar[:, k].reshape((n_x, n_u), order="F")
        nu[:, k] = X[:, k + 1] - (A @ X[:, k] + B @ U[:, k] + C @ U[:, k + 1] + p.S_bar[:, k] * sigma + p.z_bar[:, k])
    return nu


def norm1(M, mode):
    M = np.atleast_2d(M)
    return np.abs(M).sum(axis=0).max() if mode == "induced" else np.abs(M).sum()


def evaluate(p: Params, X, U, sigma):
    """Objective and hard-constraint violation of a candidate (X, U, sigma) with nu, s', S_j
    eliminated at their optimal values (nu = defect, slacks = hinge).  Used to score ANY solver's
    output against the oracle's optimal value."""
    m, K, d = p.model, p.K, p.model.d
    nu = defects(p, X, U, sigma)
    P = X[0:d, :]
    sprime = np.zeros((len(m.obstacles), K))
    for j in range(len(m.obstacles)):
        lhs = np.einsum("dk,dk->k", p.obs_a[j], P - p.obs_c[j][:, None])
        sprime[j] = np.maximum(0.0, p.obs_rhs[j] - lhs)
    obj = p.weight_nu * norm1(nu, p.norm1_mode) + p.weight_slack * sprime.sum() + p.weight_sigma * sigma
    S = []
    for nb in p.neighbors:
        lhs = np.einsum("dk,dk->k", nb["a"], P - nb["Y"])
        Sj = np.maximum(0.0, p.d_min - lhs)
        S.append(Sj)
        diff = P - nb["Y"]
        obj += (nb["Lam"] * diff).sum() + 0.5 * p.rho * (diff ** 2).sum() + p.weight_col * Sj.sum()
    # hard constraints
    viol = 0.0
    viol = max(viol, np.abs(X[:, 0] - m.x_init).max(), np.abs(X[:, -1] - m.x_final).max())
    viol = max(viol, np.abs(U[:, 0]).max(), np.abs(U[:, -1]).max())
    if m.kind == "unicycle":
        viol = max(viol, (-U[0]).max(), (U[0] - m.v_max).max(), (np.abs(U[1]) - m.w_max).max())
    else:
        viol = max(viol, (np.linalg.norm(U, axis=0) - m.v_max).max())
    viol = max(viol, (P - (m.upper_bound - m.robot_radius)).max(), ((m.lower_bound + m.robot_radius) - P).max())
    tr = norm1(X - p.X_ref, p.norm1_mode) + norm1(U - p.U_ref, p.norm1_mode) + abs(sigma - p.sigma_ref)
    viol = max(viol, tr - p.tr_radius, -sigma, 0.0)
    if p.game is not None:
        obj += p.game_cost(X, U)[0]
        for a_h, b_h in p.game["rows"]:
            viol = max(viol, (b_h - np.einsum("dk,dk->k", a_h, P)).max())
        if p.game["fix_sigma"]:
            viol = max(viol, abs(sigma - p.sigma_ref))
    return {"obj": float(obj), "viol": float(viol), "nu": nu, "s_prime": sprime, "S": S,
            "nu_norm": float(norm1(nu, "induced")), "slack_sum": float(sprime.sum()), "tr_lhs": float(tr)}


# ---------------------------------------------------------------------------------------------
class _Builder:
    def __init__(self):
        self.n = 0
        self.rows, self.cols, self.vals = [], [], []
        self.rlo, self.rhi = [], []
        self.m = 0

    def var(self, count):
        s = self.n
        self.n += count
        return s

    def row(self, idx, val, lo, hi):
        for i, v in zip(idx, val):
            self.rows.append(self.m); self.cols.append(int(i)); self.vals.append(float(v))
        self.rlo.append(lo); self.rhi.append(hi)
        self.m += 1


def solve(p: Params, solver="choose", max_cuts=60, linearize_at=None):
    """Solve the sub-problem exactly.  Returns dict(X, U, nu, sigma, s_prime, S, obj, status, ok).

    linearize_at = P0 (d, K): replace the quadratic term rho/2 * NB * |P|^2 of the ADMM variant by its
    first-order model about P0.  The resulting LP's optimal value is a LOWER bound on the QP optimum
    (convexity), and f(z0) - LB is the Frank-Wolfe gap of z0 -- see `qp_bracket`."""
    m, K = p.model, p.K
    n_x, n_u, d = m.n_x, m.n_u, m.d
    M, NB = len(m.obstacles), len(p.neighbors)
    INF = np.inf
    b = _Builder()
    iX = b.var(n_x * K); iU = b.var(n_u * K); iNu = b.var(n_x * (K - 1)); iSig = b.var(1)
    iSp = b.var(M * K); iS = b.var(NB * K)
    iEnu = b.var(n_x * (K - 1)); iEx = b.var(n_x * K); iEu = b.var(n_u * K); iEs = b.var(1)
    iTnu = b.var(1); iTx = b.var(1); iTu = b.var(1)
    n = b.n
    X = lambda i, k: iX + n_x * k + i      # noqa: E731
    Uv = lambda i, k: iU + n_u * k + i     # noqa: E731
    Nu = lambda i, k: iNu + n_x * k + i    # noqa: E731

    lo = np.full(n, -INF); hi = np.full(n, INF); c = np.zeros(n)
    # boundary conditions (unicycle_model.py:94)
    for i in range(n_x):
        lo[X(i, 0)] = hi[X(i, 0)] = m.x_init[i]
        lo[X(i, K - 1)] = hi[X(i, K - 1)] = m.x_final[i]
    # position box (unicycle_model.py:98-101); intersect with boundary equalities
    for k in range(1, K - 1):
        for i in range(d):
            lo[X(i, k)] = m.lower_bound + m.robot_radius
            hi[X(i, k)] = m.upper_bound - m.robot_radius
    box_rows_needed = []
    for k in (0, K - 1):   # box also applies at the end nodes: keep as rows so infeasibility shows up
        for i in range(d):
            box_rows_needed.append((X(i, k), m.lower_bound + m.robot_radius, m.upper_bound - m.robot_radius))
    # inputs
    if m.kind == "unicycle":        # unicycle_model.py:96
        for k in range(K):
            lo[Uv(0, k)], hi[Uv(0, k)] = 0.0, m.v_max
            lo[Uv(1, k)], hi[Uv(1, k)] = -m.w_max, m.w_max
    else:                            # SOC: start from the bounding box, tightened by cuts below
        for k in range(K):
            for i in range(n_u):
                lo[Uv(i, k)], hi[Uv(i, k)] = -m.v_max, m.v_max
    for i in range(n_u):
        lo[Uv(i, 0)] = hi[Uv(i, 0)] = 0.0
        lo[Uv(i, K - 1)] = hi[Uv(i, K - 1)] = 0.0
    lo[iSig] = 0.0                                        # sc_problem.py:26
    lo[iSp:iSp + M * K] = 0.0                             # unicycle_model.py:51 nonneg
    lo[iS:iS + NB * K] = 0.0                              # agent_solver.py:41 nonneg
    for s in (iEnu, iEx, iEu):
        pass
    lo[iEnu:iTu + 1] = 0.0

    for (col, l, h) in box_rows_needed:
        b.row([col], [1.0], l, h)
    # obstacle rows  a^T p_k + s' >= rhs + a^T c
    for j in range(M):
        for k in range(K):
            a = p.obs_a[j][:, k]
            b.row([X(i, k) for i in range(d)] + [iSp + j * K + k], list(a) + [1.0],
                  p.obs_rhs[j] + a.dot(p.obs_c[j]), INF)
    # inter-agent rows  a^T (p_k - Y_k) + S >= d_min   (agent_solver.py:85-90)
    for q, nb in enumerate(p.neighbors):
        for k in range(K):
            a = nb["a"][:, k]
            b.row([X(i, k) for i in range(d)] + [iS + q * K + k], list(a) + [1.0],
                  p.d_min + a.dot(nb["Y"][:, k]), INF)
    # hard slab rows of the Nash best response  z.(p_k - Y_k) >= r  <=>  a.p_k >= b   (game_model.py:118-124)
    if p.game is not None:
        for a_h, b_h in p.game["rows"]:
            for k in range(K):
                b.row([X(i, k) for i in range(d)], list(a_h[:, k]), float(b_h[k]), INF)
        if p.game["fix_sigma"]:
            lo[iSig] = hi[iSig] = p.sigma_ref
    # dynamics rows (sc_problem.py:53-68)
    for k in range(K - 1):
        A = p.A_bar[:, k].reshape((n_x, n_x), order="F")
        B = p.B_bar[:, k].reshape((n_x, n_u), order="F")
        C = p.C_bar[:, k].reshape((n_x, n_u), order="F")
        for i in range(n_x):
            idx = [X(i, k + 1)] + [X(j, k) for j in range(n_x)] + [Uv(j, k) for j in range(n_u)] \
                + [Uv(j, k + 1) for j in range(n_u)] + [iSig, Nu(i, k)]
            val = [1.0] + list(-A[i]) + list(-B[i]) + list(-C[i]) + [-p.S_bar[i, k], -1.0]
            # x_{k+1}[i] may coincide with no other index; duplicates are summed by the COO->CSC conversion
            b.row(idx, val, p.z_bar[i, k], p.z_bar[i, k])
    # |nu| epigraph
    for k in range(K - 1):
        for i in range(n_x):
            e = iEnu + n_x * k + i
            b.row([Nu(i, k), e], [1.0, -1.0], -INF, 0.0)
            b.row([Nu(i, k), e], [-1.0, -1.0], -INF, 0.0)
    # |dX|, |dU|, |dsigma| epigraphs
    for k in range(K):
        for i in range(n_x):
            e = iEx + n_x * k + i
            b.row([X(i, k), e], [1.0, -1.0], -INF, p.X_ref[i, k])
            b.row([X(i, k), e], [-1.0, -1.0], -INF, -p.X_ref[i, k])
        for i in range(n_u):
            e = iEu + n_u * k + i
            b.row([Uv(i, k), e], [1.0, -1.0], -INF, p.U_ref[i, k])
            b.row([Uv(i, k), e], [-1.0, -1.0], -INF, -p.U_ref[i, k])
    b.row([iSig, iEs], [1.0, -1.0], -INF, p.sigma_ref)
    b.row([iSig, iEs], [-1.0, -1.0], -INF, -p.sigma_ref)
    if p.norm1_mode == "induced":
        for k in range(K - 1):
            b.row([iEnu + n_x * k + i for i in range(n_x)] + [iTnu], [1.0] * n_x + [-1.0], -INF, 0.0)
        for k in range(K):
            b.row([iEx + n_x * k + i for i in range(n_x)] + [iTx], [1.0] * n_x + [-1.0], -INF, 0.0)
            b.row([iEu + n_u * k + i for i in range(n_u)] + [iTu], [1.0] * n_u + [-1.0], -INF, 0.0)
    else:
        b.row(list(range(iEnu, iEnu + n_x * (K - 1))) + [iTnu], [1.0] * (n_x * (K - 1)) + [-1.0], -INF, 0.0)
        b.row(list(range(iEx, iEx + n_x * K)) + [iTx], [1.0] * (n_x * K) + [-1.0], -INF, 0.0)
        b.row(list(range(iEu, iEu + n_u * K)) + [iTu], [1.0] * (n_u * K) + [-1.0], -INF, 0.0)
    b.row([iTx, iTu, iEs], [1.0, 1.0, 1.0], -INF, p.tr_radius)          # sc_problem.py:74

    # objective (sc_problem.py:77-82)
    c[iTnu] = p.weight_nu
    c[iSp:iSp + M * K] = p.weight_slack
    c[iSig] = p.weight_sigma
    Q = None
    offset = 0.0
    if NB:
        c[iS:iS + NB * K] = p.weight_col
        qdiag = np.zeros(n)
        for nb in p.neighbors:       # <Lam, P - Y> + rho/2 ||P - Y||^2   (agent_solver.py:92-94)
            for k in range(K):
                for i in range(d):
                    c[X(i, k)] += nb["Lam"][i, k] - p.rho * nb["Y"][i, k]
                    qdiag[X(i, k)] += p.rho
            offset += -(nb["Lam"] * nb["Y"]).sum() + 0.5 * p.rho * (nb["Y"] ** 2).sum()
        if linearize_at is None:
            Q = sp.diags(qdiag)
        else:
            P0 = np.asarray(linearize_at, float)
            for k in range(K):
                for i in range(d):
                    c[X(i, k)] += p.rho * NB * P0[i, k]
            offset -= 0.5 * p.rho * NB * (P0 ** 2).sum()

    if p.game is not None:
        g = p.game
        if linearize_at is None:
            raise NotImplementedError("the Nash best response is certified through qp_bracket (LP lower bound)")
        X0, U0 = linearize_at
        val0, grad0 = p.game_cost(np.asarray(X0, float), np.asarray(U0, float))
        W0 = np.vstack([X0, U0])
        for k in range(K):
            for i in range(n_x):
                c[X(i, k)] += grad0[i, k]
            for i in range(n_u):
                c[Uv(i, k)] += grad0[n_x + i, k]
        offset += val0 - float((grad0 * W0).sum())
    rows, cols, vals = list(b.rows), list(b.cols), list(b.vals)
    rlo, rhi = list(b.rlo), list(b.rhi)
    mrows = b.m
    cuts = 0
    while True:
        Amat = sp.coo_matrix((vals, (rows, cols)), shape=(mrows, n)).tocsc()
        Amat.sum_duplicates()
        r = highs.solve(c, Amat, np.array(rlo), np.array(rhi), lo, hi, Q=Q, offset=offset, solver=solver)
        if not r["ok"] or m.kind == "unicycle":
            break
        x = r["x"]
        Uval = x[iU:iU + n_u * K].reshape((K, n_u)).T
        nrm = np.linalg.norm(Uval, axis=0)
        bad = np.where(nrm > m.v_max + 1e-9)[0]
        if bad.size == 0 or cuts >= max_cuts:
            break
        for k in bad:               # tangent cut  g^T u_k <= v_max, g = u_k/||u_k||
            g = Uval[:, k] / nrm[k]
            for i in range(n_u):
                rows.append(mrows); cols.append(Uv(i, k)); vals.append(float(g[i]))
            rlo.append(-INF); rhi.append(m.v_max); mrows += 1
        cuts += 1
    x = r["x"]
    out = {"status": r["status"], "ok": r["ok"], "obj": r["obj"], "cuts": cuts}
    out["X"] = x[iX:iX + n_x * K].reshape((K, n_x)).T.copy()
    out["U"] = x[iU:iU + n_u * K].reshape((K, n_u)).T.copy()
    out["nu"] = x[iNu:iNu + n_x * (K - 1)].reshape((K - 1, n_x)).T.copy()
    out["sigma"] = float(x[iSig])
    out["s_prime"] = x[iSp:iSp + M * K].reshape((M, K)).copy()
    out["S"] = [x[iS + q * K: iS + (q + 1) * K].copy() for q in range(NB)]
    return out


def qp_bracket(p: Params, X, U, sigma):
    """Exact optimality certificate for a candidate solution of the QP variant (agent_solver.py:79-102)
    using only the exact LP solver: returns (f(z0), LB) with LB <= f_opt <= f(z0), where LB is the optimum
    of the LP obtained by linearising the quadratic term about z0 (Frank-Wolfe gap).  HiGHS' own QP
    active-set solver does not finish on these problems (>10 min at K=12), so QP parity is pinned this way."""
    f0 = evaluate(p, X, U, sigma)
    lb = solve(p, linearize_at=(X, U) if p.game is not None else X[0:p.model.d, :])
    return f0["obj"], lb["obj"], f0["viol"], lb["ok"]
