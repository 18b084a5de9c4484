"""Oracle restatement of the Distributed_opt research scripts -- TEST INFRASTRUCTURE ONLY.

Follows (reference paths):
  Distributed_opt/ADMM_decentralized.py:14-29    descete_f   (exact ZOH of the 2-D double integrator, scipy.signal)
  Distributed_opt/ADMM_decentralized.py:32-170   x_traj_opt  (5 ADMM iterations: per-robot QP, per-robot s_bar QP, dual ascent)
  Distributed_opt/ADMM_decentralized.py:174-180  x_initial   (np.linspace between x_ini and x_des)
  Distributed_opt/dist_scvx_3d.py:9-28, :31-118, :131-138    3-D twin: one QP per robot with the collision rows inside, cost_fcn

PARITY UNPINNED: the scripts cannot be imported here (cvxpy, jax and matplotlib are imported at module top and are not
installed; the module bodies also build their scenario in globals).  The QPs are restated line by line and solved with the
exact HiGHS QP solver (which does finish on these strictly convex problems, ~0.3 s each).

Shared per-robot QP data (`RobotQP`): what one `cp.Problem` of the scripts contains, as numbers.
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.sparse as sp
from scipy import signal

from . import highs


def descete_f(dt, n, m):
    """ADMM_decentralized.py:14-29 (n=4, m=2) / dist_scvx_3d.py:9-28 (n=6, m=3): double integrator, exact ZOH."""
    d = n // 2
    A = np.zeros((n, n)); A[:d, d:] = np.eye(d)
    B = np.zeros((n, m)); B[d:, :] = np.eye(d)
    sysd = signal.StateSpace(A, B, np.eye(n), np.zeros((n, m))).to_discrete(dt)
    return [sysd.A, sysd.B]


def x_initial(x_ini, x_des, T):
    """ADMM_decentralized.py:174-180: straight line in the full (state, control) vector."""
    return {name: np.linspace(x_ini[name], x_des[name], T) for name in x_ini}


class RobotQP:
    """min  c_w sum_{t<T-1} |u_t + w_t|^2 + sum_t [ lin_t . dpos_t + rho/2 |dpos_t - sbar_t|^2 ] + c_S sum_t S_t
    s.t. d_0 = 0, d_{T-1} = x_des - x_{T-1}, x_{t+1} + d_{t+1} = A(x_t + d_t) + B(u_t + w_t),
         |w_t|_1 <= tr, box_lo <= (x_t + d_t)[0:2] <= box_hi            (t = 0..T-2)
         h_tq - g_tq . d_t[0:dc] <= S_t, S_t >= 0                         (t = 0..T-2, q over the other robots)
    (ADMM_decentralized.py:52-98 with c_w=100, dc=0;  dist_scvx_3d.py:51-111 with c_w=1, rho=0, c_S=1e4, dc=3)."""

    def __init__(self, Ad, Bd, x, u, x_des, tr, c_w, box=((-1.0, 22.0), (-1.0, 20.0)), rho=0.0, lin=None, sbar=None,
                 col_h=None, col_g=None, c_S=0.0):
        self.Ad, self.Bd = np.asarray(Ad, float), np.asarray(Bd, float)
        self.x, self.u = np.asarray(x, float), np.asarray(u, float)          # (T, n), (T-1, m)
        self.T, self.n = self.x.shape
        self.m = self.u.shape[1]
        self.x_des, self.tr, self.c_w, self.box = np.asarray(x_des, float), float(tr), float(c_w), box
        self.rho = float(rho)
        self.lin = np.zeros((self.T, 2)) if lin is None else np.asarray(lin, float)      # (T, 2) linear cost on dpos
        self.sbar = np.zeros((self.T, 2)) if sbar is None else np.asarray(sbar, float)
        self.col_h = np.zeros((self.T - 1, 0)) if col_h is None else np.asarray(col_h, float)          # (T-1, nq)
        self.col_g = np.zeros((self.T - 1, 0, 0)) if col_g is None else np.asarray(col_g, float)       # (T-1, nq, dc)
        self.c_S = float(c_S)
        self.c = np.array([self.Ad @ self.x[t] + self.Bd @ self.u[t] - self.x[t + 1] for t in range(self.T - 1)])   # defects

    def objective(self, d, w):
        T = self.T
        val = self.c_w * ((self.u + w[:T - 1]) ** 2).sum()
        dp = d[:, :2]
        val += (self.lin * dp).sum() + 0.5 * self.rho * ((dp - self.sbar) ** 2).sum()
        if self.col_h.shape[1]:
            dc = self.col_g.shape[2]
            S = np.maximum(0.0, (self.col_h - np.einsum("tqc,tc->tq", self.col_g, d[:T - 1, :dc])).max(axis=1))
            val += self.c_S * S.sum()
        return float(val)

    def violation(self, d, w):
        T, n = self.T, self.n
        v = max(np.abs(d[0]).max(), np.abs(d[T - 1] + self.x[T - 1] - self.x_des[:n]).max())
        for t in range(T - 1):
            v = max(v, np.abs(d[t + 1] - (self.Ad @ d[t] + self.Bd @ w[t] + self.c[t])).max())
            v = max(v, np.abs(w[t]).sum() - self.tr)
            for i, (lo, hi) in enumerate(self.box):
                v = max(v, (self.x[t, i] + d[t, i]) - hi, lo - (self.x[t, i] + d[t, i]))
        return float(max(v, 0.0))


def solve_robot_qp(q: RobotQP, time_limit=30.0, linearize_at=None):
    """Exact solve with HiGHS.  Returns dict(d (T,n), w (T,m), S (T,), obj, ok).
    linearize_at=(d0, w0): replace the quadratic terms by their first-order model about (d0, w0); the LP optimum is then a
    LOWER bound of the QP optimum (Frank-Wolfe gap) -- used by `qp_bracket` when HiGHS' QP solver runs out of time."""
    T, n, m = q.T, q.n, q.m
    nq = q.col_h.shape[1]
    nv = T * (n + m) + T
    D = lambda t, j: t * (n + m) + j            # noqa: E731
    W = lambda t, j: t * (n + m) + n + j        # noqa: E731
    Sv = lambda t: T * (n + m) + t              # noqa: E731
    rows, cols, vals, lo, hi = [], [], [], [], []
    mr = 0

    def row(idx, val, l, h):
        nonlocal mr
        for a, b in zip(idx, val):
            rows.append(mr); cols.append(a); vals.append(b)
        lo.append(l); hi.append(h); mr += 1

    for j in range(n):
        row([D(0, j)], [1.0], 0.0, 0.0)
        e = q.x_des[j] - q.x[T - 1, j]
        row([D(T - 1, j)], [1.0], e, e)
    signs = list(itertools.product([1.0, -1.0], repeat=m))
    for t in range(T - 1):
        for j in range(n):
            row([D(t + 1, j)] + [D(t, k) for k in range(n)] + [W(t, k) for k in range(m)],
                [1.0] + list(-q.Ad[j]) + list(-q.Bd[j]), q.c[t, j], q.c[t, j])
        for sg in signs:
            row([W(t, k) for k in range(m)], list(sg), -np.inf, q.tr)
        for i, (blo, bhi) in enumerate(q.box):
            row([D(t, i)], [1.0], blo - q.x[t, i], bhi - q.x[t, i])
        for k in range(nq):
            dc = q.col_g.shape[2]
            row([D(t, c) for c in range(dc)] + [Sv(t)], list(-q.col_g[t, k]) + [-1.0], -np.inf, -q.col_h[t, k])
    cvec = np.zeros(nv); qd = np.zeros(nv)
    vlo = np.full(nv, -np.inf); vhi = np.full(nv, np.inf)
    for t in range(T - 1):
        for k in range(m):
            qd[W(t, k)] = 2.0 * q.c_w; cvec[W(t, k)] = 2.0 * q.c_w * q.u[t, k]
    for k in range(m):
        vlo[W(T - 1, k)] = vhi[W(T - 1, k)] = 0.0       # unconstrained & costless in the scripts; pinned to 0 here
    for t in range(T):
        for k in range(2):
            qd[D(t, k)] += q.rho; cvec[D(t, k)] += q.lin[t, k] - q.rho * q.sbar[t, k]
        vlo[Sv(t)] = 0.0; cvec[Sv(t)] = q.c_S
        if not nq or t == T - 1:
            vhi[Sv(t)] = 0.0 if not nq or t == T - 1 else np.inf
    offset = q.c_w * (q.u ** 2).sum() + 0.5 * q.rho * (q.sbar ** 2).sum()
    if linearize_at is not None:
        d0, w0 = linearize_at
        for t in range(T - 1):
            for k in range(m):
                cvec[W(t, k)] += 2.0 * q.c_w * w0[t, k]          # grad of c_w |u+w|^2 at w0 is 2 c_w (u + w0)
        offset -= q.c_w * (w0[:T - 1] ** 2).sum()
        for t in range(T):
            for k in range(2):
                cvec[D(t, k)] += q.rho * d0[t, k]
        offset -= 0.5 * q.rho * (d0[:, :2] ** 2).sum()
        qd[:] = 0.0
    A = sp.coo_matrix((vals, (rows, cols)), shape=(mr, nv)).tocsc()
    r = highs.solve(cvec, A, np.array(lo), np.array(hi), vlo, vhi, Q=(sp.diags(qd) if linearize_at is None else None), offset=offset,
                    time_limit=time_limit)
    z = r["x"]
    s = z[:T * (n + m)].reshape(T, n + m)
    return {"d": s[:, :n].copy(), "w": s[:, n:].copy(), "S": z[T * (n + m):].copy(), "obj": r["obj"], "ok": r["ok"],
            "status": r["status"]}


def solve_robot_qp_certified(q: RobotQP, time_limit=20.0):
    """HiGHS' QP solver reports 'Solve error' / runs out of time on some of these problems (saturated trust regions).
    Fall back to the numpy twin of the GPU algorithm and CERTIFY its point with the exact LP bracket."""
    r = solve_robot_qp(q, time_limit=time_limit)
    if r["ok"]:
        return r
    from .lti_ipm import LtiIPM
    s = LtiIPM(q).solve()
    f0, lb, viol, ok = qp_bracket(q, s["d"], s["w"])
    if not ok or viol > 1e-8 or f0 - lb > 1e-6 * max(abs(f0), 1.0):
        raise RuntimeError(f"QP could not be certified: HiGHS {r['status']}, bracket [{lb}, {f0}], violation {viol}")
    return {"d": s["d"], "w": s["w"], "S": None, "obj": f0, "ok": True, "status": "certified by LP bracket"}


def qp_bracket(q: RobotQP, d, w):
    """(f(z), LB, violation): LB <= f_opt <= f(z) from ONE exact LP (linearisation of the quadratic about z)."""
    lb = solve_robot_qp(q, linearize_at=(d, w))
    return q.objective(d, w), lb["obj"], q.violation(d, w), lb["ok"]


# ---------------------------------------------------------------------------------------------------------------
def collision_rows(X_traj, names, name, R, dc, T):
    """h_tq = 2R - |p_i,t - p_q,t|, g_tq = (p_i,t - p_q,t)/|.|  (ADMM_decentralized.py:126-137, dist_scvx_3d.py:94-107);
    no epsilon in the denominator here."""
    others = [q for q in names if q != name]
    h = np.zeros((T, len(others))); g = np.zeros((T, len(others), dc))
    for k, q in enumerate(others):
        diff = X_traj[name][:, :dc] - X_traj[q][:, :dc]
        nrm = np.linalg.norm(diff, axis=1)
        h[:, k] = 2 * R - nrm
        g[:, k] = diff / nrm[:, None]
    return h, g


def solve_sbar_qp(s_pos, r_dual, rho, h, g, c_S=1e6):
    """ADMM_decentralized.py:106-139 for ONE robot: separable over t.  min_{sbar_t, S_t} r_t.(s_t - sbar_t) + rho/2 |s_t - sbar_t|^2
    + c_S S_t  s.t. h_tq - g_tq.sbar_t <= S_t, S_t >= 0.   s_pos (T,2), r_dual (T,2), h (T,nq), g (T,nq,2) -> sbar (T,2), S (T,)."""
    T, nq = h.shape
    sbar = np.zeros((T, 2)); S = np.zeros(T)
    for t in range(T):
        # variables (sbar_x, sbar_y, S)
        rows = [[-g[t, k, 0], -g[t, k, 1], -1.0] for k in range(nq)]
        A = sp.csc_matrix(np.array(rows).reshape(nq, 3))
        cvec = np.array([-r_dual[t, 0] - rho * s_pos[t, 0], -r_dual[t, 1] - rho * s_pos[t, 1], c_S])
        r = highs.solve(cvec, A, np.full(nq, -np.inf), -h[t], np.array([-np.inf, -np.inf, 0.0]), np.full(3, np.inf),
                        Q=sp.diags([rho, rho, 0.0]), time_limit=2.0)
        if r["ok"]:
            sbar[t] = r["x"][:2]; S[t] = r["x"][2]
        else:          # HiGHS' QP solver occasionally stalls even on these 3-variable problems: exact KKT enumeration instead
            sbar[t], S[t] = _sbar_enumerate(s_pos[t], r_dual[t], rho, h[t], g[t], c_S)
    return sbar, S


def _sbar_enumerate(s_t, r_t, rho, h, g, c_S):
    """Exact solution of one 3-variable consensus QP by enumerating active sets (<= 3 rows, S pinned or free)."""
    nq = len(h)
    c = s_t + r_t / rho
    best = (np.inf, c.copy(), 0.0)
    for k in range(0, min(3, nq) + 1):
        for idx in itertools.combinations(range(nq), k):
            for free_S in ((0, 1) if k >= 1 else (0,)):
                n = 2 + free_S + k
                A = np.zeros((n, n)); b = np.zeros(n)
                A[0, 0] = A[1, 1] = rho; b[:2] = rho * c
                for j, q in enumerate(idx):
                    A[0, 2 + free_S + j] = -g[q, 0]; A[1, 2 + free_S + j] = -g[q, 1]
                row = 2
                if free_S:
                    A[row, 3:3 + k] = 1.0; b[row] = c_S; row += 1
                for j, q in enumerate(idx):
                    A[row, 0:2] = g[q]; A[row, 2] = 1.0 if free_S else A[row, 2]; b[row] = h[q]; row += 1
                try:
                    x = np.linalg.solve(A, b)
                except np.linalg.LinAlgError:
                    continue
                sb, S = x[:2], (x[2] if free_S else 0.0)
                lam = x[2 + free_S:]
                ok = S >= -1e-9 and (lam >= -1e-9 * c_S).all() and (free_S or lam.sum() <= c_S * (1 + 1e-9))
                ok = ok and ((h - g @ sb - S) <= 1e-9 * np.maximum(1.0, np.abs(h))).all()
                if ok:
                    obj = 0.5 * rho * ((sb - c) ** 2).sum() + c_S * max(S, 0.0)
                    if obj < best[0]:
                        best = (obj, sb.copy(), max(S, 0.0))
    return best[1], best[2]


def x_traj_opt_2d(X_traj, trust_region, names, x_des, Ad, Bd, T, n=4, m=2, R=2.3, n_admm=5, rho=1.0):
    """ADMM_decentralized.py:32-170.  X_traj: dict name -> (T, n+m).  Returns the updated dict (a copy) and a log."""
    r_all = {k: np.ones((T, 2)) * 10 for k in names}
    s_val = {k: np.zeros((T, n + m)) for k in names}
    sbar = {k: np.zeros((T, 2)) for k in names}
    log = []
    for _ in range(n_admm):
        for k in names:
            q = RobotQP(Ad, Bd, X_traj[k][:, :n], X_traj[k][:T - 1, n:], x_des[k], trust_region, 100.0, rho=rho,
                        lin=r_all[k], sbar=sbar[k])
            r = solve_robot_qp_certified(q)
            s_val[k] = np.hstack([r["d"], r["w"]])
        new = {}
        for k in names:
            h, g = collision_rows(X_traj, names, k, R, 2, T)
            new[k], _S = solve_sbar_qp(s_val[k][:, :2], r_all[k], rho, h, g)
        diff = 0.0
        for k in names:
            r_all[k] = r_all[k] + rho * (s_val[k][:, :2] - new[k])
            diff += np.linalg.norm(new[k] - s_val[k][:, :2], 2)
            sbar[k] = new[k]
        log.append(diff)
    return {k: X_traj[k] + s_val[k] for k in names}, log


def x_traj_opt_3d(X_traj, trust_region, names, x_des, Ad, Bd, T, n=6, m=3, R=2.3):
    """dist_scvx_3d.py:31-118: one QP per robot, collision rows linearised about every robot's PREVIOUS trajectory (Jacobi)."""
    s_val = {}
    objs = {}
    for k in names:
        h, g = collision_rows(X_traj, names, k, R, 3, T)
        q = RobotQP(Ad, Bd, X_traj[k][:, :n], X_traj[k][:T - 1, n:], x_des[k], trust_region, 1.0, col_h=h[:T - 1], col_g=g[:T - 1],
                    c_S=1e4)
        r = solve_robot_qp_certified(q)
        s_val[k] = np.hstack([r["d"], r["w"]]); objs[k] = r["obj"]
    return {k: X_traj[k] + s_val[k] for k in names}, objs


def cost_fcn(X_traj, names, T, n=6, m=3):
    """dist_scvx_3d.py:131-138."""
    return float(sum((X_traj[k][:T - 1, n:n + m] ** 2).sum() for k in names))
