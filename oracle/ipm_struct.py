"""CPU twin of the GPU sub-problem solver's ALGORITHM (csrc/solver.cu) -- TEST INFRASTRUCTURE ONLY.

This is not a restatement of reference code: the reference hands the sub-problem to cvxpy+ECOS
(scvx_solver.py:71).  It restates, in numpy, the structure-exploiting primal-dual interior-point
method that the CUDA kernel implements (same formulation, same block-tridiagonal + border
elimination, same Mehrotra predictor-corrector), so that (i) the algorithm could be validated against
the exact HiGHS oracle (oracle/subproblem.py) before any CUDA was written and (ii) kernel bugs can be
bisected iterate by iterate.  The problem statement it solves is SURVEY Appendix A.1/A.3
(sc_problem.py:21-83, unicycle_model.py:85-115, single_integrator_model.py:79-128,
agent_solver.py:79-102).

Formulation (per agent), z = (w_0..w_{K-1}, sigma, t_nu, t_x, t_u), w_k = (x_k, u_k):
  min  c_s [ w_nu t_nu + w_sig sigma + sum_hinge w_h max(0, b - a.p_k) + rho/2 |P|^2 + <lin, P> ]
  s.t. e.nu_k(z) - t_nu <= 0          for all sign patterns e in {+-1}^{n_x}, k = 0..K-2   (|nu_k|_1 <= t_nu)
       e.(x_k - xref_k) - t_x <= 0    for all e, k                                          (|dx_k|_1 <= t_x)
       e.(u_k - uref_k) - t_u <= 0    for all e, k
       t_x + t_u +- (sigma - sigma_ref) <= r ;  -sigma <= 0
       pos_lo <= p_k <= pos_hi ;  unicycle: 0 <= v_k <= v_max, |w_k| <= w_max ;  SI: 1/2(|u_k|^2 - v_max^2) <= 0
       w_0, w_{K-1} fixed (boundary conditions).
  Hinge terms are handled with a private slack xi per row that is eliminated analytically from the
  Newton system (2 inequality rows per hinge: -a.p - xi <= -b, -xi <= 0).
"""
from __future__ import annotations

import itertools

import numpy as np

from . import subproblem as spb


def _signs(n):
    return np.array(list(itertools.product([1.0, -1.0], repeat=n)))[:, ::-1].copy()   # bit i of e -> sign of comp i


def spd_inverse_frozen(Dj, piv_tol=1e-12, reg_rel=0.0):
    """Inverse of an SPD block through its Cholesky factor, the way the kernel does it in registers: a pivot that has
    lost all significance (d <= piv_tol * original diagonal) is FROZEN -- its reciprocal is set to 0, which removes
    that row/column from the step instead of injecting round-off noise.  Returns Li (inverse factor), Dinv = Li'Li."""
    ns = Dj.shape[0]
    L = np.zeros((ns, ns)); dinv = np.zeros(ns)
    reg = reg_rel * np.trace(Dj) / ns
    for j in range(ns):
        d0 = Dj[j, j] + reg
        d = d0 - (L[j, :j] ** 2).sum()
        if not (d > piv_tol * d0):
            dinv[j] = 0.0
            L[j, j] = 0.0
            L[j + 1:, j] = 0.0
            continue
        rs = 1.0 / np.sqrt(d)
        dinv[j] = rs
        L[j, j] = d * rs
        for i in range(j + 1, ns):
            L[i, j] = (Dj[i, j] - (L[i, :j] * L[j, :j]).sum()) * rs
    Li = np.zeros((ns, ns))
    for j in range(ns):
        Li[j, j] = dinv[j]
        for i in range(j + 1, ns):
            Li[i, j] = -(L[i, j:i] * Li[j:i, j]).sum() * dinv[i]
    return Li, Li.T @ Li


class StructIPM:
    def __init__(self, p: spb.Params, mu0=10.0, max_iter=60, eps_gap=1e-8, eps_feas=1e-9, verbose=False, linalg="cr", step_frac=0.999, uncapped=False,
                 gondzio=0, g_thresh=0.5, g_delta=0.3, weigh_second_order=True, adaptive_step_frac=True):
        self.weigh_second_order, self.adaptive_step_frac = bool(weigh_second_order), bool(adaptive_step_frac)
        # gondzio > 0: EXPERIMENT, not in the kernel: up to `gondzio` centrality correctors after a Mehrotra step whose
        # min(alpha_p, alpha_d) < g_thresh; n_solves counts the linear solves (2 per iteration + 1 per corrector tried)
        self.gondzio, self.g_thresh, self.g_delta, self.n_solves = int(gondzio), float(g_thresh), float(g_delta), 0
        self.p, self.verbose, self.linalg = p, verbose, linalg
        self.step_frac, self.uncapped = step_frac, uncapped
        self.mu0, self.max_iter, self.eps_gap, self.eps_feas = mu0, max_iter, eps_gap, eps_feas
        m, K = p.model, p.K
        self.K, self.nx, self.nu, self.d = K, m.n_x, m.n_u, m.d
        self.ns = self.nx + self.nu
        self.ball = (m.kind != "unicycle")
        nx, nu = self.nx, self.nu
        # interval Jacobians: nu_k = Jn_k w_{k+1} + Jp_k w_k + Js_k sigma - zbar_k
        self.Jp = np.zeros((K - 1, nx, self.ns)); self.Jn = np.zeros((K - 1, nx, self.ns)); self.Js = np.zeros((K - 1, nx))
        for k in range(K - 1):
            A = p.A_bar[:, k].reshape((nx, nx), order="F"); B = p.B_bar[:, k].reshape((nx, nu), order="F")
            C = p.C_bar[:, k].reshape((nx, nu), order="F")
            self.Jp[k, :, :nx] = -A; self.Jp[k, :, nx:] = -B
            self.Jn[k, :, :nx] = np.eye(nx); self.Jn[k, :, nx:] = -C
            self.Js[k] = -p.S_bar[:, k]
        self.Ex, self.Eu = _signs(nx), _signs(nu)
        # cost scaling
        self.cs = 1.0 / max(p.weight_nu, p.weight_sigma, 1e-300)
        # hinge tables: a (nh, d, K), b (nh, K), w (nh,)
        a_list, b_list, w_list = [], [], []
        for j in range(len(m.obstacles)):
            a_list.append(p.obs_a[j]); b_list.append(p.obs_rhs[j] + np.einsum("dk,d->k", p.obs_a[j], p.obs_c[j]))
            w_list.append(p.weight_slack)
        for nb in p.neighbors:
            a_list.append(nb["a"]); b_list.append(p.d_min + np.einsum("dk,dk->k", nb["a"], nb["Y"])); w_list.append(p.weight_col)
        self.nh = len(a_list)
        self.ha = np.array(a_list).reshape(self.nh, self.d, K)
        self.hb = np.array(b_list).reshape(self.nh, K)
        self.hw = np.array(w_list).reshape(self.nh) * self.cs
        # quadratic / linear position terms (ADMM variant)
        self.qrho = 0.0; self.qlin = np.zeros((self.d, K))
        for nb in p.neighbors:
            self.qrho += p.rho
            self.qlin += nb["Lam"] - p.rho * nb["Y"]
        self.qrho *= self.cs; self.qlin *= self.cs
        self.c_sig = p.weight_sigma * self.cs; self.c_tnu = p.weight_nu * self.cs
        lo_p, hi_p = m.lower_bound + m.robot_radius, m.upper_bound - m.robot_radius
        self.lo_p, self.hi_p = lo_p, hi_p

    # ------------------------------------------------------------------------------------------
    def _nu(self, W, sig):
        """defects for all intervals: (K-1, nx)"""
        return (np.einsum("kij,kj->ki", self.Jn, W[1:]) + np.einsum("kij,kj->ki", self.Jp, W[:-1])
                + self.Js * sig - self.p.z_bar.T)

    def solve(self):
        p, K, nx, nu, ns, d = self.p, self.K, self.nx, self.nu, self.ns, self.d
        m = p.model
        Ex, Eu = self.Ex, self.Eu
        nEx, nEu = len(Ex), len(Eu)
        free = np.ones(K, bool); free[0] = free[K - 1] = False
        Wref = np.vstack([p.X_ref, p.U_ref]).T.copy()          # (K, ns)
        r_tr = p.tr_radius
        # ---- initial point -------------------------------------------------------------------
        W = Wref.copy()
        W[0, :nx] = m.x_init; W[K - 1, :nx] = m.x_final; W[0, nx:] = 0.0; W[K - 1, nx:] = 0.0
        dlt = min(1e-2 * (self.hi_p - self.lo_p), r_tr / (16.0 * ns))
        W[1:-1, :d] = np.clip(W[1:-1, :d], self.lo_p + dlt, self.hi_p - dlt)
        if not self.ball:
            dv = min(1e-2 * m.v_max, r_tr / (16.0 * ns)); dw = min(1e-2 * 2 * m.w_max, r_tr / (16.0 * ns))
            W[1:-1, nx] = np.clip(W[1:-1, nx], dv, m.v_max - dv)
            W[1:-1, nx + 1] = np.clip(W[1:-1, nx + 1], -m.w_max + dw, m.w_max - dw)
        else:
            nrm = np.linalg.norm(W[1:-1, nx:], axis=1)
            sc = np.minimum(1.0, 0.99 * m.v_max / np.maximum(nrm, 1e-300))
            W[1:-1, nx:] *= sc[:, None]
        sig = max(p.sigma_ref, min(1e-2, r_tr / 16.0))
        nuv = self._nu(W, sig)
        t_nu = np.abs(nuv).sum(axis=1).max() * 1.1 + 1.0
        t_x = np.abs(W[:, :nx] - Wref[:, :nx]).sum(axis=1).max() + r_tr / 4
        t_u = np.abs(W[:, nx:] - Wref[:, nx:]).sum(axis=1).max() + r_tr / 4
        mu0 = self.mu0
        # a start below the default is for warm problems only: where the start point violates a hinge row the solve starts cold
        # (the kernel's rule; see csrc/solver_kernel.cuh)
        if mu0 < 10.0 and self.nh and (self.hb - np.einsum("hdk,dk->hk", self.ha, W[:, :d].T))[:, 1:-1].max() > 1e-6:
            mu0 = 10.0

        def plain_slacks(W, sig, t_nu, t_x, t_u):
            """h - G z for every plain row family (positive = satisfied)."""
            nuv = self._nu(W, sig)
            sN = t_nu - nuv @ Ex.T                                   # (K-1, 8)
            sX = t_x - (W[:, :nx] - Wref[:, :nx]) @ Ex.T             # (K, 8)
            sU = t_u - (W[:, nx:] - Wref[:, nx:]) @ Eu.T             # (K, nEu)
            sG = np.array([r_tr - t_x - t_u - (sig - p.sigma_ref), r_tr - t_x - t_u + (sig - p.sigma_ref), sig])
            sP = np.concatenate([self.hi_p - W[:, :d], W[:, :d] - self.lo_p], axis=1)   # (K, 2d)
            if not self.ball:
                sV = np.stack([m.v_max - W[:, nx], W[:, nx], m.w_max - W[:, nx + 1], m.w_max + W[:, nx + 1]], axis=1)
            else:
                sV = (0.5 * (m.v_max ** 2 - (W[:, nx:] ** 2).sum(axis=1)))[:, None]
            return sN, sX, sU, sG, sP, sV

        sN, sX, sU, sG, sP, sV = [np.maximum(a, 1e-8) for a in plain_slacks(W, sig, t_nu, t_x, t_u)]
        lN, lX, lU, lG, lP, lV = mu0 / sN, mu0 / sX, mu0 / sU, mu0 / sG, mu0 / sP, mu0 / sV
        # hinge rows (free stages only)
        P_ = W[:, :d].T                                               # (d, K)
        viol = self.hb - np.einsum("hdk,dk->hk", self.ha, P_)         # b - a.p  (nh, K)
        hw = self.hw[:, None]
        disc = np.sqrt((hw * viol) ** 2 + 4 * mu0 ** 2)
        num = np.where(viol >= 0, hw * viol + disc, 4 * mu0 ** 2 / np.maximum(disc - hw * viol, 1e-300))
        xi = (num + 2 * mu0) / (2 * hw)
        s2 = xi.copy(); s1 = xi - viol
        l1, l2 = mu0 / s1, mu0 / s2
        hfree = np.broadcast_to(free[None, :], xi.shape)

        n_rows = (K - 1) * nEx + K * nEx + K * nEu + 3 + (K - 2) * (2 * d + sV.shape[1]) + 2 * self.nh * (K - 2)
        status, it = 1, 0
        for it in range(self.max_iter):
            # ---- residuals ---------------------------------------------------------------------
            gN, gX, gU, gG, gP, gV = plain_slacks(W, sig, t_nu, t_x, t_u)     # h - Gz
            rN, rX, rU, rG, rP, rV = sN - gN, sX - gX, sU - gU, sG - gG, sP - gP, sV - gV   # r_p = Gz + s - h
            viol = self.hb - np.einsum("hdk,dk->hk", self.ha, W[:, :d].T)
            r1 = viol - xi + s1            # -a.p - xi + s1 + b
            r2 = -xi + s2
            fm = free[:, None]
            comp = (sN * lN).sum() + (sX * lX).sum() + (sU * lU).sum() + (sG * lG).sum() + ((sP * lP) * fm).sum() \
                + ((sV * lV) * fm).sum() + ((s1 * l1 + s2 * l2) * hfree).sum()
            mu = comp / n_rows
            rp_inf = max(np.abs(rN).max(), np.abs(rX).max(), np.abs(rU).max(), np.abs(rG).max(),
                         np.abs(rP[free]).max(), np.abs(rV[free]).max(),
                         np.abs(r1[:, free]).max() if self.nh else 0.0, np.abs(r2[:, free]).max() if self.nh else 0.0)
            # stationarity  q + P z + G' lam   (built through the same scatter as the rhs)
            rdW, rdg = self._scatter(lN, lX, lU, lG, lP, lV, l1, W, free)
            rdW[:, :d] += self.qrho * W[:, :d] + self.qlin.T
            rdg[0] += self.c_sig; rdg[1] += self.c_tnu
            rxi = self.hw[:, None] - l1 - l2
            rd_inf = max(np.abs(rdW[free]).max(), np.abs(rdg).max(), np.abs(rxi[:, free]).max() if self.nh else 0.0)
            obj = self.c_sig * sig + self.c_tnu * t_nu + (self.hw[:, None] * xi * hfree).sum() \
                + 0.5 * self.qrho * (W[:, :d] ** 2).sum() + (self.qlin.T * W[:, :d]).sum()
            if self.verbose:
                print(f"{it:3d} mu={mu:.2e} rp={rp_inf:.2e} rd={rd_inf:.2e} obj={obj / self.cs:.8e}")
            cmax = max(self.c_tnu, self.c_sig, self.hw.max() if self.nh else 0.0)
            gap_ok = comp <= self.eps_gap * max(abs(obj), 1e-3)
            deep = comp <= 1e-4 * self.eps_gap * max(abs(obj), 1e-3)      # far past the gap target: rd is at its roundoff floor
            if gap_ok and rp_inf <= self.eps_feas and rd_inf <= (1e-5 if deep else 1e-7) * (1.0 + cmax):
                status = 0
                break
            # stall rule: when the optimum is tiny in scaled cost units the relative gap target sits below what fp64
            # can deliver (mu stalls around 1e-11); once the gap is below 1e-7 (1 + |obj|) and stops shrinking (or the
            # dual residual blows up) the current -- primal feasible -- iterate is accepted.
            if it > 0 and comp <= 1e-7 * (1.0 + abs(obj)) and rp_inf <= self.eps_feas and \
                    (comp > 0.5 * comp_prev or (rd_inf > 10.0 * rd_prev and rd_inf > 1e-8 * (1.0 + cmax))):
                status = 0
                break
            comp_prev, rd_prev = comp, rd_inf
            if not np.isfinite(mu):
                status = 2
                break
            # ---- Newton matrix -------------------------------------------------------------------
            wN, wX, wU, wG, wP, wV = lN / sN, lX / sX, lU / sU, lG / sG, lP / sP, lV / sV
            w1, w2 = l1 / s1, l2 / s2
            weff = w1 * w2 / (w1 + w2)
            fac = self._factor(wN, wX, wU, wG, wP, wV, lV, weff, W, free)

            def newton(sigmu, cN, cX, cU, cG, cP, cV, c1, c2):
                """One solve with complementarity targets: tau = (sigmu - c + lam*r_p)/s."""
                tN = (sigmu - cN + lN * rN) / sN; tX = (sigmu - cX + lX * rX) / sX; tU = (sigmu - cU + lU * rU) / sU
                tG = (sigmu - cG + lG * rG) / sG; tP = (sigmu - cP + lP * rP) / sP; tV = (sigmu - cV + lV * rV) / sV
                t1 = (sigmu - c1 + l1 * r1) / s1; t2 = (sigmu - c2 + l2 * r2) / s2
                rhs_xi = -self.hw[:, None] + t1 + t2
                # rhs_z = -(Pz+q) - G'tau ; hinge part: + a*(t1 - w1*rhs_xi/(w1+w2))
                th = t1 - w1 * rhs_xi / (w1 + w2)
                bW, bg = self._scatter(tN, tX, tU, tG, tP, tV, th, W, free)
                bW[:, :d] += self.qrho * W[:, :d] + self.qlin.T
                bg[0] += self.c_sig; bg[1] += self.c_tnu
                bW, bg = -bW, -bg
                bW[~free] = 0.0
                dW, dg = self._solve(fac, bW, bg)
                # recover row steps
                dnu = (np.einsum("kij,kj->ki", self.Jn, dW[1:]) + np.einsum("kij,kj->ki", self.Jp, dW[:-1]) + self.Js * dg[0])
                dsN = -rN - (dnu @ Ex.T - dg[1]); dsX = -rX - (dW[:, :nx] @ Ex.T - dg[2]); dsU = -rU - (dW[:, nx:] @ Eu.T - dg[3])
                dsG = -rG - np.array([dg[2] + dg[3] + dg[0], dg[2] + dg[3] - dg[0], -dg[0]])
                dsP = -rP - np.concatenate([dW[:, :d], -dW[:, :d]], axis=1)
                if not self.ball:
                    dsV = -rV - np.stack([dW[:, nx], -dW[:, nx], dW[:, nx + 1], -dW[:, nx + 1]], axis=1)
                else:
                    dsV = -rV - (W[:, nx:] * dW[:, nx:]).sum(axis=1)[:, None]
                adp = np.einsum("hdk,dk->hk", self.ha, dW[:, :d].T)
                dxi = (rhs_xi - w1 * adp) / (w1 + w2)
                ds1 = -r1 + adp + dxi; ds2 = -r2 + dxi
                dl = lambda l, s, c, ds: -l + (sigmu - c) / s - (l / s) * ds      # noqa: E731
                return (dW, dg, dxi, (dsN, dsX, dsU, dsG, dsP, dsV, ds1, ds2),
                        (dl(lN, sN, cN, dsN), dl(lX, sX, cX, dsX), dl(lU, sU, cU, dsU), dl(lG, sG, cG, dsG),
                         dl(lP, sP, cP, dsP), dl(lV, sV, cV, dsV), dl(l1, s1, c1, ds1), dl(l2, s2, c2, ds2)))

            S = (sN, sX, sU, sG, sP, sV, s1, s2); L = (lN, lX, lU, lG, lP, lV, l1, l2)
            masks = (None, None, None, None, fm, fm, hfree, hfree)

            def maxstep(vals, dvals, cap=1.0):
                a = cap
                for v, dv, mk in zip(vals, dvals, masks):
                    neg = dv < 0
                    if mk is not None:
                        neg = neg & np.broadcast_to(mk, dv.shape)
                    if neg.any():
                        a = min(a, float((-v[neg] / dv[neg]).min()))
                return a

            z0 = 0.0
            dW, dg, dxi, dS, dL = newton(0.0, z0, z0, z0, z0, z0, z0, z0, z0)
            ap, ad = maxstep(S, dS), maxstep(L, dL)
            comp_aff = 0.0
            for v, dv, l, dl_, mk in zip(S, dS, L, dL, masks):
                t = (v + ap * dv) * (l + ad * dl_)
                comp_aff += (t * mk).sum() if mk is not None else t.sum()
            sg = (comp_aff / comp) ** 3
            # Mehrotra's second-order term presumes a full affine step; it is weighed by om = min(alpha_p, alpha_d) of the affine step
            # (round 2; on 256 sub-problems of the bench scenes, together with the step fraction below and mu0 = 1e-3: 10.4 -> 7.9 iterations)
            # (problems with one common step length -- quadratic cost, ball row -- keep the full term: weighing it costs them iterations)
            om = min(ap, ad) if (self.weigh_second_order and not (self.qrho > 0 or self.ball)) else 1.0
            cc = [om * dv * dl_ for dv, dl_ in zip(dS, dL)]
            dW, dg, dxi, dS, dL = newton(sg * mu, *cc)
            cap = 1e300 if self.uncapped else 1.0
            ap, ad = maxstep(S, dS, cap), maxstep(L, dL, cap)
            if self.qrho > 0 or self.ball:
                ap = ad = min(ap, ad)
            self.n_solves += 2
            for _g in range(self.gondzio):
                # Gondzio's centrality corrector: look a little further along the direction, project the complementarity products
                # of that trial point onto [0.1, 10] x the target, and re-solve with the projection error folded into the
                # second-order term (the Newton system is linear in it: one more solve with the same factorisation)
                if min(ap, ad) >= self.g_thresh:
                    break
                tp, td = min(1.0, ap + self.g_delta), min(1.0, ad + self.g_delta)
                mut = sg * mu
                cc_g = []
                for v, dv, l, dl_, c_ in zip(S, dS, L, dL, cc):
                    vt = (v + tp * dv) * (l + td * dl_)
                    cc_g.append(c_ - np.maximum(np.clip(vt, 0.1 * mut, 10.0 * mut) - vt, -10.0 * mut))
                cand = newton(sg * mu, *cc_g)
                self.n_solves += 1
                ap_g, ad_g = maxstep(S, cand[3], cap), maxstep(L, cand[4], cap)
                if self.qrho > 0 or self.ball:
                    ap_g = ad_g = min(ap_g, ad_g)
                if ap_g + ad_g <= 1.01 * (ap + ad):
                    break
                (dW, dg, dxi, dS, dL), ap, ad, cc = cand, ap_g, ad_g, cc_g
            # fraction of the way to the boundary: step_frac while the centring target is large, up to 1 - 1e-6 as it vanishes
            sf = min(0.999999, max(self.step_frac, 1.0 - 1000.0 * sg * mu)) if self.adaptive_step_frac else self.step_frac
            ap, ad = min(1.0, sf * ap), min(1.0, sf * ad)
            W = W + ap * dW; sig += ap * dg[0]; t_nu += ap * dg[1]; t_x += ap * dg[2]; t_u += ap * dg[3]
            xi = xi + ap * np.where(hfree, dxi, 0.0)
            # rows of the fixed stages are placeholders (the kernel never touches them): keep them frozen
            upd = lambda v, dv, a, mk: v + a * (dv if mk is None else np.where(np.broadcast_to(mk, dv.shape), dv, 0.0))   # noqa: E731
            sN, sX, sU, sG, sP, sV, s1, s2 = [upd(v, dv, ap, mk) for v, dv, mk in zip(S, dS, masks)]
            lN, lX, lU, lG, lP, lV, l1, l2 = [upd(l, dl_, ad, mk) for l, dl_, mk in zip(L, dL, masks)]
            # slacks are not independent state in the kernel: every plain row is linear and the start is strictly
            # feasible, so s = h - G z is RECOMPUTED from the iterate (consistent with z to round-off, half the traffic)
            tiny = 1e-14
            gN2, gX2, gU2, gG2, gP2, gV2 = plain_slacks(W, sig, t_nu, t_x, t_u)
            sN, sX, sU, sG = np.maximum(gN2, tiny), np.maximum(gX2, tiny), np.maximum(gU2, tiny), np.maximum(gG2, tiny)
            sP = np.where(fm, np.maximum(gP2, tiny), sP)
            if not self.ball:          # the quadratic ball row keeps its slack as an independent variable
                sV = np.where(fm, np.maximum(gV2, tiny), sV)
            viol2 = self.hb - np.einsum("hdk,dk->hk", self.ha, W[:, :d].T)
            s1 = np.where(hfree, np.maximum(xi - viol2, tiny), s1); s2 = np.where(hfree, np.maximum(xi, tiny), s2)
        X = W[:, :nx].T.copy(); U = W[:, nx:].T.copy()
        return {"X": X, "U": U, "sigma": float(sig), "iters": it, "status": status, "t": (t_nu, t_x, t_u)}

    # ------------------------------------------------------------------------------------------
    def _scatter(self, vN, vX, vU, vG, vP, vV, vH, W, free):
        """G' v for row-wise values v (per family); hinge rows contribute -a * vH on positions.
        Returns (K, ns) stage part and (4,) global part."""
        K, nx, d = self.K, self.nx, self.d
        Ex, Eu = self.Ex, self.Eu
        outW = np.zeros((K, self.ns)); outg = np.zeros(4)
        eN = vN @ Ex                                  # (K-1, nx): sum_e v_e e
        outW[:-1] += np.einsum("ki,kij->kj", eN, self.Jp)
        outW[1:] += np.einsum("ki,kij->kj", eN, self.Jn)
        outg[0] += (eN * self.Js).sum(); outg[1] -= vN.sum()
        outW[:, :nx] += vX @ Ex; outg[2] -= vX.sum()
        outW[:, nx:] += vU @ Eu; outg[3] -= vU.sum()
        outg[0] += vG[0] - vG[1] - vG[2]; outg[2] += vG[0] + vG[1]; outg[3] += vG[0] + vG[1]
        fm = free[:, None]
        vPm = vP * fm
        outW[:, :d] += vPm[:, :d] - vPm[:, d:]
        vVm = vV * fm
        if not self.ball:
            outW[:, nx] += vVm[:, 0] - vVm[:, 1]; outW[:, nx + 1] += vVm[:, 2] - vVm[:, 3]
        else:
            outW[:, nx:] += vVm[:, :1] * W[:, nx:]
        if self.nh:
            outW[:, :d] -= np.einsum("hdk,hk->kd", self.ha, vH * free[None, :])
        return outW, outg

    def _factor(self, wN, wX, wU, wG, wP, wV, lV, weff, W, free):
        """Assemble H = P + G'WG as block tridiagonal (D_k, E_k = H[k+1,k]) + 4 border columns, then
        block-Cholesky it.  Fixed stages (0, K-1) are replaced by identity rows/cols."""
        K, nx, nu, ns, d = self.K, self.nx, self.nu, self.ns, self.d
        Ex, Eu = self.Ex, self.Eu
        D = np.zeros((K, ns, ns)); E = np.zeros((K - 1, ns, ns)); Bd = np.zeros((K, ns, 4)); Gg = np.zeros((4, 4))
        MN = np.einsum("ke,ei,ej->kij", wN, Ex, Ex)          # (K-1, nx, nx)
        mN = wN @ Ex                                         # (K-1, nx)
        sNw = wN.sum(axis=1)
        D[:-1] += np.einsum("kia,kij,kjb->kab", self.Jp, MN, self.Jp)
        D[1:] += np.einsum("kia,kij,kjb->kab", self.Jn, MN, self.Jn)
        E += np.einsum("kia,kij,kjb->kab", self.Jn, MN, self.Jp)
        MJs = np.einsum("kij,kj->ki", MN, self.Js)           # (K-1, nx)
        Bd[:-1, :, 0] += np.einsum("kia,ki->ka", self.Jp, MJs); Bd[1:, :, 0] += np.einsum("kia,ki->ka", self.Jn, MJs)
        Bd[:-1, :, 1] -= np.einsum("kia,ki->ka", self.Jp, mN); Bd[1:, :, 1] -= np.einsum("kia,ki->ka", self.Jn, mN)
        Gg[0, 0] += (self.Js * MJs).sum(); Gg[0, 1] -= (mN * self.Js).sum(); Gg[1, 1] += sNw.sum()
        MX = np.einsum("ke,ei,ej->kij", wX, Ex, Ex); D[:, :nx, :nx] += MX
        Bd[:, :nx, 2] -= wX @ Ex; Gg[2, 2] += wX.sum()
        MU = np.einsum("ke,ei,ej->kij", wU, Eu, Eu); D[:, nx:, nx:] += MU
        Bd[:, nx:, 3] -= wU @ Eu; Gg[3, 3] += wU.sum()
        gG = np.array([[1.0, 0, 1, 1], [-1.0, 0, 1, 1], [-1.0, 0, 0, 0]])
        Gg += np.einsum("r,ri,rj->ij", wG, gG, gG)
        Gg[1, 0] = Gg[0, 1]
        fm = free.astype(float)
        for i in range(d):
            D[:, i, i] += (wP[:, i] + wP[:, d + i]) * fm + self.qrho
        if not self.ball:
            D[:, nx, nx] += (wV[:, 0] + wV[:, 1]) * fm; D[:, nx + 1, nx + 1] += (wV[:, 2] + wV[:, 3]) * fm
        else:
            Uv = W[:, nx:]
            D[:, nx:, nx:] += (wV[:, 0] * fm)[:, None, None] * np.einsum("ki,kj->kij", Uv, Uv)
            for i in range(nu):
                D[:, nx + i, nx + i] += lV[:, 0] * fm
        if self.nh:
            D[:, :d, :d] += np.einsum("hk,hik,hjk->kij", weff * fm[None, :], self.ha, self.ha)
        # fixed stages -> identity
        for k in (0, K - 1):
            D[k] = np.eye(ns); Bd[k] = 0.0
        E[0] = 0.0; E[K - 2] = 0.0
        if self.linalg == "cr":
            fac = {"Bd": Bd, "Gg": Gg, "ok": True}
            fac.update(self._cr_factor(D[1:K - 1].copy(), E[1:K - 2].copy()))
        else:
            # sequential block Cholesky (the first kernel version)
            Lk = np.zeros((K, ns, ns)); Li = np.zeros((K, ns, ns)); Lo = np.zeros((K - 1, ns, ns))
            ok = True
            for k in range(K):
                Dk = D[k] - (Lo[k - 1] @ Lo[k - 1].T if k > 0 else 0.0)
                Li[k], _ = spd_inverse_frozen(Dk)
                if k < K - 1:
                    Lo[k] = E[k] @ Li[k].T
            fac = {"Li": Li, "Lo": Lo, "Bd": Bd, "Gg": Gg, "ok": ok}
        # Y = T^-1 B' (4 rhs) and the Schur complement
        Y = self._tsolve(fac, Bd)
        fac["Y"] = Y
        fac["S"] = Gg - np.einsum("kia,kib->ab", Bd, Y)
        return fac

    # ---- block cyclic reduction over the free stages (the kernel's factorisation) ---------------------------
    def _cr_factor(self, D, Lc):
        """D: (n, ns, ns) diagonal blocks of the free stages, Lc: (n-1, ns, ns) with Lc[i] = T[i+1, i].
        Odd-even elimination with stride s = 1, 2, 4, ...; positions are 1-based (p = index + 1): at stride s the
        nodes p = s, 3s, 5s, ... are eliminated.  Stored per node: Dinv, P = Dinv T[j, left], Q = Dinv T[j, right]."""
        n, ns = D.shape[0], self.ns
        Lm = np.zeros((n, ns, ns))
        Lm[1:] = Lc                       # Lm[i] = coupling of node i to its current LEFT neighbour
        Dinv = np.zeros_like(D); P = np.zeros_like(D); Q = np.zeros_like(D)
        s = 1
        while s <= n:
            odd = [p - 1 for p in range(s, n + 1, 2 * s)]
            keepL = {}
            for j in odd:                                         # phase 1: invert, P, Q, push-left
                a, b = j - s, j + s
                Li, _ = spd_inverse_frozen(D[j])
                Dinv[j] = Li                       # the inverse FACTOR is stored; D^-1 v is applied as Li'(Li v)
                Lj = Lm[j].copy() if a >= 0 else np.zeros((ns, ns))
                P[j] = Li.T @ (Li @ Lj)
                Q[j] = Li.T @ (Li @ Lm[b].T) if b < n else np.zeros((ns, ns))
                if a >= 0:
                    D[a] -= Lj.T @ P[j]
            for j in odd:                                         # phase 2: push-right, new left coupling of b
                b = j + s
                if b < n:
                    D[b] -= Lm[b] @ Q[j]
                    Lm[b] = -Lm[b] @ P[j]
            s *= 2
        return {"Dinv": Dinv, "P": P, "Q": Q}

    def _cr_solve(self, fac, R):
        """Solve T V = R for R (n, ns, nrhs) with the cyclic-reduction factor."""
        Dinv, P, Q = fac["Dinv"], fac["P"], fac["Q"]
        n = R.shape[0]
        V = R.copy()
        strides = []
        s = 1
        while s <= n:
            strides.append(s)
            for p in range(2 * s, n + 1, 2 * s):                  # even nodes pull from their odd neighbours
                a = p - 1
                V[a] -= Q[a - s].T @ V[a - s]
                if a + s < n:
                    V[a] -= P[a + s].T @ V[a + s]
            s *= 2
        for s in reversed(strides):
            for p in range(s, n + 1, 2 * s):
                j = p - 1
                x = Dinv[j].T @ (Dinv[j] @ V[j])
                if j - s >= 0:
                    x -= P[j] @ V[j - s]
                if j + s < n:
                    x -= Q[j] @ V[j + s]
                V[j] = x
        return V

    def _tsolve(self, fac, R):
        """Solve T V = R for R (K, ns, nrhs) (fixed stages 0 and K-1 are identity rows)."""
        K = self.K
        if self.linalg == "cr":
            V = R.copy()
            V[1:K - 1] = self._cr_solve(fac, R[1:K - 1])
            return V
        Li, Lo = fac["Li"], fac["Lo"]
        V = np.zeros_like(R)
        for k in range(K):
            r = R[k] - (Lo[k - 1] @ V[k - 1] if k > 0 else 0.0)
            V[k] = Li[k] @ r
        for k in range(K - 1, -1, -1):
            r = V[k] - (Lo[k].T @ V[k + 1] if k < K - 1 else 0.0)
            V[k] = Li[k].T @ r
        return V

    def _solve(self, fac, bW, bg):
        v = self._tsolve(fac, bW[:, :, None])[:, :, 0]
        rg = bg - np.einsum("kia,ki->a", fac["Bd"], v)
        dg = np.linalg.solve(fac["S"], rg)
        dW = v - np.einsum("kia,a->ki", fac["Y"], dg)
        return dW, dg
