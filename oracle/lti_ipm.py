"""CPU twin of the GPU solver for the Distributed_opt per-robot QP (csrc/lti_qp.cu) -- TEST INFRASTRUCTURE ONLY.

Not reference code: the scripts call cvxpy -> CLARABEL (ADMM_decentralized.py:97, dist_scvx_3d.py:110).  This restates, in
numpy, the algorithm of the CUDA kernel: a primal-dual interior-point method whose Newton step is an equality-constrained
LQ problem solved by a Riccati recursion (dynamics as hard equalities), with the terminal equality handled by n extra
right-hand sides (one per terminal multiplier) and an n x n solve.  Problem data: oracle.distopt.RobotQP.

Rows (slack s > 0, multiplier lam > 0):
  trust   e . w_t <= tr                       for every sign pattern e in {+-1}^m,  t = 0..T-2
  box     lo_i <= (x_t + d_t)[i] <= hi_i      i = 0, 1,  t = 1..T-2   (t = 0 is fixed by d_0 = 0)
  coll    h_tq - g_tq . d_t[0:dc] - xi_t <= 0, -xi_t <= 0     t = 1..T-2; xi_t is eliminated from the Newton system
"""
from __future__ import annotations

import itertools

import numpy as np

from .distopt import RobotQP


class LtiIPM:
    def __init__(self, q: RobotQP, mu0=None, max_iter=60, eps_gap=1e-9, eps_feas=1e-9, verbose=False, xi_init='bisect', split_steps=True):
        self.xi_init, self.split_steps = xi_init, split_steps
        mu0 = max(1.0, q.c_S) if mu0 is None else mu0      # complementarity scale ~ the largest cost coefficient
        self.q, self.mu0, self.max_iter, self.eps_gap, self.eps_feas, self.verbose = q, mu0, max_iter, eps_gap, eps_feas, verbose
        self.E = np.array(list(itertools.product([1.0, -1.0], repeat=q.m)))[:, ::-1].copy()

    # -- equality-constrained LQ step by Riccati ----------------------------------------------------------------
    def _lq(self, Q, R, qv, rv, ctil, e_T):
        """min sum_t 1/2 dd_t'Q_t dd_t + qv_t.dd_t + 1/2 dw_t'R_t dw_t + rv_t.dw_t
        s.t. dd_{t+1} = A dd_t + B dw_t + ctil_t, dd_0 = 0, dd_{T-1} = e_T.   Returns dd (T,n), dw (T-1,m)."""
        q = self.q
        A, B, T, n, m = q.Ad, q.Bd, q.T, q.n, q.m
        nr = n + 1                                    # right-hand sides: the real one + one per terminal multiplier
        P = np.zeros((T, n, n)); p = np.zeros((T, n, nr))
        K = np.zeros((T - 1, m, n)); k = np.zeros((T - 1, m, nr))
        p[T - 1, :, 1:] = np.eye(n)
        for t in range(T - 2, -1, -1):
            Pn = P[t + 1]
            PB, PA = Pn @ B, Pn @ A
            Quu = R[t] + B.T @ PB
            Qux = B.T @ PA
            Qxx = Q[t] + A.T @ PA
            pc = p[t + 1].copy()
            pc[:, 0] += Pn @ ctil[t]
            qu = B.T @ pc; qx = A.T @ pc
            qu[:, 0] += rv[t]; qx[:, 0] += qv[t]
            Qi = np.linalg.inv(Quu + 1e-14 * np.trace(Quu) * np.eye(m))
            K[t] = -Qi @ Qux; k[t] = -Qi @ qu
            P[t] = Qxx + Qux.T @ K[t]
            P[t] = 0.5 * (P[t] + P[t].T)
            p[t] = qx + Qux.T @ k[t]
        dd = np.zeros((T, n, nr)); dw = np.zeros((T - 1, m, nr))
        for t in range(T - 1):
            dw[t] = K[t] @ dd[t] + k[t]
            dd[t + 1] = A @ dd[t] + B @ dw[t]
            dd[t + 1, :, 0] += ctil[t]
        M = dd[T - 1, :, 1:]                           # d(dd_{T-1}) / d(nu)
        nu = np.linalg.solve(M, e_T - dd[T - 1, :, 0])
        comb = np.concatenate([[1.0], nu])
        ddc, dwc = dd @ comb, dw @ comb
        # costates (multipliers of the dynamics rows): pi_t = P_t dd_t + p_t, the gradient of the cost-to-go
        pi = np.einsum("tij,tj->ti", P, ddc) + p @ comb
        return ddc, dwc, pi

    def solve(self):
        q = self.q
        A, B, T, n, m = q.Ad, q.Bd, q.T, q.n, q.m
        E = self.E
        nq = q.col_h.shape[1]
        dc = q.col_g.shape[2] if nq else 0
        mu0 = self.mu0
        lo = np.array([b[0] for b in q.box]); hi = np.array([b[1] for b in q.box])
        act = np.zeros(T - 1, bool); act[1:] = True                 # stages whose d_t is free (box / collision rows live there)
        # ---- start: d = 0 except the terminal value, w = 0 (strictly inside the trust region); equalities are NOT satisfied
        d = np.zeros((T, n)); w = np.zeros((T - 1, m))
        d[T - 1] = q.x_des[:n] - q.x[T - 1]
        sT = np.maximum(q.tr - w @ E.T, 1e-8); lT = mu0 / sT                         # (T-1, 2^m)
        pos = q.x[:T - 1, :2] + d[:T - 1, :2]
        sB = np.maximum(np.concatenate([hi - pos, pos - lo], axis=1), 1e-2); lB = mu0 / sB       # (T-1, 4)
        if nq:
            viol = (q.col_h - np.einsum("tqc,tc->tq", q.col_g, d[:T - 1, :dc]))       # (T-1, nq): row is viol - xi <= 0
            # dual-feasible central start of each stage's slack: sum_q mu0/(xi - viol_q) + mu0/xi = c_S  (bisection on the
            # monotone scalar equation; bracket [L, L + (nq+1) mu0 / c_S], L = max(viol_max, 0))
            L = np.maximum(viol.max(axis=1), 0.0)
            a_, b_ = L.copy(), L + (nq + 1) * mu0 / q.c_S
            for _ in range(60):
                mid = 0.5 * (a_ + b_)
                f = (mu0 / (mid[:, None] - viol)).sum(axis=1) + mu0 / mid - q.c_S
                a_ = np.where(f > 0, mid, a_); b_ = np.where(f > 0, b_, mid)
            xi = b_ if self.xi_init == 'bisect' else L + 1.0
            sC = xi[:, None] - viol; lC = mu0 / sC
            s0 = xi.copy(); l0 = mu0 / s0
        n_rows = sT.size + act.sum() * 4 + (act.sum() * (nq + 1) if nq else 0)
        cS = q.c_S
        status = 1
        pi = np.zeros((T, n))
        for it in range(self.max_iter):
            # residuals
            res = np.array([d[t + 1] - (A @ d[t] + B @ w[t] + q.c[t]) for t in range(T - 1)])     # equality residual
            rT = w @ E.T - q.tr + sT
            pos = q.x[:T - 1, :2] + d[:T - 1, :2]
            rB = np.concatenate([pos - hi, lo - pos], axis=1) + sB
            comp = (sT * lT).sum() + ((sB * lB) * act[:, None]).sum()
            rp = max(np.abs(res).max(), np.abs(rT).max(), np.abs(rB[act]).max())
            if nq:
                viol = (q.col_h - np.einsum("tqc,tc->tq", q.col_g, d[:T - 1, :dc]))
                rC = viol - xi[:, None] + sC
                r0 = -xi + s0
                comp += ((sC * lC) * act[:, None]).sum() + ((s0 * l0) * act).sum()
                rp = max(rp, np.abs(rC[act]).max(), np.abs(r0[act]).max())
            mu = comp / n_rows
            # the gap and the stagnation test are measured against the SMOOTH part of the objective: when the collision
            # slack (cost 1e4 per unit) dominates, a tolerance relative to the full value would leave w undetermined to ~1e-2
            smooth = q.c_w * ((q.u + w) ** 2).sum() + (q.lin * d[:, :2]).sum() + 0.5 * q.rho * ((d[:, :2] - q.sbar) ** 2).sum()
            obj = smooth
            if nq:
                obj += cS * (xi * act).sum()
            # stationarity with the current multipliers (lam, pi):  grad f + G'lam + A'pi_{t+1} - pi_t  /  ... + B'pi_{t+1}
            gw = 2 * q.c_w * (q.u + w) + lT @ E + pi[1:] @ B
            gd = np.zeros((T, n))
            gd[:, :2] += q.lin + q.rho * (d[:, :2] - q.sbar)
            gd[:T - 1, :2] += (lB[:, :2] - lB[:, 2:]) * act[:, None]
            if nq:
                gd[:T - 1, :dc] -= np.einsum("tq,tqc->tc", lC, q.col_g) * act[:, None]
            gd[:T - 1] += pi[1:] @ A
            gd -= pi
            rd = max(np.abs(gw).max(), np.abs(gd[1:T - 1]).max())
            if nq:
                rd = max(rd, np.abs((cS - lC.sum(axis=1) - l0)[act]).max())
            cmax = max(2 * q.c_w, cS, 1.0)
            if self.verbose:
                print(f"{it:3d} mu={mu:.2e} rp={rp:.2e} rd={rd:.2e} obj={obj:.8e}")
            # optimal: gap + primal feasibility, and either dual feasibility or a stagnated objective (on saturated problems the
            # multipliers lose accuracy long before the primal point does; the LP bracket certifies those solutions)
            scale = max(abs(smooth), 1.0)
            slack = obj - smooth
            stagn = it > 0 and abs(smooth - sm_prev) <= 1e-9 * scale and abs(slack - sl_prev) <= 1e-10 * max(abs(slack), 1.0)
            sm_prev, sl_prev = smooth, slack
            if comp <= self.eps_gap * scale and rp <= self.eps_feas and (rd <= 1e-8 * cmax or stagn):
                status = 0
                break
            WT, WB = lT / sT, lB / sB
            if nq:
                WC, W0 = lC / sC, l0 / s0
                Wsum = WC.sum(axis=1) + W0

            def step(sigmu, cT, cB, cC, c0):
                tT = (sigmu - cT + lT * rT) / sT
                tB = (sigmu - cB + lB * rB) / sB * act[:, None]
                R = np.einsum("te,ei,ej->tij", WT, E, E) + 2 * q.c_w * np.eye(m)
                rv = tT @ E + 2 * q.c_w * (q.u + w)
                Q = np.zeros((T, n, n)); qv = np.zeros((T, n))
                for i in range(2):
                    Q[:T - 1, i, i] += (WB[:, i] + WB[:, 2 + i]) * act + q.rho
                    qv[:T - 1, i] += tB[:, i] - tB[:, 2 + i]
                Q[T - 1, 0, 0] += q.rho; Q[T - 1, 1, 1] += q.rho
                qv[:, :2] += q.lin + q.rho * (d[:, :2] - q.sbar)
                if nq:
                    tC = (sigmu - cC + lC * rC) / sC
                    t0 = (sigmu - c0 + l0 * r0) / s0
                    rhs_xi = -cS + tC.sum(axis=1) + t0                       # stationarity of xi: cS - sum lamC - lam0
                    gW = np.einsum("tq,tqc->tc", WC, q.col_g)                  # sum_q W_q g_q
                    Hc = np.einsum("tq,tqc,tqe->tce", WC, q.col_g, q.col_g) - np.einsum("tc,te->tce", gW, gW) / Wsum[:, None, None]
                    gt = -np.einsum("tq,tqc->tc", tC, q.col_g) + gW * (rhs_xi / Wsum)[:, None]
                    Q[:T - 1, :dc, :dc] += Hc * act[:, None, None]
                    qv[:T - 1, :dc] += gt * act[:, None]
                dd, dw, pi_new = self._lq(Q, R, qv, rv, -res, np.zeros(n))
                self._pi_new = pi_new
                # (terminal and initial values are kept exactly: d_0 = 0, d_{T-1} fixed from the start)
                dsT = -rT - dw @ E.T
                dp = dd[:T - 1, :2]
                dsB = -rB - np.concatenate([dp, -dp], axis=1)
                dlT = -lT + (sigmu - cT) / sT - WT * dsT
                dlB = -lB + (sigmu - cB) / sB - WB * dsB
                out = [dd, dw, dsT, dsB, dlT, dlB]
                if nq:
                    gd = np.einsum("tqc,tc->tq", q.col_g, dd[:T - 1, :dc])
                    dxi = (rhs_xi - (WC * gd).sum(axis=1)) / Wsum
                    dsC = -rC + gd + dxi[:, None]
                    ds0 = -r0 + dxi
                    dlC = -lC + (sigmu - cC) / sC - WC * dsC
                    dl0 = -l0 + (sigmu - c0) / s0 - W0 * ds0
                    out += [dxi, dsC, ds0, dlC, dl0]
                return out

            def maxstep(pairs):
                a = 1.0
                for v, dv, mk in pairs:
                    neg = dv < 0
                    if mk is not None:
                        neg = neg & np.broadcast_to(mk, dv.shape)
                    if neg.any():
                        a = min(a, float((-v[neg] / dv[neg]).min()))
                return a

            z = 0.0
            o = step(0.0, z, z, z, z)
            mk1 = act[:, None]
            prs = [(sT, o[2], None), (sB, o[3], mk1), (lT, o[4], None), (lB, o[5], mk1)]
            if nq:
                prs += [(sC, o[7], mk1), (s0, o[8], act), (lC, o[9], mk1), (l0, o[10], act)]
            a_aff = maxstep(prs)
            comp_aff = ((sT + a_aff * o[2]) * (lT + a_aff * o[4])).sum() + (((sB + a_aff * o[3]) * (lB + a_aff * o[5])) * mk1).sum()
            if nq:
                comp_aff += (((sC + a_aff * o[7]) * (lC + a_aff * o[9])) * mk1).sum() + (((s0 + a_aff * o[8]) * (l0 + a_aff * o[10])) * act).sum()
            sg = min(max(comp_aff / comp, 0.0), 1.0) ** 3
            cT, cB = o[2] * o[4], o[3] * o[5]
            cC, c0 = (o[7] * o[9], o[8] * o[10]) if nq else (z, z)
            o = step(sg * mu, cT, cB, cC, c0)
            prs = [(sT, o[2], None), (sB, o[3], mk1), (lT, o[4], None), (lB, o[5], mk1)]
            if nq:
                prs += [(sC, o[7], mk1), (s0, o[8], act), (lC, o[9], mk1), (l0, o[10], act)]
            if self.split_steps:
                ap = min(1.0, 0.995 * maxstep([pr_ for pr_ in prs if pr_[0] is sT or pr_[0] is sB or (nq and (pr_[0] is sC or pr_[0] is s0))]))
                ad = min(1.0, 0.995 * maxstep([pr_ for pr_ in prs if pr_[0] is lT or pr_[0] is lB or (nq and (pr_[0] is lC or pr_[0] is l0))]))
            else:
                ap = ad = min(1.0, 0.995 * maxstep(prs))
            d = d + ap * o[0]; w = w + ap * o[1]
            pi = pi + ad * (self._pi_new - pi)
            sT = sT + ap * o[2]; lT = lT + ad * o[4]
            sB = np.where(mk1, sB + ap * o[3], sB); lB = np.where(mk1, lB + ad * o[5], lB)
            if nq:
                xi = np.where(act, xi + ap * o[6], xi)
                sC = np.where(mk1, sC + ap * o[7], sC); s0 = np.where(act, s0 + ap * o[8], s0)
                lC = np.where(mk1, lC + ad * o[9], lC); l0 = np.where(act, l0 + ad * o[10], l0)
        wfull = np.vstack([w, np.zeros((1, m))])
        return {"d": d, "w": wfull, "iters": it, "status": status}
