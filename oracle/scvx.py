"""Oracle restatement of the reference's outer loops (drivers around stages 1-3).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (reference paths):
  SCvx/optimization/scvx_solver.py:33-115   SCVXSolver.solve (metrics, convergence test, returns the
                                            PREVIOUS iterate on convergence: break at :105 precedes :111)
  SCvx/optimization/scvx_solver.py:125-133  _update_trust_region (grow-only, cap 50, floor 1e-3)
  SCvx/optimization/admm_coordinator.py:39-118     ADMMCoordinator.solve (Gauss-Seidel sweep)
  SCvx/optimization/si_admm_coordinator.py:42-127  SI twin: setup() gets the INITIAL X_refs/U_refs every round
  SCvx/optimization/admm_utils.py:8-58      primal_residual / dual_residual / update_rho_admm
"""
from __future__ import annotations

import numpy as np

from . import subproblem as spb
from .foh import OracleFOH
from .models import linearize_collision

# SCvx/global_parameters.py:4-18
MAX_ITER, TRUST_RADIUS0, CONV_TOL = 30, 100.0, 1e-3
WEIGHT_NU, WEIGHT_SLACK, WEIGHT_SIGMA = 1e4, 1e6, 100.0


def primal_residual(p_j, Y_ij):          # admm_utils.py:8-18
    return float(np.linalg.norm(p_j - Y_ij))


def dual_residual(Y_new, Y_old):         # admm_utils.py:21-31
    return float(np.linalg.norm(Y_new - Y_old))


def update_rho_admm(rho, primal_res, dual_res, mu=10.0, tau_inc=2.0, tau_dec=2.0):   # admm_utils.py:34-58
    if primal_res > mu * dual_res:
        return rho * tau_inc
    elif dual_res > mu * primal_res:
        return rho / tau_dec
    return rho


def update_trust_region(tr, nu_norm, slack_norm):     # scvx_solver.py:125-133
    tr = min(tr * 1.5, 50.0) if (nu_norm < 1e-2 and slack_norm < 1e-2) else min(tr * 1.2, 50.0)
    return max(tr, 1e-3)


def scvx_solve(model, K, max_iter=MAX_ITER, tr_radius=TRUST_RADIUS0, conv_tol=CONV_TOL,
               weight_nu=WEIGHT_NU, weight_slack=WEIGHT_SLACK, weight_sigma=WEIGHT_SIGMA,
               initial_sigma=1.0, norm1_mode="induced", foh_tol="reference", solver="choose",
               X0=None, U0=None, return_problems=False):
    """scvx_solver.py:33-115.  Returns X, U, sigma, records (and the per-iteration Params if asked)."""
    foh = OracleFOH(model, K)
    X, U = model.initialize_trajectory(K) if X0 is None else (X0.copy(), U0.copy())
    sigma = float(initial_sigma)
    records, problems = [], []
    for it in range(max_iter):
        mats = foh.calculate_discretization(X, U, sigma, tol=foh_tol)
        p = spb.Params(model, K, mats, X, U, sigma, tr_radius, weight_nu, weight_slack, weight_sigma, norm1_mode)
        r = spb.solve(p, solver=solver)
        if not r["ok"]:
            raise RuntimeError(f"SCvx iteration {it}: convex subproblem infeasible")
        if return_problems:
            problems.append((p, r))
        X_new, U_new, nu_new, sigma_new = r["X"], r["U"], r["nu"], r["sigma"]
        nu_norm = float(np.abs(nu_new).sum(axis=0).max())       # np.linalg.norm(nu, 1), scvx_solver.py:82
        slack_norm = float(r["s_prime"].sum())
        dx = float(np.linalg.norm(X_new - X)); du = float(np.linalg.norm(U_new - U)); ds = abs(sigma_new - sigma)
        records.append({"iter": it, "nu_norm": nu_norm, "slack_norm": slack_norm, "dx": dx, "du": du,
                        "ds": ds, "sigma": sigma_new, "obj": r["obj"], "tr_radius": tr_radius})
        if nu_norm < conv_tol and slack_norm < conv_tol and dx < conv_tol and ds < conv_tol:
            break
        tr_radius = update_trust_region(tr_radius, nu_norm, slack_norm)
        X, U, sigma = X_new, U_new, sigma_new
    if return_problems:
        return X, U, sigma, records, problems
    return X, U, sigma, records


def _solve_qp(p):
    """HiGHS' QP solver does not finish on these problems; the QP variant is solved with the CPU twin of
    the interior-point algorithm and CERTIFIED by the exact LP bracket (subproblem.qp_bracket)."""
    from .ipm_struct import StructIPM
    s = StructIPM(p).solve()
    ev = spb.evaluate(p, s["X"], s["U"], s["sigma"])
    return {"ok": s["status"] == 0, "status": s["status"], "X": s["X"], "U": s["U"], "sigma": s["sigma"], "obj": ev["obj"]}


def admm_solve(models, d_min, K, X_refs, U_refs, sigma_ref, rho=1.0, max_iter=10, sweep="gauss_seidel",
               si_variant=False, norm1_mode="induced", foh_tol="reference", solver="choose"):
    """admm_coordinator.py:39-118 (si_variant=True: si_admm_coordinator.py:80-86 passes the INITIAL
    refs to setup()).  sweep="jacobi" is the batched-GPU order: every agent linearises about the
    previous round's neighbours (SURVEY hard part 5)."""
    N, d = len(models), models[0].d
    fohs = [OracleFOH(m, K) for m in models]
    Y = [X_refs[j][0:d, :].copy() for j in range(N)]        # identical for every observer i
    Lam = [np.zeros((d, K)) for _ in range(N)]
    X_curr = [x.copy() for x in X_refs]
    U_curr = [u.copy() for u in U_refs]
    primal_hist, dual_hist, objs = [], [], []
    for _ in range(max_iter):
        X_prev = [x.copy() for x in X_curr]
        new_pos = [None] * N
        round_objs = []
        for i in range(N):
            nbr_src = X_curr if sweep == "gauss_seidel" else X_prev
            own_X = X_curr[i] if sweep == "gauss_seidel" else X_prev[i]
            own_U = U_curr[i]
            mats = fohs[i].calculate_discretization(own_X, own_U, sigma_ref, tol=foh_tol)
            Xr, Ur = (X_refs[i], U_refs[i]) if si_variant else (own_X, own_U)
            nbrs = []
            for j in range(N):
                if j == i:
                    continue
                a, _b = linearize_collision(d, d_min, Xr, nbr_src[j])
                nbrs.append({"a": a, "Y": Y[j], "Lam": Lam[j]})
            p = spb.Params(models[i], K, mats, Xr, Ur, sigma_ref, TRUST_RADIUS0, WEIGHT_NU, WEIGHT_SLACK,
                           WEIGHT_SIGMA, norm1_mode, neighbors=nbrs, rho=rho, d_min=d_min)
            r = _solve_qp(p)
            if not r["ok"]:
                raise RuntimeError(f"agent {i}: sub-problem status {r['status']}")
            round_objs.append(r["obj"])
            X_curr[i], U_curr[i] = r["X"], r["U"]
            new_pos[i] = r["X"][0:d, :]
        pr, du = [], []
        for j in range(N):
            Y_old = Y[j]
            Y_new = 0.5 * (Y_old + new_pos[j])
            Y[j] = Y_new
            Lam[j] = Lam[j] + rho * (new_pos[j] - Y_new)
            pr.append(primal_residual(new_pos[j], Y_new)); du.append(dual_residual(Y_new, Y_old))
        primal_hist.append(float(np.mean(pr))); dual_hist.append(float(np.mean(du)))
        objs.append(round_objs)
    return X_curr, U_curr, sigma_ref, primal_hist, dual_hist, objs
