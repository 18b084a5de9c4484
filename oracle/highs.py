"""Thin wrapper over scipy's vendored HiGHS for the oracle's LPs / convex QPs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference hands its sub-problems to cvxpy -> ECOS / CLARABEL (scvx_solver.py:71,
admm_coordinator.py:76, Distributed_opt/ADMM_decentralized.py:97) which are not installed in the
build container; HiGHS is an exact LP/QP solver, and the optimal value of a convex programme does
not depend on the solver.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize._highspy import _core as hc

INF = hc.kHighsInf


def solve(c, A, row_lo, row_hi, col_lo, col_hi, Q=None, offset=0.0, solver="choose", time_limit=600.0):
    """min 0.5 x'Qx + c'x + offset  s.t. row_lo <= A x <= row_hi, col_lo <= x <= col_hi.

    A: scipy sparse (m, n); Q: scipy sparse symmetric PSD (n, n) or None.
    Returns dict(x, obj, status, ok).
    """
    A = sp.csc_matrix(A)
    m, n = A.shape
    lp = hc.HighsLp()
    lp.num_col_, lp.num_row_ = n, m
    lp.col_cost_ = np.asarray(c, dtype=float)
    lp.col_lower_ = np.asarray(np.clip(col_lo, -INF, INF), dtype=float)
    lp.col_upper_ = np.asarray(np.clip(col_hi, -INF, INF), dtype=float)
    lp.row_lower_ = np.asarray(np.clip(row_lo, -INF, INF), dtype=float)
    lp.row_upper_ = np.asarray(np.clip(row_hi, -INF, INF), dtype=float)
    lp.offset_ = float(offset)
    lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
    lp.a_matrix_.num_col_, lp.a_matrix_.num_row_ = n, m
    lp.a_matrix_.start_ = A.indptr.astype(np.int32)
    lp.a_matrix_.index_ = A.indices.astype(np.int32)
    lp.a_matrix_.value_ = A.data.astype(float)

    h = hc._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("time_limit", float(time_limit))
    h.setOptionValue("primal_feasibility_tolerance", 1e-9)
    h.setOptionValue("dual_feasibility_tolerance", 1e-9)
    if solver != "choose":
        h.setOptionValue("solver", solver)
    h.passModel(lp)
    if Q is not None:
        Ql = sp.csc_matrix(sp.tril(sp.csc_matrix(Q)))
        Ql.sort_indices()
        hs = hc.HighsHessian()
        hs.dim_ = n
        hs.format_ = hc.HessianFormat.kTriangular
        hs.start_ = Ql.indptr.astype(np.int32)
        hs.index_ = Ql.indices.astype(np.int32)
        hs.value_ = Ql.data.astype(float)
        h.passHessian(hs)
    h.run()
    status = h.modelStatusToString(h.getModelStatus())
    sol = h.getSolution()
    x = np.array(sol.col_value)
    return {"x": x, "obj": h.getObjectiveValue(), "status": status, "ok": status == "Optimal",
            "row_dual": np.array(sol.row_dual), "col_dual": np.array(sol.col_dual)}
