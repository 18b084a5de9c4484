"""CPU: the stage-3 oracle (HiGHS LP restatement of sc_problem.py:21-83) on its own -- known optimal value from the
survey, evaluator consistency, norm-mode switch, the LP bracket used to pin the QP variant, and the CPU twin of the
GPU algorithm against the exact solver."""
import numpy as np
import pytest

from oracle import foh as ofoh, models as omodels, scvx as oscvx, subproblem as ospb
from oracle.ipm_struct import StructIPM
from oracle.models import linearize_collision


@pytest.fixture(scope="module")
def prob0():
    m = omodels.unicycle()
    K = 50
    X, U = m.initialize_trajectory(K)
    mats = ofoh.OracleFOH(m, K).calculate_discretization(X, U, 1.0)
    return m, K, X, U, mats


def test_iteration0_optimal_value_matches_survey(prob0):
    """with the documented weights (w_nu=1e3, w_sigma=1, r=20) the first LP's optimum is 7.050690e+03
    for both HiGHS simplex and IPM."""
    m, K, X, U, mats = prob0
    p = ospb.Params(m, K, mats, X, U, 1.0, 20.0, weight_nu=1e3, weight_sigma=1.0)
    r1, r2 = ospb.solve(p), ospb.solve(p, solver="ipm")
    assert r1["ok"] and r2["ok"]
    assert r1["obj"] == pytest.approx(7050.690, rel=1e-6) and r2["obj"] == pytest.approx(r1["obj"], rel=1e-9)
    e = ospb.evaluate(p, r1["X"], r1["U"], r1["sigma"])
    assert e["obj"] == pytest.approx(r1["obj"], rel=1e-9) and e["viol"] <= 1e-9


def test_norm_mode_switch(prob0):
    """cvxpy's norm(M, 1) is the induced norm; the entry-wise reading gives a different (larger) penalty."""
    m, K, X, U, mats = prob0
    a = ospb.solve(ospb.Params(m, K, mats, X, U, 1.0, 100.0, norm1_mode="induced"))
    b = ospb.solve(ospb.Params(m, K, mats, X, U, 1.0, 100.0, norm1_mode="entrywise"))
    assert a["ok"] and b["ok"] and b["obj"] > a["obj"] * 1.5
    assert ospb.norm1(np.array([[1.0, -2.0], [3.0, 4.0]]), "induced") == 6.0
    assert ospb.norm1(np.array([[1.0, -2.0], [3.0, 4.0]]), "entrywise") == 10.0


def test_algorithm_twin_matches_exact_lp(prob0):
    m, K, X, U, mats = prob0
    p = ospb.Params(m, K, mats, X, U, 1.0, 100.0)
    r = ospb.solve(p)
    s = StructIPM(p).solve()
    e = ospb.evaluate(p, s["X"], s["U"], s["sigma"])
    assert s["status"] == 0 and e["viol"] <= 1e-9 and abs(e["obj"] - r["obj"]) <= 1e-8 * abs(r["obj"])


def test_qp_bracket_certifies_and_rejects():
    """LB <= f_opt <= f(z): tight at the optimum, visibly open at a perturbed (feasible) point."""
    N, K = 3, 16
    rng = np.random.default_rng(4)
    ang = np.linspace(0, 2 * np.pi, N, endpoint=False)
    ms = [omodels.unicycle(r_init=[6 * np.cos(a), 6 * np.sin(a), 0], r_final=[-6 * np.cos(a), -6 * np.sin(a), 0],
                           obstacles=[([0.0, 0.0], 1.0)]) for a in ang]
    XU = [m.initialize_trajectory(K) for m in ms]
    mats = ofoh.OracleFOH(ms[0], K).calculate_discretization(XU[0][0], XU[0][1], 15.0)
    nbrs = []
    for j in (1, 2):
        a, _ = linearize_collision(2, 0.5, XU[0][0], XU[j][0])
        nbrs.append({"a": a, "Y": XU[j][0][:2].copy(), "Lam": 0.1 * rng.normal(size=(2, K))})
    p = ospb.Params(ms[0], K, mats, XU[0][0], XU[0][1], 15.0, 100.0, neighbors=nbrs, rho=1.0, d_min=0.5)
    s = StructIPM(p).solve()
    f0, lb, viol, ok = ospb.qp_bracket(p, s["X"], s["U"], s["sigma"])
    assert ok and viol <= 1e-9 and -1e-9 * abs(f0) <= f0 - lb <= 1e-7 * abs(f0)
    f1, lb1, _, _ = ospb.qp_bracket(p, p.X_ref, p.U_ref, p.sigma_ref)          # the reference point: feasible, not optimal
    assert f1 - lb1 > 1e-3 * abs(f1) and lb1 <= f0 * (1 + 1e-9)


def test_outer_loop_quirks():
    """Grow-only trust region (cap 50, floor 1e-3) and record keys (scvx_solver.py:89-97,125-133)."""
    assert oscvx.update_trust_region(100.0, 1.0, 0.0) == 50.0
    assert oscvx.update_trust_region(10.0, 1e-3, 1e-3) == 15.0
    assert oscvx.update_trust_region(10.0, 1.0, 0.0) == 12.0
    assert oscvx.update_trust_region(1e-4, 1.0, 0.0) == 1e-3
    _, _, _, rec = oscvx.scvx_solve(omodels.unicycle(), 20, max_iter=2)
    assert {"iter", "nu_norm", "slack_norm", "dx", "du", "ds", "sigma"} <= set(rec[0])
