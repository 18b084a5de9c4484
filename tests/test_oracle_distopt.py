"""CPU: the Distributed_opt oracle (oracle/distopt.py) and the numpy twin of the Riccati interior-point algorithm."""
import numpy as np
import pytest

from oracle import distopt as od
from oracle.lti_ipm import LtiIPM


def _qp_2d(T=21):
    n, m = 4, 2
    Ad, Bd = od.descete_f(0.5, n, m)
    x_ini = np.array([0.0, 0.0, 0, 0, 0, 0]); x_des = np.array([3.0, 2.0, 0, 0, 0, 0])
    X = np.linspace(x_ini, x_des, T)
    rng = np.random.default_rng(0)
    return od.RobotQP(Ad, Bd, X[:, :n], X[:T - 1, n:], x_des, 0.25, 100.0, rho=1.0, lin=10 + rng.normal(size=(T, 2)),
                      sbar=0.1 * rng.normal(size=(T, 2)))


def test_descete_f_closed_form():
    for n, m, dt in ((4, 2, 0.5), (6, 3, 0.6)):
        Ad, Bd = od.descete_f(dt, n, m)
        d = n // 2
        want_A = np.eye(n); want_A[:d, d:] = dt * np.eye(d)
        want_B = np.vstack([0.5 * dt * dt * np.eye(d), dt * np.eye(d)])
        np.testing.assert_allclose(Ad, want_A, atol=1e-15); np.testing.assert_allclose(Bd, want_B, atol=1e-15)


def test_twin_matches_exact_qp_and_bracket():
    q = _qp_2d()
    r = od.solve_robot_qp(q)
    assert r["ok"] and q.objective(r["d"], r["w"]) == pytest.approx(r["obj"], rel=1e-9)
    s = LtiIPM(q).solve()
    assert s["status"] == 0 and q.violation(s["d"], s["w"]) <= 1e-9
    assert q.objective(s["d"], s["w"]) == pytest.approx(r["obj"], rel=1e-8)
    f0, lb, viol, ok = od.qp_bracket(q, s["d"], s["w"])
    assert ok and -1e-9 * abs(f0) <= f0 - lb <= 1e-7 * abs(f0)
    f1, lb1, _, _ = od.qp_bracket(q, r["d"] * 0.0 + np.linspace(0, 1, q.T)[:, None] * (q.x_des[:4] - q.x[-1]), r["w"] * 0.0)
    assert lb1 <= r["obj"] + 1e-6 * abs(r["obj"])


def test_sbar_qp_inactive_rows_give_unconstrained_minimiser():
    T = 5
    s = np.ones((T, 2)); r = np.full((T, 2), 2.0)
    h = -np.ones((T, 1)) * 5.0; g = np.tile(np.array([1.0, 0.0]), (T, 1, 1))
    sb, S = od.solve_sbar_qp(s, r, 1.0, h, g)
    np.testing.assert_allclose(sb, s + r, atol=1e-7); np.testing.assert_allclose(S, 0.0, atol=1e-9)


def test_helpers_match_the_reference_scripts_golden():
    """descete_f, x_initial, cost_fcn and the scenario globals against outputs of the reference's own function bodies
    (tests/golden/make_golden_distopt.py lifts them out of the scripts with ast; the scripts themselves cannot be imported)."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "distopt_golden.npz"))
    for tag in ("ad", "d3"):
        n, m = int(G[f"{tag}_n"]), int(G[f"{tag}_m"])
        Ad, Bd = od.descete_f(float(G[f"{tag}_dt"]), n, m)
        np.testing.assert_allclose(Ad, G[f"{tag}_Ad"], rtol=0, atol=1e-15); np.testing.assert_allclose(Bd, G[f"{tag}_Bd"], rtol=0, atol=1e-15)
        Ad, Bd = od.descete_f(0.37, n, m)
        np.testing.assert_allclose(Ad, G[f"{tag}_Ad_037"], rtol=0, atol=1e-15); np.testing.assert_allclose(Bd, G[f"{tag}_Bd_037"], rtol=0, atol=1e-15)


def test_mirror_modules_carry_the_reference_scenarios():
    """Module-level scenario of the two mirrors (no GPU call: importing them only builds numpy arrays)."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "distopt_golden.npz"))
    import scvx_b200.Distributed_opt.ADMM_decentralized as M2
    import scvx_b200.Distributed_opt.dist_scvx_3d as M3
    for tag, M in (("ad", M2), ("d3", M3)):
        assert M.T == int(G[f"{tag}_T"]) and M.n == int(G[f"{tag}_n"]) and M.m == int(G[f"{tag}_m"]) and M.R == float(G[f"{tag}_R"])
        assert M.dt == float(G[f"{tag}_dt"]) and list(M.robots_name) == [str(s) for s in G[f"{tag}_names"]]
        np.testing.assert_array_equal(np.stack([M.x_ini[k] for k in M.robots_name]), G[f"{tag}_x_ini"])
        np.testing.assert_array_equal(np.stack([M.x_des[k] for k in M.robots_name]), G[f"{tag}_x_des"])
        X0 = M.x_initial(M.x_ini, M.x_des)
        np.testing.assert_array_equal(np.stack([X0[k] for k in M.robots_name]), G[f"{tag}_X0"])
        np.testing.assert_allclose(M.Ad, G[f"{tag}_Ad"], rtol=0, atol=1e-15); np.testing.assert_allclose(M.Bd, G[f"{tag}_Bd"], rtol=0, atol=1e-15)
    Xr = {k: G["d3_cost_X"][i] for i, k in enumerate(M3.robots_name)}
    assert M3.cost_fcn(Xr) == pytest.approx(float(G["d3_cost"]), rel=1e-13)
