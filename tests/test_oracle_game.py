"""Oracle statement of the Nash best-response cost (game_model.py:69-126) and the mirror's quadratic tables."""
import numpy as np

from oracle import foh as ofoh, models as omodels, subproblem as ospb


def _direct_cost(X, U, X_prev, cw, rw, tw, iw):
    """sum_squares terms exactly as GameUnicycleModel.get_cost_function writes them."""
    c = cw * (U ** 2).sum()
    c += rw * ((U[:, 1:] - U[:, :-1]) ** 2).sum()
    c += tw * ((X[2, 1:] - X[2, :-1]) ** 2).sum()
    c += iw * ((X - X_prev) ** 2).sum()
    return c


def _params(K=12):
    m = omodels.unicycle(r_init=[0.0, -1.0, 0.0], r_final=[2.0, 3.0, 0.0], obstacles=[([1.0, 1.0], 0.25)])
    X, U = m.initialize_trajectory(K)
    mats = ofoh.OracleFOH(m, K).calculate_discretization(X, U, 1.0)
    return ospb.Params(m, K, mats, X, U, 1.0, 100.0), X, U


def test_tables_reproduce_the_reference_cost_and_gradient():
    from scvx_b200.optimization.agent_best_response import game_tables

    class W:       # weights only: game_tables needs nothing else from the model
        def __init__(self, **kw): self.kw = kw
        def get_cost_function(self): return self.kw

    rng = np.random.default_rng(3)
    p, X0, U0 = _params()
    K = p.K
    X_prev = X0 + 0.1 * rng.normal(size=X0.shape)
    cw, rw, tw, iw = 5.0, 5.0, 100.0, 0.7
    qd, lw, qp, const = game_tables(W(control_weight=cw, control_rate_weight=rw, curvature_weight=tw, inertia_weight=iw,
                                      path_weight=0.0), X_prev, K)
    p.set_game(quad_diag=qd, lin_w=lw, quad_pair=qp, const=const)
    X = X0 + 0.3 * rng.normal(size=X0.shape); U = U0 + 0.3 * rng.normal(size=U0.shape)
    val, grad = p.game_cost(X, U)
    assert abs(val - _direct_cost(X, U, X_prev, cw, rw, tw, iw)) <= 1e-10 * abs(val)
    # gradient by central differences
    W0 = np.vstack([X, U]); num = np.zeros_like(W0); h = 1e-6
    for i in range(W0.shape[0]):
        for k in range(K):
            Wp, Wm = W0.copy(), W0.copy(); Wp[i, k] += h; Wm[i, k] -= h
            num[i, k] = (p.game_cost(Wp[:3], Wp[3:])[0] - p.game_cost(Wm[:3], Wm[3:])[0]) / (2 * h)
    assert np.abs(num - grad).max() <= 1e-5 * (1 + np.abs(grad).max())


def test_lp_lower_bound_is_below_every_feasible_value():
    """The certificate used by the GPU tests: linearising the convex quadratic about ANY point gives an LP whose optimum
    bounds the best-response optimum from below; evaluating at the LP's own minimiser bounds it from above."""
    p, X0, U0 = _params()
    p.set_game(quad_diag=[0, 0, 0, 10.0, 10.0], quad_pair=[0, 0, 200.0, 10.0, 10.0],
               hard_rows=[(np.tile(np.array([[0.0], [1.0]]), (1, p.K)), np.full(p.K, -5.0))])
    lb1 = ospb.solve(p, linearize_at=(X0, U0))
    assert lb1["ok"]
    e1 = ospb.evaluate(p, lb1["X"], lb1["U"], lb1["sigma"])
    assert e1["viol"] <= 1e-7 and lb1["obj"] <= e1["obj"] * (1 + 1e-9)
    lb2 = ospb.solve(p, linearize_at=(lb1["X"], lb1["U"]))
    assert lb2["obj"] <= e1["obj"] * (1 + 1e-9)
    assert abs(lb2["sigma"] - 1.0) < 1e-9            # sigma == sigma_ref
