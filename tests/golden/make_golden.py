"""Generate golden vectors from the UNMODIFIED reference Python modules (stages 1-2).

Run once in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Writes tests/golden/stage12_golden.npz.  The reference modules are imported through
oracle/refshim.py (an inert `cvxpy` stub on sys.path, SCvx.global_parameters.K patched before any
other SCvx import).  Everything stored is an output of the reference's own code:

  * FirstOrderHold.calculate_discretization at odeint DEFAULT tolerances (= what the reference runs)
  * the reference's own FirstOrderHold._ode_dVdt integrated at rtol=1e-13/atol=1e-14 (the parity
    target for the CUDA RK4 kernel: same right-hand side, tight integration)
  * integrate_nonlinear_piecewise / integrate_nonlinear_full
  * UnicycleModel f/A/B lambdas, initialize_trajectory
  * MultiAgentModel.linearize_collision, SI_MultiAgentModel.linearize_inter_agent_collision
"""
import os
import sys

import numpy as np
from scipy.integrate import odeint

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402


def tight_discretization(foh, X, U, sigma):
    n_x, n_u, K = foh.n_x, foh.n_u, foh.K
    A_bar = np.zeros((n_x * n_x, K - 1)); B_bar = np.zeros((n_x * n_u, K - 1))
    C_bar = np.zeros((n_x * n_u, K - 1)); S_bar = np.zeros((n_x, K - 1)); z_bar = np.zeros((n_x, K - 1))
    V0 = foh.V0.copy()
    for k in range(K - 1):
        V0[foh.x_ind] = X[:, k]
        V = odeint(foh._ode_dVdt, V0, [0.0, foh.dt], args=(U[:, k], U[:, k + 1], sigma),
                   rtol=1e-13, atol=1e-14, mxstep=100000)[1]
        Phi = V[foh.A_ind].reshape((n_x, n_x), order="F")
        A_bar[:, k] = Phi.flatten(order="F")
        B_bar[:, k] = (Phi @ V[foh.B_ind].reshape((n_x, n_u), order="F")).flatten(order="F")
        C_bar[:, k] = (Phi @ V[foh.C_ind].reshape((n_x, n_u), order="F")).flatten(order="F")
        S_bar[:, k] = Phi @ V[foh.S_ind]
        z_bar[:, k] = Phi @ V[foh.z_ind]
    return A_bar, B_bar, C_bar, S_bar, z_bar


def main():
    out = {}
    rng = np.random.default_rng(20261018)

    # ---- unicycle, K = 50 -------------------------------------------------------------------
    K = 50
    refshim.load(K)
    from SCvx.discretization.first_order_hold import FirstOrderHold
    from SCvx.models.multi_agent_model import MultiAgentModel
    from SCvx.models.unicycle_model import UnicycleModel

    m = UnicycleModel()
    foh = FirstOrderHold(m, K)
    for tag, sigma in (("s1", 1.0), ("s24", 24.0)):
        X = np.vstack([rng.uniform(-8, 8, (2, K)), rng.uniform(-3, 3, (1, K))])
        U = np.vstack([rng.uniform(0, 1, (1, K)), rng.uniform(-0.5, 0.5, (1, K))])
        out[f"uni_{tag}_X"], out[f"uni_{tag}_U"], out[f"uni_{tag}_sigma"] = X, U, sigma
        for nm, arr in zip("ABCSz", foh.calculate_discretization(X, U, sigma)):
            out[f"uni_{tag}_ref_{nm}"] = arr.copy()      # FOH returns aliases of its buffers
        for nm, arr in zip("ABCSz", tight_discretization(foh, X, U, sigma)):
            out[f"uni_{tag}_tight_{nm}"] = arr
        out[f"uni_{tag}_nl_piecewise"] = foh.integrate_nonlinear_piecewise(X, U, sigma)
        out[f"uni_{tag}_nl_full"] = foh.integrate_nonlinear_full(X[:, 0], U, sigma)

    # warm start (straight line, U = 0) + its discretisation -- config 1, iteration 0
    X0, U0 = m.initialize_trajectory(np.zeros((3, K)), np.zeros((2, K)))
    out["uni_init_X"], out["uni_init_U"] = X0, U0
    for nm, arr in zip("ABCSz", foh.calculate_discretization(X0, U0, 1.0)):
        out[f"uni_init_ref_{nm}"] = arr.copy()

    # f / A / B lambdas at random points
    f, A, B = m.get_equations()
    xs = rng.uniform(-3, 3, (8, 3)); us = rng.uniform(-1, 1, (8, 2))
    out["uni_fab_x"], out["uni_fab_u"] = xs, us
    out["uni_fab_f"] = np.stack([np.asarray(f(x, u), dtype=float).reshape(3) for x, u in zip(xs, us)])
    out["uni_fab_A"] = np.stack([np.asarray(A(x, u), dtype=float) for x, u in zip(xs, us)])
    out["uni_fab_B"] = np.stack([np.asarray(B(x, u), dtype=float) for x, u in zip(xs, us)])

    # inter-agent collision linearisation (2-D), incl. the coincident-agents corner (diff = 0)
    mam = MultiAgentModel([{"r_init": np.zeros(3), "r_final": np.ones(3)}] * 2, d_min=0.5)
    Xi = rng.uniform(-4, 4, (3, K)); Xj = rng.uniform(-4, 4, (3, K))
    Xj[:, 7] = Xi[:, 7]
    A_ij, b_ij = mam.linearize_collision(0, 1, Xi, Xj)
    out["col2_Xi"], out["col2_Xj"], out["col2_A"], out["col2_b"], out["col2_dmin"] = Xi, Xj, A_ij, b_ij, 0.5

    # ---- single integrator, K = 20 ------------------------------------------------------------
    K = 20
    refshim.load(K)
    from SCvx.discretization.first_order_hold import FirstOrderHold
    from SCvx.models.SI_multi_agent_model import SI_MultiAgentModel
    from SCvx.models.single_integrator_model import SingleIntegratorModel

    m = SingleIntegratorModel()
    foh = FirstOrderHold(m, K)
    X = rng.uniform(-8, 8, (3, K)); U = rng.uniform(-1, 1, (3, K)); sigma = 7.5
    out["si_X"], out["si_U"], out["si_sigma"] = X, U, sigma
    for nm, arr in zip("ABCSz", foh.calculate_discretization(X, U, sigma)):
        out[f"si_ref_{nm}"] = arr.copy()
    for nm, arr in zip("ABCSz", tight_discretization(foh, X, U, sigma)):
        out[f"si_tight_{nm}"] = arr
    out["si_nl_piecewise"] = foh.integrate_nonlinear_piecewise(X, U, sigma)

    smam = SI_MultiAgentModel([{"r_init": np.zeros(3), "r_final": np.ones(3)}] * 2, d_min=0.7)
    Xi = rng.uniform(-4, 4, (3, K)); Xj = rng.uniform(-4, 4, (3, K))
    A_ij, b_ij = smam.linearize_inter_agent_collision(0, 1, Xi, Xj)
    out["col3_Xi"], out["col3_Xj"], out["col3_A"], out["col3_b"], out["col3_dmin"] = Xi, Xj, A_ij, b_ij, 0.7

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage12_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
