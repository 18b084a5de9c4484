"""Oracle restatement of the reference's dynamics models and stage-2 linearisations.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy, fp64.

Follows (reference paths):
  SCvx/models/unicycle_model.py:27-71            UnicycleModel defaults, f/A/B
  SCvx/models/unicycle_model.py:73-83            initialize_trajectory
  SCvx/models/unicycle_model.py:103-114          obstacle half-space linearisation
  SCvx/models/single_integrator_model.py:24-63   SingleIntegratorModel defaults, f/A/B
  SCvx/models/single_integrator_model.py:113-126 obstacle linearisation (3-D, + MARGIN_OBS)
  SCvx/models/multi_agent_model.py:61-79         linearize_collision
  SCvx/models/SI_multi_agent_model.py:49-74      linearize_inter_agent_collision
"""
from __future__ import annotations

import numpy as np

EPS_NORMAL = 1e-6  # the "+1e-6" in every normal's denominator (unicycle_model.py:111)
MARGIN_OBS = 0.0   # SCvx/config/SI_default_game.py:9


class OracleModel:
    """Plain-data description of one agent's model (no cvxpy objects)."""

    def __init__(self, kind, x_init, x_final, v_max, w_max, bounds, robot_radius, obstacles):
        assert kind in ("unicycle", "single_integrator")
        self.kind = kind
        self.n_x = 3
        self.n_u = 2 if kind == "unicycle" else 3
        self.d = 2 if kind == "unicycle" else 3          # position dimension
        self.x_init = np.asarray(x_init, dtype=float).reshape(-1)
        self.x_final = np.asarray(x_final, dtype=float).reshape(-1)
        self.v_max = float(v_max)
        self.w_max = float(w_max)
        self.lower_bound, self.upper_bound = float(bounds[0]), float(bounds[1])
        self.robot_radius = float(robot_radius)
        self.obstacles = [(np.asarray(c, dtype=float).reshape(-1), float(r)) for c, r in obstacles]

    # -- dynamics ---------------------------------------------------------------------------
    def f(self, x, u):
        if self.kind == "unicycle":      # unicycle_model.py:56
            return np.array([u[0] * np.cos(x[2]), u[0] * np.sin(x[2]), u[1]])
        return np.array(u, dtype=float)  # single_integrator_model.py:54

    def A(self, x, u):
        if self.kind == "unicycle":      # unicycle_model.py:57 (jacobian wrt x)
            return np.array([[0.0, 0.0, -u[0] * np.sin(x[2])],
                             [0.0, 0.0, u[0] * np.cos(x[2])],
                             [0.0, 0.0, 0.0]])
        return np.zeros((3, 3))          # single_integrator_model.py:56

    def B(self, x, u):
        if self.kind == "unicycle":      # unicycle_model.py:58 (jacobian wrt u)
            return np.array([[np.cos(x[2]), 0.0],
                             [np.sin(x[2]), 0.0],
                             [0.0, 1.0]])
        return np.eye(3)                 # single_integrator_model.py:57

    # -- warm start -------------------------------------------------------------------------
    def initialize_trajectory(self, K):
        """Straight line X, zero U (unicycle_model.py:73-83)."""
        X = np.zeros((self.n_x, K))
        for k in range(K):
            a1 = (K - 1 - k) / (K - 1)
            a2 = k / (K - 1)
            X[:, k] = a1 * self.x_init + a2 * self.x_final
        return X, np.zeros((self.n_u, K))

    def obstacle_clearance(self, j):
        c, r = self.obstacles[j]
        extra = MARGIN_OBS if self.kind == "single_integrator" else 0.0
        return r + self.robot_radius + extra


def unicycle(r_init=(-8.0, -8.0, 0.0), r_final=(8.0, 8.0, 0.0), v_max=1.0, w_max=np.pi / 6,
             bounds=(-10.0, 10.0), robot_radius=0.5, obstacles=None):
    """UnicycleModel defaults (unicycle_model.py:27-49)."""
    if obstacles is None:
        obstacles = [([5.0, 4.0], 3.0), ([-5.0, -4.0], 3.0), ([0.0, 0.0], 2.0)]
    return OracleModel("unicycle", r_init, r_final, v_max, w_max, bounds, robot_radius, obstacles)


def single_integrator(r_init=(-8.0, -8.0, -8.0), r_final=(8.0, 8.0, 8.0), v_max=1.0,
                      bounds=(-10.0, 10.0), robot_radius=0.5, obstacles=None):
    """SingleIntegratorModel defaults (single_integrator_model.py:24-49)."""
    if obstacles is None:
        obstacles = [([-5.0, -4.0, -5.0], 2.0), ([0.0, 0.0, 4.0], 2.0)]
    return OracleModel("single_integrator", r_init, r_final, v_max, 0.0, bounds, robot_radius, obstacles)


# -- stage 2 --------------------------------------------------------------------------------
def linearize_obstacles(model: OracleModel, X_ref):
    """Obstacle half-spaces  a_jk^T (p_k - c_j) >= r_j + r_rob (+margin) - s'_jk.

    unicycle_model.py:103-114 / single_integrator_model.py:113-126.
    Returns a (M, d, K), rhs (M,) [the clearance], centres (M, d).
    """
    d, K = model.d, X_ref.shape[1]
    M = len(model.obstacles)
    a = np.zeros((M, d, K))
    rhs = np.zeros(M)
    ctr = np.zeros((M, d))
    for j, (c, _r) in enumerate(model.obstacles):
        ctr[j] = c
        rhs[j] = model.obstacle_clearance(j)
        for k in range(K):
            diff = X_ref[0:d, k] - c
            a[j, :, k] = diff / (np.linalg.norm(diff) + EPS_NORMAL)
    return a, rhs, ctr


def linearize_collision(d, d_min, X_ref_i, X_ref_j):
    """multi_agent_model.py:61-79 / SI_multi_agent_model.py:49-74."""
    p_i = X_ref_i[0:d, :]
    p_j = X_ref_j[0:d, :]
    K = p_i.shape[1]
    A_ij = np.zeros((d, K))
    b_ij = np.zeros(K)
    for k in range(K):
        diff = p_i[:, k] - p_j[:, k]
        a = diff / (np.linalg.norm(diff) + EPS_NORMAL)
        A_ij[:, k] = a
        b_ij[k] = d_min + a.dot(p_j[:, k])
    return A_ij, b_ij
