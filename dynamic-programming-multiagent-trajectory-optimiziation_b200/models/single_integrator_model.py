"""SingleIntegratorModel -- mirrors SCvx/models/single_integrator_model.py:12-141 (3-D, f = u)."""
from typing import List, Optional, Tuple

import numpy as np

from .. import _lib
from ..global_parameters import K
from .base_model import BaseModel, ConstraintTables, SlackValue, straight_line_guess

MARGIN_OBS = 0.0   # SCvx/config/SI_default_game.py:9


class SingleIntegratorModel(BaseModel):
    n_x = 3
    n_u = 3
    device_model_id = _lib.MODEL_SINGLE_INTEGRATOR

    def __init__(
        self,
        r_init: np.ndarray = np.array([-8.0, -8.0, -8.0]),
        r_final: np.ndarray = np.array([8.0, 8.0, 8.0]),
        v_max: float = 1.0,
        bounds: Tuple[float, float] = (-10.0, 10.0),
        robot_radius: float = 0.5,
        obstacles: Optional[List[Tuple[List[float], float]]] = None,
    ):
        super().__init__()
        self.x_init = np.asarray(r_init, dtype=float).reshape(-1)
        self.x_final = np.asarray(r_final, dtype=float).reshape(-1)
        self.v_max = v_max
        self.lower_bound, self.upper_bound = bounds
        self.robot_radius = robot_radius
        self.obstacles = (
            obstacles if obstacles is not None else [([-5.0, -4.0, -5.0], 2.0), ([0.0, 0.0, 4.0], 2.0)]
        )
        self.s_prime = [SlackValue(K) for _ in self.obstacles]
        self.f = lambda x, u: u
        self.A = lambda x, u: np.zeros((self.n_x, self.n_x))
        self.B = lambda x, u: np.eye(self.n_x)

    def get_equations(self) -> Tuple:
        return self.f, self.A, self.B

    def initialize_trajectory(self, X: np.ndarray, U: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Straight line in state space, zero controls; fills and returns the caller's arrays."""
        return straight_line_guess(self.x_init, self.x_final, X, U)

    def get_constraints(self, X=None, U=None, X_ref=None, U_ref=None) -> ConstraintTables:
        """single_integrator_model.py:79-128: boundary, ||u_k||_2 <= v_max, box, spherical obstacles."""
        M = len(self.obstacles)
        return ConstraintTables(
            x_init=self.x_init, x_final=self.x_final,
            pos_lo=self.lower_bound + self.robot_radius, pos_hi=self.upper_bound - self.robot_radius,
            v_max=self.v_max, w_max=0.0, input_kind="ball",
            obs_centres=np.array([np.asarray(p, dtype=float).reshape(3) for p, _ in self.obstacles]).reshape(M, 3),
            obs_clearance=np.array([r + self.robot_radius + MARGIN_OBS for _, r in self.obstacles], dtype=float),
        )

    def get_objective(self, X=None, U=None, X_ref=None, U_ref=None):
        return {"kind": "slack_sum", "weight": 1e5}
