"""UnicycleModel -- mirrors SCvx/models/unicycle_model.py:12-122 (state [x, y, theta], input [v, w])."""
from typing import List, Optional, Tuple

import numpy as np

from .. import _lib
from ..global_parameters import K
from .base_model import BaseModel, ConstraintTables, SlackValue, straight_line_guess


class UnicycleModel(BaseModel):
    n_x = 3
    n_u = 2
    device_model_id = _lib.MODEL_UNICYCLE

    def __init__(
        self,
        r_init: np.ndarray = np.array([-8.0, -8.0, 0.0]),
        r_final: np.ndarray = np.array([8.0, 8.0, 0.0]),
        v_max: float = 1.0,
        w_max: float = np.pi / 6,
        bounds: Tuple[float, float] = (-10.0, 10.0),
        robot_radius: float = 0.5,
        obstacles: Optional[List[Tuple[List[float], float]]] = None,
    ):
        super().__init__()
        self.x_init = np.asarray(r_init, dtype=float).reshape(-1)
        self.x_final = np.asarray(r_final, dtype=float).reshape(-1)
        self.v_max = v_max
        self.w_max = w_max
        self.lower_bound, self.upper_bound = bounds
        self.robot_radius = robot_radius
        self.obstacles = (
            obstacles if obstacles is not None else [([5.0, 4.0], 3.0), ([-5.0, -4.0], 3.0), ([0.0, 0.0], 2.0)]
        )
        self.s_prime = [SlackValue(K) for _ in self.obstacles]

    # host-side twins of the device functions in csrc/common.cuh (unicycle_model.py:54-63)
    @staticmethod
    def f(x, u):
        return np.array([[u[0] * np.cos(x[2])], [u[0] * np.sin(x[2])], [u[1]]], dtype=float)

    @staticmethod
    def A(x, u):
        return np.array([[0.0, 0.0, -u[0] * np.sin(x[2])], [0.0, 0.0, u[0] * np.cos(x[2])], [0.0, 0.0, 0.0]])

    @staticmethod
    def B(x, u):
        return np.array([[np.cos(x[2]), 0.0], [np.sin(x[2]), 0.0], [0.0, 1.0]])

    def get_equations(self) -> Tuple:
        return self.f, self.A, self.B

    def initialize_trajectory(self, X: np.ndarray, U: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Straight line in state space, zero controls; fills and returns the caller's arrays."""
        return straight_line_guess(self.x_init, self.x_final, X, U)

    def get_constraints(self, X=None, U=None, X_ref=None, U_ref=None) -> ConstraintTables:
        """Boundary / input / box / obstacle tables (unicycle_model.py:85-115).  The obstacle normals
        themselves are recomputed on the device every iteration from X_ref (stage 2)."""
        M = len(self.obstacles)
        return ConstraintTables(
            x_init=self.x_init, x_final=self.x_final,
            pos_lo=self.lower_bound + self.robot_radius, pos_hi=self.upper_bound - self.robot_radius,
            v_max=self.v_max, w_max=self.w_max, input_kind="box",
            obs_centres=np.array([np.asarray(p, dtype=float).reshape(2) for p, _ in self.obstacles]).reshape(M, 2),
            obs_clearance=np.array([r + self.robot_radius for _, r in self.obstacles], dtype=float),
        )

    def get_objective(self, X=None, U=None, X_ref=None, U_ref=None):
        """1e5 * sum of obstacle slack (unicycle_model.py:117-122) -- returned as a descriptor; like the
        reference's SCProblem, the sub-problem does not use it."""
        return {"kind": "slack_sum", "weight": 1e5}
