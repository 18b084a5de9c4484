"""SingleIntegratorModel -- mirrors SCvx/models/single_integrator_model.py:12-141 (3-D, f = u)."""
import numpy as np

from .. import _lib
from ..global_parameters import K
from .base_model import BaseModel, ConstraintTables, SlackValue, straight_line_guess

MARGIN_OBS = 0.0   # SCvx/config/SI_default_game.py:9


_DEFAULT_SPHERES = (([-5.0, -4.0, -5.0], 2.0), ([0.0, 0.0, 4.0], 2.0))   # the reference's two default obstacles


class SingleIntegratorModel(BaseModel):
    n_x = n_u = 3
    device_model_id = _lib.MODEL_SINGLE_INTEGRATOR

    def __init__(self, r_init=(-8.0, -8.0, -8.0), r_final=(8.0, 8.0, 8.0), v_max=1.0, bounds=(-10.0, 10.0),
                 robot_radius=0.5, obstacles=None):
        """Same keywords and defaults as single_integrator_model.py:17-51; the dynamics x' = u need no symbolic build."""
        super().__init__()
        self.x_init, self.x_final = (np.array(r, dtype=float).ravel() for r in (r_init, r_final))
        self.v_max, self.robot_radius = v_max, robot_radius
        self.lower_bound, self.upper_bound = bounds
        self.obstacles = [(list(c), r) for c, r in _DEFAULT_SPHERES] if obstacles is None else obstacles
        self.s_prime = [SlackValue(K) for _ in self.obstacles]
        eye, zero = np.eye(3), np.zeros((3, 3))
        self.f, self.A, self.B = (lambda x, u: u), (lambda x, u: zero.copy()), (lambda x, u: eye.copy())

    def get_equations(self):
        return self.f, self.A, self.B

    def initialize_trajectory(self, X, U):
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
