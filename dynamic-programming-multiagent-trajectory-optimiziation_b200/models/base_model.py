"""Model plug-in contract -- mirrors SCvx/models/base_model.py:7-88.

Differences forced by the GPU path (no cvxpy objects exist here):
  * a model that wants to run on the device names its compiled dynamics with the class attribute
    `device_model_id` (one of scvx_b200._lib.MODEL_*); the device functions f/A/B live in
    csrc/common.cuh.  `get_equations()` still returns host callables with the reference's shapes
    so user code that probes f, A, B keeps working.
  * `get_constraints` returns a `ConstraintTables` descriptor (plain arrays) instead of a list of
    cvxpy constraints; the sub-problem kernel consumes exactly these tables.
"""
from abc import ABC, abstractmethod
from dataclasses import dataclass, field

import numpy as np


class SlackValue:
    """Stand-in for the reference's `cvx.Variable((K, 1), nonneg=True)` obstacle slack: after a solve,
    `.value` holds the (K, 1) slack column, which is all callers read (scvx_solver.py:117-123)."""

    def __init__(self, K):
        self.shape = (K, 1)
        self.value = None


@dataclass
class ConstraintTables:
    """Everything Model.get_constraints contributes to the sub-problem, as numbers
    (unicycle_model.py:85-115 / single_integrator_model.py:79-128)."""
    x_init: np.ndarray
    x_final: np.ndarray
    pos_lo: float
    pos_hi: float
    v_max: float
    w_max: float
    input_kind: str                      # "box" (0<=v<=v_max, |w|<=w_max) or "ball" (||u||_2<=v_max)
    obs_centres: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))   # (M, d)
    obs_clearance: np.ndarray = field(default_factory=lambda: np.zeros(0))      # (M,) r_j + r_rob (+margin)


class BaseModel(ABC):
    n_x = 0
    n_u = 0
    device_model_id = None

    def __init__(self):
        super().__init__()

    @abstractmethod
    def get_equations(self):
        """Returns f(x,u), A(x,u)=df/dx, B(x,u)=df/du as host callables."""

    @abstractmethod
    def get_constraints(self, X=None, U=None, X_ref=None, U_ref=None):
        """Returns a ConstraintTables descriptor (arguments kept for signature parity)."""

    @abstractmethod
    def get_objective(self, X=None, U=None, X_ref=None, U_ref=None):
        """Model-specific cost descriptor (unused by SCProblem, as in the reference: sc_problem.py:77-82)."""

    @abstractmethod
    def initialize_trajectory(self, X: np.ndarray, U: np.ndarray):
        """Provide the initial guess (modified in place); returns (X, U)."""

    # optional scaling hooks, identity by default (base_model.py:66-88)
    def nondimensionalize(self):
        return

    def redimensionalize(self):
        return

    def x_nondim(self, x):
        return x

    def u_nondim(self, u):
        return u

    def x_redim(self, X):
        return X

    def u_redim(self, U):
        return U
