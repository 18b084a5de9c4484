"""Model plug-in contract -- mirrors SCvx/models/base_model.py:7-88.

Differences forced by the GPU path (no cvxpy objects exist here):
  * the shipped models name their compiled dynamics with the class attribute `device_model_id` (one of
    scvx_b200._lib.MODEL_*; device functions in csrc/common.cuh).  ANY OTHER model runs on the device by returning its
    symbolic right-hand side from `symbolic_dynamics()`: f, A = df/dx, B = df/du are generated as CUDA C++ and compiled
    with NVRTC at first use (scvx_b200/codegen.py).  `get_equations()` still returns host callables with the
    reference's shapes so user code that probes f, A, B keeps working.
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
    position_dim = 2          # the first `position_dim` states are the position (obstacle / box / inter-agent rows act on them)

    def __init__(self):
        super().__init__()

    def symbolic_dynamics(self):
        """Optional hook of the GPU path: return (x_symbols, u_symbols, f_expr) as sympy objects -- the expression a model
        written for the reference already holds when it builds its lambdas (unicycle_model.py:54-63) -- and the model's
        dynamics are compiled for the device at first use (scvx_b200/codegen.py, NVRTC).  None: no device dynamics."""
        return None

    @property
    def device_model_id(self):
        """Id of the compiled device dynamics.  The shipped models override this with a constant; any other model gets one
        by returning its symbolic right-hand side from `symbolic_dynamics()`."""
        sym = self.symbolic_dynamics()
        if sym is None:
            return None
        from .. import codegen
        x_syms, u_syms, f_expr = sym
        return codegen.register(x_syms, u_syms, f_expr, self.position_dim)

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


def straight_line_guess(x_init, x_final, X, U):
    """The reference models' `initialize_trajectory` (unicycle_model.py:73-83, single_integrator_model.py:65-77): states on
    the straight line from x_init to x_final, zero inputs, written into the caller's arrays.  Column k is
    (K-1-k)/(K-1) * x_init + k/(K-1) * x_final -- evaluated for all k at once with the same two products and one sum per
    entry, hence bit-identical to the reference's loop."""
    n = X.shape[1]
    k = np.arange(n)
    X[:, :] = np.outer(x_init, (n - 1 - k) / (n - 1)) + np.outer(x_final, k / (n - 1))
    U[:] = 0
    return X, U


class AgentCollection:
    """What MultiAgentModel / SI_MultiAgentModel share (multi_agent_model.py:8-79, SI_multi_agent_model.py:7-74): a list of
    per-agent models built from parameter dicts, the minimum separation, and per-agent accessors."""
    _MODEL = None
    _KEYS = ()

    def __init__(self, agent_params, d_min=1.0):
        self.N, self.d_min = len(agent_params), d_min
        self.models = [self._build(params) for params in agent_params]

    def _build(self, params):
        return self._MODEL(**{k: params[k] for k in self._KEYS if params.get(k) is not None})

    def get_local_dynamics(self, i):
        return self.models[i].get_equations()

    def get_static_constraints(self, i, X=None, U=None, X_ref=None, U_ref=None):
        return self.models[i].get_constraints(X, U, X_ref, U_ref)

    def get_objective(self, i, X=None, U=None, X_ref=None, U_ref=None):
        return self.models[i].get_objective(X, U, X_ref, U_ref)

    def _pair(self, X_ref_i, X_ref_j):
        """One (i, j) pair through the batched stage-2 kernel (n_local = n_agents = 1): (a (d, K), b (K,))."""
        import torch
        from .. import _device
        dev = torch.device("cuda")
        Xi = torch.as_tensor(np.ascontiguousarray(X_ref_i, dtype=np.float64)).unsqueeze(0).to(dev)
        Xj = torch.as_tensor(np.ascontiguousarray(X_ref_j, dtype=np.float64)).unsqueeze(0).to(dev)
        a, b = _device.linearize_collision(self._MODEL.device_model_id, Xi, Xj, self.d_min, i0=1)   # i0=1: slot 0 is not "self"
        return a[0, 0].cpu().numpy(), b[0, 0].cpu().numpy()
