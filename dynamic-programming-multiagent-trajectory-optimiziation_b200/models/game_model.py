"""GameUnicycleModel -- mirrors SCvx/models/game_model.py:11-126 (unicycle with per-agent Nash-game cost weights).

The reference keeps the slab normals in cvxpy Parameters and returns a cvxpy cost expression; here `z` holds the normals
as a (n_neighbours, 2, K) array (computed by slab_normals_kernel) and `get_cost_function` returns the cost DESCRIPTOR that
AgentBestResponse turns into the kernel's quadratic tables (include/scvx_b200.h, `quad_diag` / `lin_w` / `quad_pair`).
"""
from typing import List, Optional

import numpy as np
import torch

from .. import _device
from .unicycle_model import UnicycleModel


class GameUnicycleModel(UnicycleModel):
    def __init__(self, *, r_init: np.ndarray, r_final: np.ndarray, obstacles: Optional[List] = None,
                 control_weight: float = 1.0, collision_weight: float = 10.0, collision_radius: float = 0.50,
                 control_rate_weight: float = 5.0, curvature_weight: float = 100.0, inertia_weight: float = 0.0,
                 path_weight: float = 0.0, **kwargs):
        for key in ("control_weight", "collision_weight", "collision_radius", "control_rate_weight", "curvature_weight",
                    "inertia_weight", "path_weight"):
            kwargs.pop(key, None)
        super().__init__(r_init=r_init, r_final=r_final, obstacles=obstacles, **kwargs)
        self.control_weight = control_weight
        self.collision_weight = collision_weight
        self.collision_radius = collision_radius
        self.control_rate_weight = control_rate_weight
        self.curvature_weight = curvature_weight
        self.inertia_weight = inertia_weight
        self.path_weight = path_weight
        self.extra_constraints: List = []
        self.z_params: List[np.ndarray] = []      # per neighbour: (2, K) unit normals (zero where undefined)
        self.z_degenerate = 0                     # number of vanished normals after the last update

    def update_slabs(self, p_i: np.ndarray, neighbour_prev_pos: List[np.ndarray]):
        """z* = d/||d|| (zero when ||d|| < 1e-6), d = p_i - P_j, for every neighbour and knot (game_model.py:56-67)."""
        if not neighbour_prev_pos:
            self.z_params, self.z_degenerate = [], 0
            return
        dev = torch.device("cuda")
        K = np.asarray(p_i).shape[1]
        own = torch.zeros((1, 3, K), dtype=torch.float64, device=dev)
        own[0, 0:2] = torch.as_tensor(np.ascontiguousarray(p_i, dtype=np.float64)[0:2], device=dev)
        nbr = torch.zeros((len(neighbour_prev_pos), 3, K), dtype=torch.float64, device=dev)
        nbr[:, 0:2] = torch.as_tensor(np.stack([np.asarray(P, dtype=np.float64)[0:2] for P in neighbour_prev_pos]), device=dev)
        rad = torch.full((1,), float(self.collision_radius), dtype=torch.float64, device=dev)
        a, _, deg = _device.slab_normals(self.device_model_id, own, nbr, nbr, rad, i0=len(neighbour_prev_pos))
        z = a[0].cpu().numpy()
        self.z_params = [z[j] for j in range(z.shape[0])]
        self.z_degenerate = int(deg[0].item())

    @staticmethod
    def _horizon_of(first_neighbour, X_v, X_prev, neighbour_prev_pos):
        """Number of knots when the model has no obstacle slack to read it from: the shape of whatever trajectory-like
        argument is at hand.  The arguments may be `.shape`-carrying holders (AgentBestResponse passes those), not arrays,
        so np.asarray must not be applied to them."""
        candidates = [first_neighbour, X_v, X_prev] + list(neighbour_prev_pos or [])
        for c in candidates:
            shape = getattr(c, "shape", None)
            if shape is None and c is not None and not hasattr(c, "value"):
                shape = np.shape(c)
            if shape is not None and len(shape) == 2:
                return int(shape[1])
        raise ValueError("GameUnicycleModel.get_cost_function: cannot infer the number of knots from its arguments")

    def get_cost_function(self, X_v=None, U_v=None, neighbour_pos=None, X_prev=None, neighbour_prev_pos=None):  # noqa: ARG002
        """Cost descriptor (game_model.py:69-126): control effort, control-rate and curvature smoothing, inertia, path
        length; (re)initialises the slab normals to zero when the neighbour count changed, like the reference's lazy init."""
        n_nbr = 0 if neighbour_pos is None else len(neighbour_pos)
        K = self.s_prime[0].shape[0] if self.s_prime else None
        if not self.z_params or len(self.z_params) != n_nbr:
            Kz = K if K is not None else (self._horizon_of(neighbour_pos[0], X_v, X_prev, neighbour_prev_pos) if n_nbr else 0)
            self.z_params = [np.zeros((2, Kz)) for _ in range(n_nbr)]
        self.extra_constraints = [{"kind": "slab", "neighbour": j, "radius": self.collision_radius} for j in range(n_nbr)]
        return {"control_weight": self.control_weight, "control_rate_weight": self.control_rate_weight,
                "curvature_weight": self.curvature_weight, "inertia_weight": self.inertia_weight,
                "path_weight": getattr(self, "path_weight", 0.0)}
