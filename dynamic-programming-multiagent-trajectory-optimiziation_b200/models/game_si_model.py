"""GameSIModel -- mirrors SCvx/models/game_si_model.py:11-187 (3-D single integrator with Nash-game cost weights).

What the reference's class effectively does is kept, quirks included (they change the sub-problem):
  * `get_cost_function` adds one slab row per neighbour and knot WITHOUT the slack it declares (the collision slacks only
    appear in the cost, so they are zero at the optimum), game_si_model.py:115-135;
  * `update_intersample_constraints` then RESETS `extra_constraints` (the slab rows are gone) and computes the inter-sample
    linearisations but never appends them as constraints (game_si_model.py:150-187: `expr` is built and dropped).
So the best response the reference solves is: base problem + quadratic costs + sigma == sigma_ref.  The linearisations are
still computed here (one launch for all segments and obstacles) and exposed as `inter_samples` for callers that want them.
"""
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import _device
from .single_integrator_model import SingleIntegratorModel

GLOBAL_COLL_RAD = 2 * 0.5 + 0.0     # AGT_COLL_RAD of SCvx/config/SI_default_game.py:22 (2 * ROBOT_RADIUS + MARGIN_AGT)


class GameSIModel(SingleIntegratorModel):
    _COST_KEYS = {"control_weight", "collision_weight", "control_rate_weight", "curvature_weight", "inertia_weight", "path_weight"}

    def __init__(self, *, r_init: np.ndarray, r_final: np.ndarray, robot_radius: float = 0.5,
                 collision_radius: Optional[float] = None, obstacles: Optional[List[Tuple[List[float], float]]] = None, **kwargs):
        self.agent_coll_rad = collision_radius if collision_radius is not None else GLOBAL_COLL_RAD
        self.control_weight = kwargs.pop("control_weight", 1.0)
        self.collision_weight = kwargs.pop("collision_weight", 80.0)
        self.control_rate_weight = kwargs.pop("control_rate_weight", 5.0)
        self.curvature_weight = kwargs.pop("curvature_weight", 0.0)
        self.inertia_weight = kwargs.pop("inertia_weight", 0.0)
        self.path_weight = kwargs.pop("path_weight", 0.0)
        for key in list(kwargs):
            if key in self._COST_KEYS:
                kwargs.pop(key)
        super().__init__(r_init=r_init, r_final=r_final, robot_radius=robot_radius, obstacles=obstacles)
        self.collision_radius = self.agent_coll_rad
        self.coll_slacks: List = []
        self.z_params: List[np.ndarray] = []
        self.z_degenerate = 0
        self.inter_slacks: List = []
        self.inter_samples: List[dict] = []
        self.extra_constraints: List = []

    def update_slabs(self, p_i: np.ndarray, neighbour_prev_pos: List[np.ndarray]) -> None:
        """z* = (x_i - x_j)/||x_i - x_j|| per neighbour and knot, zero below 1e-6 (game_si_model.py:69-88)."""
        if not neighbour_prev_pos:
            self.z_params, self.z_degenerate = [], 0
            return
        dev = torch.device("cuda")
        own = torch.as_tensor(np.ascontiguousarray(p_i, dtype=np.float64)[None, 0:3], device=dev)
        nbr = torch.as_tensor(np.stack([np.asarray(P, dtype=np.float64)[0:3] for P in neighbour_prev_pos]), device=dev)
        rad = torch.full((1,), float(self.agent_coll_rad), dtype=torch.float64, device=dev)
        a, _, deg = _device.slab_normals(self.device_model_id, own, nbr, nbr, rad, i0=len(neighbour_prev_pos))
        z = a[0].cpu().numpy()
        self.z_params = [z[j] for j in range(z.shape[0])]
        self.z_degenerate = int(deg[0].item())

    def get_cost_function(self, X_v=None, U_v=None, neighbour_pos=None, X_prev=None, neighbour_prev_pos=None):  # noqa: ARG002
        """Cost descriptor (game_si_model.py:90-137); refreshes the normals from (X_prev, neighbour_prev_pos) and lists the
        slab rows -- which update_intersample_constraints will drop again, as in the reference."""
        n_nbr = 0 if neighbour_pos is None else len(neighbour_pos)
        self.coll_slacks = [{"kind": "coll_slack", "neighbour": j} for j in range(n_nbr)]
        if X_prev is not None and neighbour_prev_pos:
            Xp = X_prev.value if hasattr(X_prev, "value") else X_prev
            self.update_slabs(np.asarray(Xp), neighbour_prev_pos)
        self.extra_constraints = [{"kind": "slab", "neighbour": j, "radius": self.agent_coll_rad} for j in range(n_nbr)]
        return {"control_weight": self.control_weight, "control_rate_weight": self.control_rate_weight, "curvature_weight": 0.0,
                "inertia_weight": self.inertia_weight, "path_weight": self.path_weight}

    def update_intersample_constraints(self, X_v, U_v, X_nom: np.ndarray, U_nom: np.ndarray, foh, sigma_ref: float) -> None:  # noqa: ARG002
        """Linearised inter-sample obstacle clearances about (X_nom, U_nom) with sigma = 1 and t in (0, 1), exactly the calls of
        game_si_model.py:139-187 -- all (segment, obstacle) pairs in one launch.  Resets `extra_constraints`, like the reference."""
        self.extra_constraints = []
        self.inter_slacks = []
        self.inter_samples = []
        if not self.obstacles:
            return
        dev = torch.device("cuda")
        X = torch.as_tensor(np.ascontiguousarray(X_nom, dtype=np.float64)[None], device=dev)
        U = torch.as_tensor(np.ascontiguousarray(U_nom, dtype=np.float64)[None], device=dev)
        C = torch.as_tensor(np.array([np.asarray(c, dtype=float).reshape(3) for c, _ in self.obstacles])[None], device=dev)
        R = torch.as_tensor(np.array([float(r) for _, r in self.obstacles])[None], device=dev)
        one = torch.ones(1, dtype=torch.float64, device=dev)
        nr, ts, h0, gx = _device.intersample(self.device_model_id, X, U, one, C, R, proj_dim=3, t_range=1.0)
        nr, ts, h0, gx = nr[0].cpu().numpy(), ts[0].cpu().numpy(), h0[0].cpu().numpy(), gx[0].cpu().numpy()
        for k in range(nr.shape[0]):
            for j in range(nr.shape[1]):
                for q in range(min(int(nr[k, j]), ts.shape[2])):
                    self.inter_samples.append({"k": k, "obs": j, "t_star": float(ts[k, j, q]), "h0": float(h0[k, j, q]),
                                               "grad_x": gx[k, j, q].copy(), "grad_u": np.zeros(self.n_u)})
                    self.inter_slacks.append({"kind": "inter_slack", "name": f"s_intersample_k{k}_obs{j}_t{int(ts[k, j, q] * 1e3)}"})
