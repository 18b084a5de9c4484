"""MultiAgentModel -- mirrors SCvx/models/multi_agent_model.py:8-79 (wrapper over UnicycleModel agents)."""
import numpy as np
import torch

from .. import _device
from ..global_parameters import K as GLOBAL_K
from .base_model import SlackValue
from .unicycle_model import UnicycleModel


def _pair_linearize(model_id, d_min, X_ref_i, X_ref_j):
    """One (i, j) pair through the batched stage-2 kernel (n_local = n_agents = 1)."""
    dev = torch.device("cuda")
    Xi = torch.as_tensor(np.ascontiguousarray(X_ref_i, dtype=np.float64)).unsqueeze(0).to(dev)
    Xj = torch.as_tensor(np.ascontiguousarray(X_ref_j, dtype=np.float64)).unsqueeze(0).to(dev)
    a, b = _device.linearize_collision(model_id, Xi, Xj, d_min, i0=1)   # i0=1: slot 0 is not "self"
    return a[0, 0].cpu().numpy(), b[0, 0].cpu().numpy()


class MultiAgentModel:
    def __init__(self, agent_params, d_min=1.0):
        self.N = len(agent_params)
        self.models = []
        for params in agent_params:
            kwargs = {}
            if "r_init" in params:
                kwargs["r_init"] = params["r_init"]
            if "r_final" in params:
                kwargs["r_final"] = params["r_final"]
            for key in ("v_max", "w_max", "bounds", "robot_radius"):
                if key in params and params[key] is not None:
                    kwargs[key] = params[key]
            m = UnicycleModel(**kwargs)
            if "obstacles" in params and params["obstacles"] is not None:
                m.obstacles = params["obstacles"]
                m.s_prime = [SlackValue(GLOBAL_K) for _ in m.obstacles]
            self.models.append(m)
        self.d_min = d_min

    def get_local_dynamics(self, i):
        return self.models[i].get_equations()

    def get_static_constraints(self, i, X=None, U=None, X_ref=None, U_ref=None):
        return self.models[i].get_constraints(X, U, X_ref, U_ref)

    def linearize_collision(self, i, j, X_ref_i, X_ref_j):  # noqa: ARG002
        """a_k = (p_i,k - p_j,k)/(||.|| + 1e-6), b_k = d_min + a_k.p_j,k  (multi_agent_model.py:61-79),
        computed by scvx_linearize_collision_batched."""
        return _pair_linearize(UnicycleModel.device_model_id, self.d_min, X_ref_i, X_ref_j)
