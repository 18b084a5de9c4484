"""SI_MultiAgentModel -- mirrors SCvx/models/SI_multi_agent_model.py:7-74 (3-D single-integrator agents)."""
from .multi_agent_model import _pair_linearize
from .single_integrator_model import SingleIntegratorModel


class SI_MultiAgentModel:  # noqa: N801
    _ALLOWED_KEYS = {"r_init", "r_final", "v_max", "bounds", "robot_radius", "obstacles"}

    def __init__(self, agent_params: list, d_min: float = 1.0):
        self.N = len(agent_params)
        self.models = []
        for params in agent_params:
            basic = {k: v for k, v in params.items() if k in self._ALLOWED_KEYS}
            self.models.append(SingleIntegratorModel(**basic))
        self.d_min = d_min

    def get_local_dynamics(self, i: int):
        return self.models[i].get_equations()

    def get_static_constraints(self, i: int, X=None, U=None, X_ref=None, U_ref=None):
        return self.models[i].get_constraints(X, U, X_ref, U_ref)

    def get_objective(self, i: int, X=None, U=None, X_ref=None, U_ref=None):
        return self.models[i].get_objective(X, U, X_ref, U_ref)

    def linearize_inter_agent_collision(self, i: int, j: int, X_ref_i, X_ref_j) -> tuple:  # noqa: ARG002
        """SI_multi_agent_model.py:49-74, computed by scvx_linearize_collision_batched."""
        return _pair_linearize(SingleIntegratorModel.device_model_id, self.d_min, X_ref_i, X_ref_j)
