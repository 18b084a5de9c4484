"""SI_MultiAgentModel -- the 3-D single-integrator agent collection with the reference's interface
(SCvx/models/SI_multi_agent_model.py:7-74); `linearize_inter_agent_collision(i, j, X_ref_i, X_ref_j) -> (A_ij (3, K), b_ij (K,))`."""
from .base_model import AgentCollection
from .single_integrator_model import SingleIntegratorModel


class SI_MultiAgentModel(AgentCollection):  # noqa: N801
    _MODEL = SingleIntegratorModel
    _KEYS = ("r_init", "r_final", "v_max", "bounds", "robot_radius", "obstacles")

    def _build(self, params):
        # unknown keys are dropped, known ones passed through even when None (the reference forwards them as they are)
        return self._MODEL(**{k: v for k, v in params.items() if k in self._KEYS})

    def linearize_inter_agent_collision(self, i: int, j: int, X_ref_i, X_ref_j) -> tuple:  # noqa: ARG002
        """SI_multi_agent_model.py:49-74, computed by scvx_linearize_collision_batched."""
        return self._pair(X_ref_i, X_ref_j)
