"""MultiAgentModel -- the unicycle-agent collection with the reference's interface (SCvx/models/multi_agent_model.py:8-79):
`MultiAgentModel(agent_params, d_min)`, `.N`, `.models`, `.d_min`, `get_local_dynamics(i)`, `get_static_constraints(i, ...)`,
`linearize_collision(i, j, X_ref_i, X_ref_j) -> (A_ij (2, K), b_ij (K,))`."""
from ..global_parameters import K as GLOBAL_K
from .base_model import AgentCollection, SlackValue
from .unicycle_model import UnicycleModel


class MultiAgentModel(AgentCollection):
    _MODEL = UnicycleModel
    _KEYS = ("r_init", "r_final", "v_max", "w_max", "bounds", "robot_radius")

    def _build(self, params):
        model = super()._build(params)
        if params.get("obstacles") is not None:       # the reference assigns the list after construction and rebuilds the slacks
            model.obstacles = params["obstacles"]
            model.s_prime = [SlackValue(GLOBAL_K) for _ in model.obstacles]
        return model

    def linearize_collision(self, i, j, X_ref_i, X_ref_j):  # noqa: ARG002
        """a_k = (p_i,k - p_j,k)/(||.|| + 1e-6), b_k = d_min + a_k.p_j,k  (multi_agent_model.py:61-79), computed by
        scvx_linearize_collision_batched."""
        return self._pair(X_ref_i, X_ref_j)
