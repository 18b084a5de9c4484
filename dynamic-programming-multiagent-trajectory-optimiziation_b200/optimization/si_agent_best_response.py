"""SI_AgentBestResponse -- mirrors SCvx/optimization/si_agent_best_response.py:15-124 (3-D single-integrator best response).

The sub-problem the reference ends up solving has NO slab and NO inter-sample rows (see models/game_si_model.py): it is the
SOCP of SCProblem + the quadratic game costs + sigma == sigma_ref; the slab normals and inter-sample linearisations are
still refreshed in `setup`, as the reference does."""
from typing import Dict, Tuple

import numpy as np
import torch

from ..discretization.first_order_hold import FirstOrderHold
from ..global_parameters import TRUST_RADIUS0, WEIGHT_NU, WEIGHT_SIGMA, WEIGHT_SLACK, K
from .agent_best_response import game_tables
from .sc_problem import SCProblem, _Holder


class SI_AgentBestResponse:
    _D = 3

    def __init__(self, i: int, multi_agent_model, K=K):
        self.i = i
        self.multi_model = multi_agent_model
        self.model = multi_agent_model.models[i]
        self.K = K
        self.foh = FirstOrderHold(self.model, K)
        self.Y_params: Dict[int, _Holder] = {j: _Holder((3, K)) for j in range(multi_agent_model.N) if j != i}
        self.X_prev_param = _Holder((3, K))
        self.scp = None

    def setup(self, X_ref, U_ref, sigma_ref: float, discr_mats: Tuple, neighbour_refs: Dict[int, np.ndarray], X_prev,
              neighbour_prev_refs: Dict[int, np.ndarray], tr_radius: float = TRUST_RADIUS0):
        self.scp = SCProblem(self.model, self.K)
        self.X_prev_param.value = np.array(X_prev)
        for j, P in self.Y_params.items():
            P.value = np.array(neighbour_refs[j][0:3, :])
        neighbour_prev_pos = [neighbour_prev_refs[j][0:3, :] for j in self.Y_params]
        self.model.get_cost_function(neighbour_pos=list(self.Y_params.values()), X_prev=self.X_prev_param,
                                     neighbour_prev_pos=neighbour_prev_pos)
        self.model.update_slabs(np.asarray(X_prev)[0:3, :], neighbour_prev_pos)
        self.model.update_intersample_constraints(X_v=None, U_v=None, X_nom=np.asarray(X_ref), U_nom=np.asarray(U_ref),
                                                  foh=self.foh, sigma_ref=sigma_ref)
        A_bar, B_bar, C_bar, S_bar, z_bar = discr_mats
        self.scp.set_parameters(
            A_bar=np.array(A_bar), B_bar=np.array(B_bar), C_bar=np.array(C_bar), S_bar=np.array(S_bar), z_bar=np.array(z_bar),
            X_ref=X_ref, U_ref=U_ref, sigma_ref=sigma_ref, weight_nu=WEIGHT_NU, weight_slack=WEIGHT_SLACK,
            weight_sigma=WEIGHT_SIGMA, tr_radius=tr_radius)

    def solve(self, solver: str = "ECOS", **solver_kwargs):  # noqa: ARG002
        if self.scp is None:
            raise RuntimeError("setup() must be called first")
        dev = self.scp._batch.device
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)   # noqa: E731
        qd, lw, qp, const = game_tables(self.model, self.X_prev_param.value, self.K, 3, 3)
        self.scp._solve_device(max_iter=int(solver_kwargs.get("max_iter", 0)), quad_diag=up(qd).unsqueeze(0),
                               lin_w=up(lw).unsqueeze(0), quad_pair=up(qp).unsqueeze(0), fix_sigma=True)
        if self.scp.prob.status not in ("optimal", "optimal_inaccurate"):
            raise RuntimeError(f"SCProblem for agent {self.i} was not solved successfully.")
        self.scp.prob.value += const
        X_i = self.scp.get_variable("X")
        U_i = self.scp.get_variable("U")
        nu_i = self.scp.get_variable("nu")
        p_i = X_i[0:3, :]
        slack_i = getattr(self.model, "get_linear_cost", lambda: 0.0)()
        return X_i, U_i, nu_i, slack_i, p_i
