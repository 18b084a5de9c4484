"""AgentBestResponse -- mirrors SCvx/optimization/agent_best_response.py:15-113 (one agent's pure-Nash best response).

Sub-problem = SCProblem + the GameUnicycleModel cost + slab rows z_jk.(p_ik - Y_jk) >= collision_radius + sigma == sigma_ref.
On the device the cost becomes the kernel's quadratic tables, sigma is frozen, and the HARD slab rows are carried as hinge
rows with an exact penalty (weights SLAB_PENALTIES, escalated only while a slab slack remains): when the hard-constrained problem is feasible
the minimisers coincide and every slab slack is zero; when it is not -- the reference's ECOS reports infeasibility -- the
residual slack is detected and RuntimeError is raised, as the reference does.
"""
from typing import Dict, Tuple

import numpy as np
import torch

from ..discretization.first_order_hold import FirstOrderHold
from ..global_parameters import TRUST_RADIUS0, WEIGHT_NU, WEIGHT_SIGMA, WEIGHT_SLACK, K
from .sc_problem import SCProblem, _Holder

# Exact-penalty weights of the hard slab rows, tried in order until every slab slack vanishes.  The weight must dominate
# WEIGHT_SLACK = 1e6 of the SOFT obstacle rows (at 1e6 a slab that conflicts with an obstacle keeps a residual slack); 1e7
# converges like the plain problem (tools/robustness_scan_game.py: 128 agents x 127 slots, 17 IPM iterations, all optimal),
# 1e8 is exact in tighter conflicts but needs 2-4x the iterations, so it is only the fall-back.
SLAB_PENALTIES = (1e7, 1e8)
SLAB_PENALTY = SLAB_PENALTIES[0]
SLAB_TOL = 1e-7             # a larger residual slack means the hard rows are infeasible


def game_tables(model, X_prev, K, n_x=3, n_u=2):
    """Weights of GameUnicycleModel.get_cost_function as the kernel's tables (include/scvx_b200.h):
    quad_diag (n_x+n_u,), lin_w (n_x+n_u, K), quad_pair (n_x+n_u,), constant."""
    cost = model.get_cost_function()
    if cost.get("path_weight", 0.0) > 0:
        raise NotImplementedError("path_weight > 0 (sum of segment norms) is not supported by the device sub-problem")
    ns = n_x + n_u
    qd = np.zeros(ns); qp = np.zeros(ns); lw = np.zeros((ns, K)); const = 0.0
    qd[n_x:] += 2.0 * cost["control_weight"]                      # c ||U||^2
    if cost["control_rate_weight"] > 0:
        qp[n_x:] += 2.0 * cost["control_rate_weight"]             # c sum ||u_{k+1} - u_k||^2
    if cost["curvature_weight"] > 0:
        qp[2] += 2.0 * cost["curvature_weight"]                   # c sum (theta_{k+1} - theta_k)^2
    if cost["inertia_weight"] > 0:
        w = cost["inertia_weight"]                                # c ||X - X_prev||^2
        qd[:n_x] += 2.0 * w
        lw[:n_x] += -2.0 * w * np.asarray(X_prev, dtype=float)
        const += w * float((np.asarray(X_prev, dtype=float) ** 2).sum())
    return qd, lw, qp, const


class AgentBestResponse:
    _D = 2

    def __init__(self, i: int, multi_agent_model, K=K):
        self.i = i
        self.multi_model = multi_agent_model
        self.model = multi_agent_model.models[i]
        self.K = K
        self.foh = FirstOrderHold(self.model, K)
        self.Y_params: Dict[int, _Holder] = {j: _Holder((self._D, K)) for j in range(multi_agent_model.N) if j != i}
        self.X_prev_param = _Holder((3, K))
        self.scp = None
    def _after_slabs(self, X_ref, U_ref, sigma_ref):
        """Hook for model-specific refreshes that follow the slab update (the single-integrator game's inter-sample rows)."""

    def setup(self, X_ref, U_ref, sigma_ref: float, discr_mats: Tuple, neighbour_refs: Dict[int, np.ndarray], X_prev,
              neighbour_prev_refs: Dict[int, np.ndarray], tr_radius: float = TRUST_RADIUS0) -> None:
        d = self._D
        self.scp = SCProblem(self.model, self.K)
        self.X_prev_param.value = np.array(X_prev)
        for j, P in self.Y_params.items():
            P.value = np.array(neighbour_refs[j][0:d, :])
        neighbour_prev_pos = [neighbour_prev_refs[j][0:d, :] for j in self.Y_params]
        self.model.get_cost_function(neighbour_pos=list(self.Y_params.values()), X_prev=self.X_prev_param,
                                     neighbour_prev_pos=neighbour_prev_pos)
        # normals along X_prev -> neighbour_prev so that the slabs hold at the previous iterate (agent_best_response.py:66-72)
        self.model.update_slabs(np.asarray(X_prev)[0:d, :], neighbour_prev_pos)
        self._after_slabs(X_ref, U_ref, sigma_ref)
        A_bar, B_bar, C_bar, S_bar, z_bar = discr_mats
        self.scp.set_parameters(
            A_bar=np.array(A_bar), B_bar=np.array(B_bar), C_bar=np.array(C_bar), S_bar=np.array(S_bar), z_bar=np.array(z_bar),
            X_ref=X_ref, U_ref=U_ref, sigma_ref=sigma_ref, weight_nu=WEIGHT_NU, weight_slack=WEIGHT_SLACK,
            weight_sigma=WEIGHT_SIGMA, tr_radius=tr_radius)

    def solve(self, solver: str = "ECOS", **solver_kwargs):  # noqa: ARG002
        if self.scp is None:
            raise RuntimeError("call setup() before solve()")
        d, K = self._D, self.K
        dev = self.scp._batch.device
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)   # noqa: E731
        js = list(self.Y_params.keys())
        col_a = col_b = None
        if js:
            if self.model.z_degenerate:
                # a vanished normal makes the reference's row 0 >= collision_radius infeasible
                raise RuntimeError("SCProblem error inside AgentBestResponse")
            z = np.stack(self.model.z_params)                                  # (n_nbr, d, K)
            Y = np.stack([self.Y_params[j].value for j in js])                 # (n_nbr, d, K)
            col_a = up(z).unsqueeze(0)
            col_b = up(self.model.collision_radius + (z * Y).sum(axis=1)).unsqueeze(0)
        qd, lw, qp, const = game_tables(self.model, self.X_prev_param.value, K)
        for penalty in (SLAB_PENALTIES if js else SLAB_PENALTIES[:1]):
            ws = self.scp._solve_device(max_iter=int(solver_kwargs.get("max_iter", 0)), col_a=col_a, col_b=col_b,
                                        weight_col=penalty, quad_diag=up(qd).unsqueeze(0), lin_w=up(lw).unsqueeze(0),
                                        quad_pair=up(qp).unsqueeze(0), fix_sigma=True)
            if self.scp.status == 2:
                raise RuntimeError("SCProblem error inside AgentBestResponse")
            if not js or float(ws.col_slack[0].max().item()) <= SLAB_TOL:
                break
        else:
            raise RuntimeError("SCProblem error inside AgentBestResponse")      # hard slab rows infeasible
        self.scp.prob.value += const
        X_i = self.scp.get_variable("X")
        U_i = self.scp.get_variable("U")
        nu_i = self.scp.get_variable("nu")
        p_i = X_i[0:d, :]
        slack_i = getattr(self.model, "get_linear_cost", lambda: 0.0)()
        return X_i, U_i, nu_i, slack_i, p_i
