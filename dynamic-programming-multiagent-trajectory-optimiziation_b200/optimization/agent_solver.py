"""AgentSolver -- mirrors SCvx/optimization/agent_solver.py:10-117 (per-agent ADMM sub-problem).

`setup` records the numbers the reference would bind into a fresh cvxpy problem (base SCProblem
parameters with tr_radius = TRUST_RADIUS0, neighbour references); `solve` linearises the inter-agent
rows on the device about (X_ref_i, X_ref_j), collapses the augmented-Lagrangian terms to
rho/2 (N-1) |P|^2 + <sum_j Lambda_j - rho sum_j Y_j, P> and runs the batched interior-point kernel.
"""
import numpy as np
import torch

from .. import _device
from ..global_parameters import TRUST_RADIUS0, WEIGHT_NU, WEIGHT_SIGMA, WEIGHT_SLACK, K
from .admm_utils import WEIGHT_COLLISION_SLACK
from .sc_problem import SCProblem, _Holder


class AgentSolver:
    _D = 2   # position dimension of the wrapped model family

    def __init__(self, agent_index: int, multi_agent_model, rho_admm: float, K=K):
        self.i = agent_index
        self.multi_agent_model = multi_agent_model
        self.model_i = multi_agent_model.models[self.i]
        self.K = K
        self.d_min = multi_agent_model.d_min
        self.scp = SCProblem(self.model_i, K)
        self.rho_admm = rho_admm
        self.Y, self.Lambda, self.S = {}, {}, {}
        for j in range(multi_agent_model.N):
            if j == self.i:
                continue
            self.Y[j] = _Holder((self._D, K))
            self.Lambda[j] = _Holder((self._D, K))
            self.S[j] = _Holder((K, 1))
        self._neighbor_refs = None
        self.prob = self.scp.prob

    def setup(self, X_ref_i, U_ref_i, sigma_ref_i, discretization_mats, neighbor_refs: dict):
        A_bar, B_bar, C_bar, S_bar, z_bar = discretization_mats
        self.scp.set_parameters(
            A_bar=np.array(A_bar), B_bar=np.array(B_bar), C_bar=np.array(C_bar), S_bar=np.array(S_bar), z_bar=np.array(z_bar),
            X_ref=X_ref_i, U_ref=U_ref_i, sigma_ref=sigma_ref_i,
            weight_nu=WEIGHT_NU, weight_slack=WEIGHT_SLACK, weight_sigma=WEIGHT_SIGMA, tr_radius=TRUST_RADIUS0,
        )
        self._neighbor_refs = dict(neighbor_refs)

    def solve(self, **kwargs):
        """Returns X_i, U_i, nu_i, slacks, p_i (agent_solver.py:104-117)."""
        if self._neighbor_refs is None:
            raise RuntimeError("call setup() before solve()")
        d, K = self._D, self.K
        dev = self.scp._batch.device
        js = list(self._neighbor_refs.keys())
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)   # noqa: E731
        X_own = up(np.asarray(self.scp.par["X_ref"].value)).unsqueeze(0)
        col_a = col_b = quad = lin = None
        if js:
            X_nbr = up(np.stack([self._neighbor_refs[j] for j in js]))
            Y = up(np.stack([self.Y[j].value for j in js]))
            Lam = up(np.stack([self.Lambda[j].value for j in js]))
            # i0 = len(js): no slot is "self"
            col_a, _ = _device.linearize_collision(self.scp._batch.model_id, X_own, X_nbr, self.d_min, i0=len(js))
            col_b = self.d_min + (col_a * Y[None]).sum(dim=2)
            quad = torch.full((1,), self.rho_admm * len(js), dtype=torch.float64, device=dev)
            lin = (Lam.sum(dim=0) - self.rho_admm * Y.sum(dim=0)).unsqueeze(0).contiguous()
        ws = self.scp._solve_device(max_iter=int(kwargs.get("max_iter", 0)), col_a=col_a, col_b=col_b, quad_rho=quad,
                                    lin_p=lin, weight_col=WEIGHT_COLLISION_SLACK)
        if js:
            cs = ws.col_slack[0].cpu().numpy()
            for q, j in enumerate(js):
                self.S[j].value = cs[q].reshape(K, 1).copy()
            const = 0.5 * self.rho_admm * float((Y * Y).sum().item()) - float((Lam * Y).sum().item())
            self.scp.prob.value += const            # full augmented-Lagrangian objective (agent_solver.py:92-98)
        X_i = self.scp.get_variable("X")
        U_i = self.scp.get_variable("U")
        nu_i = self.scp.get_variable("nu")
        slacks = {j: self.S[j].value for j in self.S}
        p_i = X_i[0:d, :]
        return X_i, U_i, nu_i, slacks, p_i
