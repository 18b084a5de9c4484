"""SI_AgentBestResponse -- one 3-D single-integrator agent's best response with the reference's interface
(SCvx/optimization/si_agent_best_response.py:15-124).

The sub-problem the reference ends up solving has NO slab and NO inter-sample rows (see models/game_si_model.py): it is the
SOCP of SCProblem + the quadratic game costs + sigma == sigma_ref; the slab normals and inter-sample linearisations are
still refreshed in `setup`, as the reference does.  Everything else is the unicycle response (agent_best_response.py) with
three position rows and three inputs."""
import numpy as np
import torch

from .agent_best_response import AgentBestResponse, game_tables


class SI_AgentBestResponse(AgentBestResponse):  # noqa: N801
    _D = 3

    def _after_slabs(self, X_ref, U_ref, sigma_ref):
        self.model.update_intersample_constraints(X_v=None, U_v=None, X_nom=np.asarray(X_ref), U_nom=np.asarray(U_ref),
                                                  foh=self.foh, sigma_ref=sigma_ref)

    def solve(self, solver: str = "ECOS", **solver_kwargs):  # noqa: ARG002
        if self.scp is None:
            raise RuntimeError("setup() must be called first")
        dev = self.scp._batch.device
        tables = game_tables(self.model, self.X_prev_param.value, self.K, 3, 3)
        qd, lw, qp = (torch.as_tensor(np.ascontiguousarray(t, dtype=np.float64)).to(dev).unsqueeze(0) for t in tables[:3])
        self.scp._solve_device(max_iter=int(solver_kwargs.get("max_iter", 0)), quad_diag=qd, lin_w=lw, quad_pair=qp, fix_sigma=True)
        if self.scp.prob.status not in ("optimal", "optimal_inaccurate"):
            raise RuntimeError(f"SCProblem for agent {self.i} was not solved successfully.")
        self.scp.prob.value += tables[3]
        X_i, U_i, nu_i = (self.scp.get_variable(name) for name in ("X", "U", "nu"))
        extra = getattr(self.model, "get_linear_cost", None)
        return X_i, U_i, nu_i, (extra() if extra else 0.0), X_i[0:3, :]
