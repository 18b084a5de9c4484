"""ADMMCoordinator -- mirrors SCvx/optimization/admm_coordinator.py:13-118.

`solve` is the reference's sequential Gauss-Seidel sweep, statement for statement, with the per-agent
discretisation and sub-problem on the GPU.  `solve_batched` is the device-resident Jacobi variant
(`scvx_b200.batch.BatchedADMM`): all agents of a round in one launch, shardable over GPUs with one
all-gather per round (SURVEY hard part 5 documents the Gauss-Seidel -> Jacobi difference).
"""
import time

import numpy as np
import torch

from ..batch import BatchedADMM
from ..discretization.first_order_hold import FirstOrderHold
from ..global_parameters import K
from ..utils.multi_agent_logging import print_iteration, print_summary
from .admm_utils import dual_residual, primal_residual
from .agent_solver import AgentSolver


class ADMMCoordinator:
    _SOLVER = AgentSolver
    _D = 2
    _SI_VARIANT = False

    def __init__(self, multi_agent_model, rho_admm: float = 1.0, max_iter: int = 10, K=K):
        self.model = multi_agent_model
        self.N = multi_agent_model.N
        self.rho_admm = rho_admm
        self.max_iter = max_iter
        self.K = K
        self.agent_solvers = []
        self.discretizers = []
        for i in range(self.N):
            self.agent_solvers.append(self._SOLVER(i, multi_agent_model, rho_admm, K))
            self.discretizers.append(FirstOrderHold(multi_agent_model.models[i], K))

    def solve(self, X_refs: list, U_refs: list, sigma_ref: float, verbose: bool = True):
        d, K = self._D, self.K
        for solver in self.agent_solvers:
            for j in solver.Y:
                solver.Y[j].value = np.array(X_refs[j][0:d, :])
                solver.Lambda[j].value = np.zeros((d, K))
        X_curr = list(X_refs)
        U_curr = list(U_refs)
        primal_hist, dual_hist = [], []
        t0 = time.time()
        for it in range(self.max_iter):
            new_positions = [None] * self.N
            for i, solver in enumerate(self.agent_solvers):
                mats = self.discretizers[i].calculate_discretization(X_curr[i], U_curr[i], sigma_ref)
                neighbor_refs = {j: X_curr[j] for j in range(self.N) if j != i}
                if self._SI_VARIANT:     # si_admm_coordinator.py:80-86: the INITIAL references every round
                    solver.setup(X_refs[i], U_refs[i], sigma_ref, mats, neighbor_refs)
                else:
                    solver.setup(X_curr[i], U_curr[i], sigma_ref, mats, neighbor_refs)
                X_i, U_i, nu_i, slacks_i, p_i = solver.solve(solver="ECOS")
                X_curr[i], U_curr[i] = X_i, U_i
                new_positions[i] = p_i
            pr_vals, du_vals = [], []
            for solver in self.agent_solvers:
                for j in solver.Y:
                    p_j = new_positions[j]
                    Y_old = solver.Y[j].value
                    Y_new = 0.5 * (Y_old + p_j)
                    solver.Y[j].value = Y_new
                    solver.Lambda[j].value = solver.Lambda[j].value + self.rho_admm * (p_j - Y_new)
                    pr_vals.append(primal_residual(p_j, Y_new))
                    du_vals.append(dual_residual(Y_new, Y_old))
            pr_avg = float(np.mean(pr_vals))
            du_avg = float(np.mean(du_vals))
            primal_hist.append(pr_avg)
            dual_hist.append(du_avg)
            if verbose:
                print_iteration(it, nu_norm=0.0, slack_norm=0.0, primal_res=pr_avg, dual_res=du_avg, dx=0.0, ds=0.0,
                                sigma=sigma_ref, tr_radius=self.rho_admm)
        runtime = time.time() - t0
        if verbose:
            print_summary(len(primal_hist), sigma_ref, runtime)
        return X_curr, U_curr, sigma_ref, primal_hist, dual_hist

    def solve_batched(self, X_refs: list, U_refs: list, sigma_ref: float, group=None, neighbor_radius=None):
        """Jacobi rounds on the device (and across the ranks of `group`).  Same return tuple as `solve`."""
        dev = torch.device("cuda")
        up = lambda lst: torch.as_tensor(np.ascontiguousarray(np.stack(lst), dtype=np.float64)).to(dev)   # noqa: E731
        eng = BatchedADMM(self.model.models, self.model.d_min, self.K, self.rho_admm, self.max_iter,
                          si_variant=self._SI_VARIANT, group=group, neighbor_radius=neighbor_radius)
        out = eng.solve(up(X_refs), up(U_refs), sigma_ref)
        X = out["X"].cpu().numpy(); U = out["U"].cpu().numpy()
        return [X[i] for i in range(self.N)], [U[i] for i in range(self.N)], sigma_ref, out["primal_hist"], out["dual_hist"]
