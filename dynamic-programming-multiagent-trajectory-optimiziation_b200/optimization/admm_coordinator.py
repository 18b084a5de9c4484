"""ADMMCoordinator -- mirrors SCvx/optimization/admm_coordinator.py:13-118.

`solve` keeps the reference's sequential Gauss-Seidel order (agent i sees the trajectories agents 0..i-1 produced in the
same round) with the per-agent discretisation and sub-problem on the GPU.  `solve_batched` is the device-resident Jacobi variant
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
        self.model, self.N, self.K = multi_agent_model, multi_agent_model.N, K
        self.rho_admm, self.max_iter = rho_admm, max_iter
        self.agent_solvers = [self._SOLVER(i, multi_agent_model, rho_admm, K) for i in range(self.N)]
        self.discretizers = [FirstOrderHold(m, K) for m in multi_agent_model.models]

    # -- the literal coordinator: agents one after the other (Gauss-Seidel), host arrays in and out ------------------------
    def _reset_consensus(self, X_refs):
        """Every agent's copy Y_j of neighbour j starts at j's reference positions, the multipliers at zero."""
        d = self._D
        for solver in self.agent_solvers:
            for j, holder in solver.Y.items():
                holder.value = np.array(X_refs[j][0:d, :])
                solver.Lambda[j].value = np.zeros((d, self.K))

    def _agent_round(self, i, X_now, U_now, X_refs, U_refs, sigma_ref):
        """Agent i's sub-problem of this round about its current trajectory (the single-integrator coordinator keeps
        linearising about the INITIAL references, si_admm_coordinator.py:80-86); returns (X_i, U_i, p_i)."""
        mats = self.discretizers[i].calculate_discretization(X_now[i], U_now[i], sigma_ref)
        about = (X_refs[i], U_refs[i]) if self._SI_VARIANT else (X_now[i], U_now[i])
        solver = self.agent_solvers[i]
        solver.setup(about[0], about[1], sigma_ref, mats, {j: X_now[j] for j in range(self.N) if j != i})
        X_i, U_i, _nu, _slacks, p_i = solver.solve(solver="ECOS")
        return X_i, U_i, p_i

    def _consensus_round(self, positions):
        """Y_j <- (Y_j + p_j) / 2, Lambda_j += rho (p_j - Y_j) in every agent's copy; mean residuals over all (i, j) copies."""
        primal, dual = [], []
        for solver in self.agent_solvers:
            for j, holder in solver.Y.items():
                before = holder.value
                after = 0.5 * (before + positions[j])
                holder.value = after
                solver.Lambda[j].value = solver.Lambda[j].value + self.rho_admm * (positions[j] - after)
                primal.append(primal_residual(positions[j], after))
                dual.append(dual_residual(after, before))
        return float(np.mean(primal)), float(np.mean(dual))

    def solve(self, X_refs: list, U_refs: list, sigma_ref: float, verbose: bool = True):
        self._reset_consensus(X_refs)
        X_now, U_now = list(X_refs), list(U_refs)
        history = {"primal": [], "dual": []}
        started = time.time()
        for rnd in range(self.max_iter):
            positions = {}
            for i in range(self.N):
                X_now[i], U_now[i], positions[i] = self._agent_round(i, X_now, U_now, X_refs, U_refs, sigma_ref)
            pr, du = self._consensus_round(positions)
            history["primal"].append(pr); history["dual"].append(du)
            if verbose:
                print_iteration(rnd, nu_norm=0.0, slack_norm=0.0, primal_res=pr, dual_res=du, dx=0.0, ds=0.0, sigma=sigma_ref,
                                tr_radius=self.rho_admm)
        if verbose:
            print_summary(len(history["primal"]), sigma_ref, time.time() - started)
        return X_now, U_now, sigma_ref, history["primal"], history["dual"]

    def solve_batched(self, X_refs: list, U_refs: list, sigma_ref: float, group=None, neighbor_radius=None):
        """Jacobi rounds on the device (and across the ranks of `group`).  Same return tuple as `solve`."""
        dev = torch.device("cuda")
        up = lambda lst: torch.as_tensor(np.ascontiguousarray(np.stack(lst), dtype=np.float64)).to(dev)   # noqa: E731
        eng = BatchedADMM(self.model.models, self.model.d_min, self.K, self.rho_admm, self.max_iter,
                          si_variant=self._SI_VARIANT, group=group, neighbor_radius=neighbor_radius)
        out = eng.solve(up(X_refs), up(U_refs), sigma_ref)
        X = out["X"].cpu().numpy(); U = out["U"].cpu().numpy()
        return [X[i] for i in range(self.N)], [U[i] for i in range(self.N)], sigma_ref, out["primal_hist"], out["dual_hist"]
