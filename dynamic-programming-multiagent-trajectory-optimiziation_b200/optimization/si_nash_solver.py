"""SI_NashSolver -- mirrors SCvx/optimization/si_nash_solver.py:16-125 (iterative best response, 3-D single integrators)."""
import time
from typing import List, Tuple

import numpy as np

from ..discretization.first_order_hold import FirstOrderHold
from ..global_parameters import TRUST_RADIUS0, K
from ..utils.multi_agent_logging import print_iteration, print_summary
from .si_agent_best_response import SI_AgentBestResponse


class SI_NashSolver:
    def __init__(self, multi_agent_model, max_iter: int = 20, tol: float = 1e-3, max_acs_iters: int = 5,
                 acs_tol: float = 1e-3, K=K) -> None:
        self.mam = multi_agent_model
        self.N = multi_agent_model.N
        self.max_iter, self.tol, self.max_acs_iters, self.acs_tol, self.K = max_iter, tol, max_acs_iters, acs_tol, K
        self.br_solvers = [SI_AgentBestResponse(i, multi_agent_model, K) for i in range(self.N)]
        self.fohs = [FirstOrderHold(m, K) for m in multi_agent_model.models]

    def solve(self, X_refs: List[np.ndarray], U_refs: List[np.ndarray], sigma_ref: float = 1.0, verbose: bool = False,
              show_progress: bool = True) -> Tuple[List[np.ndarray], List[np.ndarray], List[float]]:  # noqa: ARG002
        X_curr = [np.array(x, dtype=float) for x in X_refs]
        U_curr = [np.array(u, dtype=float) for u in U_refs]
        change_hist: List[float] = []
        t0 = time.time()
        for it in range(self.max_iter):
            max_change = 0.0
            X_prev_all = [x.copy() for x in X_curr]
            for i in range(self.N):
                br = self.br_solvers[i]
                mats = self.fohs[i].calculate_discretization(X_curr[i], U_curr[i], sigma_ref)
                neigh_cur = {j: X_curr[j] for j in range(self.N) if j != i}
                neigh_prev = {j: X_prev_all[j] for j in range(self.N) if j != i}
                br.setup(X_ref=X_curr[i], U_ref=U_curr[i], sigma_ref=sigma_ref, discr_mats=mats, neighbour_refs=neigh_cur,
                         X_prev=X_prev_all[i], neighbour_prev_refs=neigh_prev, tr_radius=TRUST_RADIUS0)
                X_new = U_new = None
                for _ in range(self.max_acs_iters):
                    X_new, U_new, *_ = br.solve()
                    br.model.update_slabs(X_new[0:3, :], [X_curr[j][0:3, :] for j in range(self.N) if j != i])
                    if np.linalg.norm(X_new - X_curr[i]) < self.acs_tol:
                        break
                delta = float(np.linalg.norm(X_new - X_curr[i]))
                max_change = max(max_change, delta)
                X_curr[i], U_curr[i] = X_new, U_new
                if verbose:
                    print(f"  Agent {i}: dX={delta:.2e}")
            change_hist.append(max_change)
            if verbose:
                print_iteration(it, 0.0, 0.0, max_change, 0.0, max_change, 0.0, sigma_ref, 0.0)
            if max_change < self.tol:
                break
        if verbose:
            print_summary(len(change_hist), sigma_ref, time.time() - t0)
        return X_curr, U_curr, change_hist
