"""NashSolver -- the iterative-best-response driver with the reference's interface (SCvx/optimization/nash_solver.py:16-149):
`NashSolver(mam, max_iter=20, tol=1e-3, max_acs_iters=5, acs_tol=1e-3).solve(X_refs, U_refs, sigma_ref=1.0, verbose=False)
-> (X_list, U_list, change_hist)`.

Semantics kept from the reference: agents respond one after the other within a sweep (Gauss-Seidel: agent i sees the
trajectories agents 0..i-1 have just produced), the slab normals of a response are initialised from the snapshot taken at the
start of the sweep, a response is re-solved with refreshed normals up to `max_acs_iters` times (the ACS loop) or until it
stops moving, and the sweeps stop when no agent moved by more than `tol`.  The discretisation and every sub-problem run on
the GPU; `scvx_b200.batch.BatchedNash` is the many-agents form (all responses of a sweep, or of a colour phase, in one launch).
"""
import time
from typing import List, Tuple

import numpy as np

from ..discretization.first_order_hold import FirstOrderHold
from ..global_parameters import TRUST_RADIUS0, K
from ..utils.multi_agent_logging import print_iteration, print_summary
from .agent_best_response import AgentBestResponse


class _BestResponseSweeps:
    """Shared machinery of NashSolver / SI_NashSolver; subclasses name the response class and the position dimension."""
    _RESPONSE = None
    _D = 2

    def __init__(self, multi_agent_model, max_iter: int = 20, tol: float = 1e-3, max_acs_iters: int = 5,
                 acs_tol: float = 1e-3, K=K) -> None:
        self.mam, self.N, self.K = multi_agent_model, multi_agent_model.N, K
        self.max_iter, self.tol, self.max_acs_iters, self.acs_tol = max_iter, tol, max_acs_iters, acs_tol
        self.br_solvers = [self._RESPONSE(i, multi_agent_model, K) for i in range(self.N)]
        self.fohs = [FirstOrderHold(m, K) for m in multi_agent_model.models]

    # -- one agent -------------------------------------------------------------------------------------------------
    def _others(self, trajectories, i):
        return {j: trajectories[j] for j in range(self.N) if j != i}

    def _respond(self, i, X_now, U_now, snapshot, sigma_ref):
        """Best response of agent i against X_now (others) with the ACS loop; returns (X_i, U_i, how far it moved)."""
        d, br = self._D, self.br_solvers[i]
        mats = self.fohs[i].calculate_discretization(X_now[i], U_now[i], sigma_ref)
        br.setup(X_ref=X_now[i], U_ref=U_now[i], sigma_ref=sigma_ref, discr_mats=mats, neighbour_refs=self._others(X_now, i),
                 X_prev=snapshot[i], neighbour_prev_refs=self._others(snapshot, i), tr_radius=TRUST_RADIUS0)
        rivals = [X_now[j][0:d, :] for j in range(self.N) if j != i]
        X_i = U_i = None
        for _attempt in range(self.max_acs_iters):
            X_i, U_i = br.solve()[:2]
            br.model.update_slabs(X_i[0:d, :], rivals)          # normals along (new own position) - (rivals' positions)
            if np.linalg.norm(X_i - X_now[i]) < self.acs_tol:
                break
        return X_i, U_i, float(np.linalg.norm(X_i - X_now[i]))

    def _agent_line(self, i, moved):
        return f"  Agent {i}: cost={self.br_solvers[i].scp.prob.value:8.3f}, dX={moved:6.2e}"

    # -- the driver ------------------------------------------------------------------------------------------------
    def solve(self, X_refs: List[np.ndarray], U_refs: List[np.ndarray], sigma_ref: float = 1.0, verbose: bool = False,
              **_ignored) -> Tuple[List[np.ndarray], List[np.ndarray], List[float]]:
        X_now = [np.array(x, dtype=float) for x in X_refs]
        U_now = [np.array(u, dtype=float) for u in U_refs]
        started, history = time.time(), []
        while len(history) < self.max_iter:
            snapshot = [x.copy() for x in X_now]
            largest = 0.0
            for i in range(self.N):
                X_now[i], U_now[i], moved = self._respond(i, X_now, U_now, snapshot, sigma_ref)
                largest = max(largest, moved)
                if verbose:
                    print(self._agent_line(i, moved))
            history.append(largest)
            if verbose:
                print_iteration(len(history) - 1, 0.0, 0.0, largest, 0.0, largest, 0.0, sigma_ref, 0.0)
            if largest < self.tol:
                break
        if verbose:
            print_summary(len(history), sigma_ref, time.time() - started)
        return X_now, U_now, history


class NashSolver(_BestResponseSweeps):
    _RESPONSE = AgentBestResponse
    _D = 2
