"""SCVXSolver -- the single-agent driver with the reference's interface (SCvx/optimization/scvx_solver.py:10-133):
`SCVXSolver(model).solve(verbose=False, initial_sigma=1.0) -> (X, U, sigma, logger)`, public attributes `max_iter`,
`conv_tol`, `tr_radius`, the three weights, `discretizer`, `problem`, `logger`.

It is the batch-of-one view of `scvx_b200.batch.BatchedSCvx`: the whole outer loop -- discretisation, obstacle
linearisation, sub-problem, and the bookkeeping of scvx_solver.py:82-111,125-133 (metrics, convergence test, break before
accepting the converged iterate, grow-only trust region; `scvx_outer_update` in csrc/linearize.cu) -- runs on the device
with no host round trip per iteration.  The host looks at the device state every `check_every` iterations, then turns the
device metrics table into the reference's log records.
"""
import numpy as np
import torch

from .. import _lib
from ..global_parameters import CONV_TOL, MAX_ITER, TRUST_RADIUS0, WEIGHT_NU, WEIGHT_SIGMA, WEIGHT_SLACK, K
from ..utils.logging import Logger


class SCVXSolver:
    check_every = 5          # outer iterations between two host looks at the device state

    def __init__(self, model, K=K):
        self.model, self.K = model, K
        self.max_iter, self.conv_tol, self.tr_radius = MAX_ITER, CONV_TOL, TRUST_RADIUS0
        self.weight_nu, self.weight_slack, self.weight_sigma = WEIGHT_NU, WEIGHT_SLACK, WEIGHT_SIGMA
        self.logger = Logger()
        self._discretizer = self._problem = None

    # The reference builds these two in its constructor and user code reaches into them (`solver.problem.prob.value`,
    # `solver.discretizer.dt`); here they are not part of the loop, so they come into being on first access.
    @property
    def discretizer(self):
        if self._discretizer is None:
            from ..discretization.first_order_hold import FirstOrderHold
            self._discretizer = FirstOrderHold(self.model, self.K)
        return self._discretizer

    @property
    def problem(self):
        if self._problem is None:
            from .sc_problem import SCProblem
            self._problem = SCProblem(self.model, self.K)
        return self._problem

    def _engine(self):
        from ..batch import BatchedSCvx
        return BatchedSCvx([self.model], self.K, max_iter=self.max_iter, tr_radius0=self.tr_radius, conv_tol=self.conv_tol,
                           weight_nu=self.weight_nu, weight_slack=self.weight_slack, weight_sigma=self.weight_sigma)

    def solve(self, verbose=False, initial_sigma=1.0):
        model, K = self.model, self.K
        X0, U0 = model.initialize_trajectory(np.zeros((model.n_x, K)), np.zeros((model.n_u, K)))
        eng = self._engine()
        dev = eng.batch.device
        to_dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)[None]).to(dev)   # noqa: E731
        out = eng.solve(to_dev(X0), to_dev(U0), initial_sigma=float(initial_sigma), early_exit=True,
                        check_every=max(1, int(self.check_every)))

        ran = out["was_active"][:, 0].cpu().numpy().astype(bool)          # iterations the agent actually took
        status = out["status"][:, 0].cpu().numpy()
        metrics = out["metrics"][:, 0].cpu().numpy()
        self.logger.clear()
        for it in np.flatnonzero(ran):
            if status[it] == _lib.ST_NUMERICAL:
                raise RuntimeError(f"SCvx iteration {it}: convex subproblem infeasible")
            self.logger.log_metrics([metrics[it]], first_iter=int(it))
            if verbose:
                nu_norm, slack_norm, dx, _, ds, _ = metrics[it]
                print(f"Iter {it}: nu={nu_norm:.3e}, slack={slack_norm:.3e}, dx={dx:.3e}, ds={ds:.3e}")

        # what a caller of the reference finds afterwards: the grown trust radius on the solver and the obstacle slacks of
        # the last sub-problem on the model (scvx_solver.py:117-133)
        self.tr_radius = float(out["tr_radius"][0].item())
        slack = eng.ws.s_prime[0].cpu().numpy()
        for row, holder in zip(slack, model.s_prime):
            holder.value = row.reshape(K, 1).copy()
        X, U = out["X"][0].cpu().numpy(), out["U"][0].cpu().numpy()
        return model.x_redim(X), model.u_redim(U), float(out["sigma"][0].item()), self.logger
