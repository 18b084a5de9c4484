"""SCProblem -- mirrors SCvx/optimization/sc_problem.py:6-129 on the GPU.

Same constructor, `.var` / `.par` keys, `set_parameters(**kw)` (KeyError on unknown names),
`solve(**kw) -> bool` error flag, `get_variable(name)`, and `.prob.value` (objective).  Where the
reference builds a cvxpy graph and calls ECOS, `solve` uploads the parameters, runs the obstacle
linearisation kernel about X_ref and the batched interior-point kernel with a batch of one
(scvx_solve_batched), and downloads X, U, nu, sigma and the obstacle slacks.
"""
import warnings

import numpy as np
import torch

from .. import _device, _lib
from ..batch import AgentBatch
from ..global_parameters import K


class _Holder:
    """Minimal stand-in for a cvxpy Variable / Parameter: a `.value` slot (+ shape)."""

    def __init__(self, shape=()):
        self.shape = shape
        self.value = None


class _ProblemView:
    """`.value` = objective of the last solve, `.status` in cvxpy's vocabulary."""

    def __init__(self):
        self.value = None
        self.status = None


_PAR_SHAPES = lambda n_x, n_u, K: {   # noqa: E731  (sc_problem.py:29-44)
    "A_bar": (n_x * n_x, K - 1), "B_bar": (n_x * n_u, K - 1), "C_bar": (n_x * n_u, K - 1),
    "S_bar": (n_x, K - 1), "z_bar": (n_x, K - 1), "X_ref": (n_x, K), "U_ref": (n_u, K), "sigma_ref": (),
    "weight_nu": (), "weight_sigma": (), "tr_radius": (), "weight_slack": (),
}


class SCProblem:
    def __init__(self, model, K=K):
        self.model = model
        self.n_x = model.n_x
        self.n_u = model.n_u
        self.K = K
        n_x, n_u = self.n_x, self.n_u
        self.var = {"X": _Holder((n_x, K)), "U": _Holder((n_u, K)), "nu": _Holder((n_x, K - 1)), "sigma": _Holder(())}
        self.par = {k: _Holder(s) for k, s in _PAR_SHAPES(n_x, n_u, K).items()}
        self.prob = _ProblemView()
        self.status = None
        self.iters = None
        self._batch = AgentBatch([model], K)
        b = self._batch
        self._ws = {}                                          # keyed by number of neighbour slots
        self._obs_a = torch.empty((1, b.M, b.d, K), dtype=torch.float64, device=b.device)
        self._obs_b = torch.empty((1, b.M, K), dtype=torch.float64, device=b.device)

    # -- reference API -----------------------------------------------------------------------------
    def set_parameters(self, **kwargs):
        for key, val in kwargs.items():
            if key in self.par:
                self.par[key].value = val
            else:
                raise KeyError(f"Parameter '{key}' not found in SCProblem.")

    def solve(self, **kwargs):
        """Returns True on solver error (the reference's convention, sc_problem.py:95-105).
        Solver keyword arguments of the reference (solver="ECOS", warm_start=..., verbose=...) are accepted
        and ignored; `max_iter=` caps the interior-point iterations."""
        self._solve_device(max_iter=int(kwargs.get("max_iter", 0)))
        if self.status == _lib.ST_MAXITER:
            # the reference's solver would report "optimal_inaccurate" here and cvxpy lets that pass as success too; say so
            warnings.warn(f"SCProblem.solve: the interior-point method stopped at its iteration cap ({self.iters}); the returned "
                          "iterate is primal feasible but not certified optimal (prob.status == 'optimal_inaccurate')", RuntimeWarning,
                          stacklevel=2)
        return self.status != _lib.ST_OPTIMAL and self.status != _lib.ST_MAXITER

    def get_variable(self, name):
        if name in self.var:
            return self.var[name].value
        raise KeyError(f"Variable '{name}' not found.")

    def print_available_parameters(self):
        print("Available parameters:")
        for k in self.par:
            print(f"  {k}")

    def print_available_variables(self):
        print("Available variables:")
        for k in self.var:
            print(f"  {k}")

    # -- device path -------------------------------------------------------------------------------
    def _dev(self, name):
        v = self.par[name].value
        if v is None:
            raise ValueError(f"parameter '{name}' has no value")
        shape = self.par[name].shape
        a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(shape if shape else (1,)))
        return torch.as_tensor(a).to(self._batch.device).unsqueeze(0) if shape else torch.as_tensor(a).to(self._batch.device)

    def _solve_device(self, max_iter=0, col_a=None, col_b=None, quad_rho=None, lin_p=None, weight_col=1e5,
                      quad_diag=None, lin_w=None, quad_pair=None, fix_sigma=False):
        b = self._batch
        n_nbr = 0 if col_a is None else col_a.shape[1]
        ws = self._ws.get(n_nbr)
        if ws is None:
            ws = self._ws[n_nbr] = _device.SubproblemWorkspace(b.model_id, 1, self.K, b.M, n_nbr, b.device)
        mats = tuple(self._dev(k) for k in ("A_bar", "B_bar", "C_bar", "S_bar", "z_bar"))
        X_ref, U_ref = self._dev("X_ref"), self._dev("U_ref")
        if b.M:
            _device.linearize_obstacles(b.model_id, X_ref, b.obs_c, b.obs_clear, out=(self._obs_a, self._obs_b))
        _device.solve_subproblem(
            ws, mats, X_ref, U_ref, self._dev("sigma_ref"), self._dev("tr_radius"), b.x_init, b.x_final,
            b.pos_lo, b.pos_hi, b.v_max, b.w_max, self._obs_a, self._obs_b,
            float(self.par["weight_nu"].value), float(self.par["weight_slack"].value), float(self.par["weight_sigma"].value),
            col_a=col_a, col_b=col_b, quad_rho=quad_rho, lin_p=lin_p, weight_col=weight_col, max_iter=max_iter,
            quad_diag=quad_diag, lin_w=lin_w, quad_pair=quad_pair, fix_sigma=fix_sigma)
        self.var["X"].value = ws.X[0].cpu().numpy()
        self.var["U"].value = ws.U[0].cpu().numpy()
        self.var["nu"].value = ws.nu[0].cpu().numpy()
        self.var["sigma"].value = float(ws.sigma[0].item())
        sp = ws.s_prime[0].cpu().numpy()
        for j, holder in enumerate(self.model.s_prime):
            holder.value = sp[j].reshape(self.K, 1).copy()
        self.status = int(ws.status[0].item())
        self.iters = int(ws.iters[0].item())
        self.prob.value = float(ws.objective[0].item())
        self.prob.status = {0: "optimal", 1: "optimal_inaccurate", 2: "solver_error"}[self.status]
        self._last_ws = ws
        return ws
