"""Inter-sample obstacle clearance -- same calls as SCvx/utils/intersample_collision.py:7-125, computed by
intersample_kernel (csrc/intersample.cu).

`make_segment_f(foh, U_ref_k, U_ref_kp1, sigma) -> (f_seg, dt_phys)` returns a SegmentFlow: callable like the reference's
closure (`f_seg(xk, _u, t)` = state after the fraction t of the segment, one launch per call) and, because it carries the
segment's data, usable by the batched kernel: `find_critical_times` / `linearize_h` / `h_i` accept it as their `f` and run
on the device.  An arbitrary Python callable cannot run inside a kernel: passing one raises NotImplementedError.
`T` must select leading state components (the first len(p_c) rows of the identity), which is what the reference uses.
`critical_times_batch` does every (agent, segment, obstacle) of a trajectory batch in ONE launch.
"""
from typing import List, Tuple

import numpy as np
import torch

from .. import _device


class SegmentFlow:
    def __init__(self, foh, U_ref_k, U_ref_kp1, sigma):
        self.foh = foh
        self.model_id = foh.model_id
        self.K = foh.K
        self.u0 = np.asarray(U_ref_k, dtype=float).reshape(-1)
        self.u1 = np.asarray(U_ref_kp1, dtype=float).reshape(-1)
        self.sigma = float(sigma)
        self.dt_phys = foh.dt * sigma

    def _tensors(self, xk):
        """A two-node 'trajectory' whose single segment is this one (K is kept: dt = 1/(K-1) enters the flow)."""
        dev = torch.device("cuda")
        n_x, n_u, K = self.foh.n_x, self.foh.n_u, self.K
        X = torch.zeros((1, n_x, K), dtype=torch.float64, device=dev)
        U = torch.zeros((1, n_u, K), dtype=torch.float64, device=dev)
        X[0, :, 0] = torch.as_tensor(np.asarray(xk, dtype=float).reshape(-1), device=dev)
        U[0, :, 0] = torch.as_tensor(self.u0, device=dev); U[0, :, 1] = torch.as_tensor(self.u1, device=dev)
        sig = torch.full((1,), self.sigma, dtype=torch.float64, device=dev)
        return X, U, sig

    def __call__(self, xk, _u_dummy, t):
        """x(t * dt_phys) from xk (intersample_collision.py:113-123)."""
        X, U, sig = self._tensors(xk)
        return _flow_point(self, X, U, sig, float(t))


def _flow_point(seg, X, U, sig, t):
    """The flow over the fraction t of a first-order-hold segment equals the flow over a WHOLE segment with end control
    u0 + t (u1 - u0) and time scale sigma * t: one call of the piecewise integrator (scvx_integrate_piecewise_batched)."""
    Ut = U.clone()
    Ut[0, :, 1] = U[0, :, 0] + t * (U[0, :, 1] - U[0, :, 0])
    out = _device.integrate_piecewise(seg.model_id, X, Ut, sig * t)
    return out[0, :, 1].cpu().numpy()


def make_segment_f(foh, U_ref_k, U_ref_kp1, sigma):
    seg = SegmentFlow(foh, U_ref_k, U_ref_kp1, sigma)
    return seg, seg.dt_phys


def _check(f, T, obstacle):
    if not isinstance(f, SegmentFlow):
        raise NotImplementedError("the device path needs the SegmentFlow returned by make_segment_f (no Python callables in kernels)")
    p_c, r = obstacle
    p_c = np.asarray(p_c, dtype=float).reshape(-1)
    T = np.asarray(T, dtype=float)
    m = p_c.size
    if T.shape != (m, f.foh.n_x) or not np.array_equal(T, np.eye(f.foh.n_x)[:m]):
        raise NotImplementedError("T must be the first len(p_c) rows of the identity")
    return p_c, float(r), m


def _run(f, xk, p_c, r, m, dt, num_samples, eps, tol, max_roots=8):
    X, U, sig = f._tensors(xk)
    dev = X.device
    # only segment 0 of the two-node trajectory is meaningful; the kernel walks K-1 segments, read back the first
    c = torch.as_tensor(p_c.reshape(1, 1, m), device=dev); rr = torch.full((1, 1), r, dtype=torch.float64, device=dev)
    n, t, h0, g = _device.intersample(f.model_id, X, U, sig, c, rr, proj_dim=m, t_range=dt, num_samples=num_samples, eps=eps,
                                      tol=tol, max_roots=max_roots)
    return int(n[0, 0, 0].item()), t[0, 0, 0].cpu().numpy(), h0[0, 0, 0].cpu().numpy(), g[0, 0, 0].cpu().numpy()


def h_i(xk, uk, t, f, T, obstacle) -> float:  # noqa: ARG001
    """|| T f(xk, uk, t) - p_c || - r (intersample_collision.py:7-26)."""
    p_c, r, m = _check(f, T, obstacle)
    xt = f(xk, uk, t)
    return float(np.linalg.norm(xt[:m] - p_c) - r)


def find_critical_times(xk, uk, f, T, obstacle, dt, num_samples: int = 100, eps: float = 1e-4, tol: float = 1e-6) -> List[float]:  # noqa: ARG001
    """All interior minima t* of the clearance on (0, dt) (intersample_collision.py:29-72)."""
    p_c, r, m = _check(f, T, obstacle)
    n, t, _, _ = _run(f, xk, p_c, r, m, dt, num_samples, eps, tol)
    return [float(v) for v in t[:min(n, t.size)]]


def linearize_h(xk, uk, t_star, f, T, obstacle, eps: float = 1e-4) -> Tuple[float, np.ndarray, np.ndarray]:
    """h0, grad_x (central differences in xk), grad_u (identically zero for a segment flow, as in the reference:
    intersample_collision.py:75-97 perturbs an argument the flow ignores)."""
    p_c, r, m = _check(f, T, obstacle)
    xk = np.asarray(xk, dtype=float).reshape(-1)
    h0 = h_i(xk, uk, t_star, f, T, obstacle)
    grad_x = np.zeros_like(xk)
    for j in range(xk.size):
        xp, xm = xk.copy(), xk.copy()
        xp[j] += eps; xm[j] -= eps
        grad_x[j] = (h_i(xp, uk, t_star, f, T, obstacle) - h_i(xm, uk, t_star, f, T, obstacle)) / (2 * eps)
    return h0, grad_x, np.zeros_like(np.asarray(uk, dtype=float))


def critical_times_batch(model_id, X, U, sigma, obs_c, obs_r, **kw):
    """Every (agent, segment, obstacle) of a batch in one launch; see _device.intersample."""
    return _device.intersample(model_id, X, U, sigma, obs_c, obs_r, **kw)
