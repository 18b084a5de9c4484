"""Shared device engine of the Distributed_opt mirrors: batches every robot's QP of one x_traj_opt call into one launch
of `scvx_lti_qp_batched` (and `scvx_sbar_qp_batched` for the consensus step of ADMM_decentralized)."""
import ctypes

import numpy as np
import torch

from .. import _lib
from .._lib import ScvxError, LtiArgs, check, load, ptr, stream_ptr

F64 = torch.float64


def zoh_double_integrator(dt, n, m):
    """Exact zero-order-hold discretisation of the double integrator xdot = [0 I; 0 0] x + [0; I] u (what
    scipy.signal.StateSpace(...).to_discrete(dt) returns in the scripts' descete_f): closed form, host side."""
    d = n // 2
    Ad = np.eye(n); Ad[:d, d:] = dt * np.eye(d)
    Bd = np.zeros((n, m)); Bd[:d, :] = 0.5 * dt * dt * np.eye(d); Bd[d:, :] = dt * np.eye(d)
    return [Ad, Bd]


def _dev(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


def collision_tables(P, R):
    """P (N, T, dc) positions of all robots -> h (N, N-1, T), g (N, N-1, dc, T) on the device:
    h = 2R - |p_i - p_q|, g = (p_i - p_q)/|p_i - p_q| for every other robot q (ADMM_decentralized.py:126-137,
    dist_scvx_3d.py:94-107; no epsilon in the denominator).  Elementwise torch on the device (set-up of the tables)."""
    N, T, dc = P.shape
    idx = torch.tensor([[q for q in range(N) if q != i] for i in range(N)], device=P.device)       # (N, N-1)
    diff = P[:, None, :, :] - P[idx]                                                                  # (N, N-1, T, dc)
    nrm = torch.linalg.norm(diff, dim=3)                                                              # (N, N-1, T)
    h = 2.0 * R - nrm
    g = (diff / nrm[..., None]).permute(0, 1, 3, 2).contiguous()                                      # (N, N-1, dc, T)
    return h.contiguous(), g


def solve_robot_qps(Ad, Bd, X, x_des, tr, c_w, box, rho=0.0, lin=None, sbar=None, col_h=None, col_g=None, c_S=0.0,
                    max_iter=0):
    """X (R, T, n+m) device tensor of current trajectories [state | control]; returns s (R, T, n+m) = [d | w], objective,
    status, iters, S (R, T)."""
    lib = load()
    dev = X.device
    R, T, nm = X.shape
    n, m = Ad.shape[0], Bd.shape[1]
    assert nm == n + m
    x = X[:, :, :n].contiguous(); u = X[:, :, n:].contiguous()
    nq = 0 if col_h is None else col_h.shape[1]
    d = torch.empty((R, T, n), dtype=F64, device=dev); w = torch.empty((R, T, m), dtype=F64, device=dev)
    S = torch.zeros((R, T), dtype=F64, device=dev)
    obj = torch.empty(R, dtype=F64, device=dev)
    status = torch.empty(R, dtype=torch.int32, device=dev); iters = torch.empty(R, dtype=torch.int32, device=dev)
    nbytes = int(lib.scvx_lti_qp_workspace_bytes(R, T, n, nq))
    wsb = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev)
    keep = [_dev(Ad, dev), _dev(Bd, dev), _dev(x_des, dev)]
    a = LtiArgs()
    a.n_robots, a.T, a.n, a.m, a.nq, a.max_iter = R, T, n, m, nq, int(max_iter)
    a.Ad, a.Bd, a.x, a.u, a.x_des = ptr(keep[0]), ptr(keep[1]), ptr(x), ptr(u), ptr(keep[2])
    a.tr, a.c_w, a.rho, a.c_S = float(tr), float(c_w), float(rho), float(c_S)
    (a.box_lo0, a.box_hi0), (a.box_lo1, a.box_hi1) = box
    if lin is not None:
        lin = lin.contiguous(); a.lin = ptr(lin)
    if sbar is not None:
        sbar = sbar.contiguous(); a.sbar = ptr(sbar)
    if nq:
        ch = col_h[:, :, :T - 1].contiguous(); cg = col_g[:, :, :, :T - 1].contiguous()
        keep += [ch, cg]
        a.col_h, a.col_g = ptr(ch), ptr(cg)
    a.d, a.w, a.S, a.objective, a.status, a.iters = ptr(d), ptr(w), ptr(S), ptr(obj), ptr(status), ptr(iters)
    a.workspace, a.workspace_bytes = ptr(wsb), nbytes
    check(lib.scvx_lti_qp_batched(ctypes.byref(a), stream_ptr()), "scvx_lti_qp_batched")
    return torch.cat([d, w], dim=2), obj, status, iters, S


def solve_sbar_qps(s_pos, r_dual, rho, col_h, col_g, c_S=1e6):
    """s_pos, r_dual (R, T, 2); col_h (R, nq, T); col_g (R, nq, 2, T) -> sbar (R, T, 2), S (R, T)."""
    lib = load()
    dev = s_pos.device
    R, T, _ = s_pos.shape
    nq = col_h.shape[1]
    sbar = torch.empty((R, T, 2), dtype=F64, device=dev); S = torch.empty((R, T), dtype=F64, device=dev)
    wsb = torch.empty(R * T * 2 * nq, dtype=F64, device=dev)
    s_pos, r_dual, col_h, col_g = s_pos.contiguous(), r_dual.contiguous(), col_h.contiguous(), col_g.contiguous()
    check(lib.scvx_sbar_qp_batched(R, T, nq, float(rho), float(c_S), ptr(s_pos), ptr(r_dual), ptr(col_h), ptr(col_g), ptr(sbar),
                                   ptr(S), ptr(wsb), wsb.numel() * 8, stream_ptr()), "scvx_sbar_qp_batched")
    # The kernel has no per-(robot, t) status word; a non-finite input (coincident robots give NaN normals in the collision tables)
    # or a step it had to abandon shows up as a non-finite result, which must not flow into the dual update unnoticed.
    if not bool(torch.isfinite(col_g).all().item()):
        raise ScvxError("scvx_sbar_qp_batched: non-finite collision normals (coincident robots?)")
    if not bool((torch.isfinite(sbar).all() & torch.isfinite(S).all()).item()):
        raise ScvxError("scvx_sbar_qp_batched: non-finite consensus solution")
    return sbar, S
