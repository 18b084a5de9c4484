"""Warm start for the unicycle -- same call as SCvx/utils/initial_guess.py:61-107, computed by warm_start_kernel<2>.

`initial_guess(p0, p1, obstacles, clearance, K) -> (X0 (3, K), U0 (2, K))` for one agent, `initial_guess_batch` for many
(one launch).  Errors follow the reference: ValueError when start or goal lies inside / on an inflated obstacle that the
straight segment crosses, ValueError when numpy.linspace would be asked for a negative number of samples."""
import numpy as np
import torch

from .. import _device, _lib

_MODEL = _lib.MODEL_UNICYCLE
_D = 2
_ERRORS = {
    1: "Point inside/on circle; no tangents.",
    2: "Number of samples, {n}, must be non-negative.",
    3: "p0 and p1 are too close for detour computation.",
}


def _pack(p0s, p1s, obstacle_lists, d):
    n = len(p0s)
    M = max((len(o) for o in obstacle_lists), default=0)
    P0 = np.zeros((n, 3)); P1 = np.zeros((n, 3))
    C = np.zeros((n, max(M, 1), d)); R = np.zeros((n, max(M, 1))); cnt = np.zeros(n, dtype=np.int32)
    for i in range(n):
        a = np.asarray(p0s[i], dtype=float).ravel(); b = np.asarray(p1s[i], dtype=float).ravel()
        P0[i, :min(3, a.size)] = a[:3]; P1[i, :min(3, b.size)] = b[:3]
        cnt[i] = len(obstacle_lists[i])
        for j, (c, r) in enumerate(obstacle_lists[i]):
            C[i, j] = np.asarray(c, dtype=float).ravel()[:d]; R[i, j] = float(r)
    return P0, P1, C, R, cnt, M


def _run(model_id, d, p0s, p1s, obstacle_lists, clearance, K):
    P0, P1, C, R, cnt, M = _pack(p0s, p1s, obstacle_lists, d)
    dev = torch.device("cuda")
    X0, U0, st = _device.warm_start(
        model_id, torch.as_tensor(P0, device=dev), torch.as_tensor(P1, device=dev),
        torch.as_tensor(C, device=dev) if M else None, torch.as_tensor(R, device=dev) if M else None,
        clearance, K, obs_count=torch.as_tensor(cnt, device=dev) if M else None)
    st = st.cpu().numpy()
    bad = np.nonzero(st)[0]
    if bad.size:
        raise ValueError(_ERRORS[int(st[bad[0]])].format(n="<0") + (f" (agent {int(bad[0])})" if len(p0s) > 1 else ""))
    return X0, U0


def initial_guess_batch(p0s, p1s, obstacle_lists, clearance, K):
    """Device tensors X0 (n, 3, K), U0 (n, 2, K) for n agents with their own obstacle lists."""
    return _run(_MODEL, _D, p0s, p1s, obstacle_lists, clearance, K)


def initial_guess(p0, p1, obstacles, clearance, K):
    X0, U0 = _run(_MODEL, _D, [p0], [p1], [list(obstacles)], clearance, K)
    return X0[0].cpu().numpy(), U0[0].cpu().numpy()
