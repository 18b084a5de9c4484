"""Trajectory analysis metrics -- same calls as SCvx/utils/analysis.py:10-62, computed on the device.

`X_list` may be the reference's list of (n_x, K) numpy arrays or a stacked (N, n_x, K) CUDA tensor."""
import numpy as np
import torch

from .. import _device


def _stack(X_list):
    if isinstance(X_list, torch.Tensor):
        return X_list
    return torch.as_tensor(np.stack([np.asarray(x, dtype=float) for x in X_list]), device="cuda")


def min_inter_agent_distance(X_list):
    """(d_min_global, d_mat (N, N)): pairwise minimum distance over the horizon on rows 0:3 (analysis.py:10-31).
    Raises ValueError when no pair has a positive distance (numpy's min of an empty selection)."""
    X = _stack(X_list)
    d_min, d_mat = _device.min_inter_agent_distance(X, n_rows=min(3, X.shape[1]))
    v = float(d_min.item())
    if not np.isfinite(v):
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    return v, d_mat.cpu().numpy()


def min_agent_obstacle_distance(X_list, obstacles, robot_radius):
    """(d_min_global, d_mat (N, M)): min_k ||p_i(k) - c_j|| - (robot_radius + r_j) (analysis.py:34-62)."""
    X = _stack(X_list)
    n_rows = min(3, X.shape[1])
    M = len(obstacles)
    if X.shape[0] * M == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    C = torch.as_tensor(np.array([np.asarray(c, dtype=float).ravel()[:n_rows] for c, _ in obstacles]), device="cuda")
    R = torch.as_tensor(np.array([float(r) for _, r in obstacles]), device="cuda")
    d_min, d_mat = _device.min_agent_obstacle_distance(X, C, R, robot_radius, n_rows=n_rows)
    return float(d_min.item()), d_mat.cpu().numpy()
