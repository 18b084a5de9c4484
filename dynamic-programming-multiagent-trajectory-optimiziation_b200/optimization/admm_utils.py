"""ADMM helpers -- same functions and values as SCvx/optimization/admm_utils.py:5-58.

These are host-side scalars-in/scalars-out utilities of the reference's public API; the batched
consensus round computes the same norms on the device (scvx_consensus_update)."""
import numpy as np

WEIGHT_COLLISION_SLACK = 1e5


def primal_residual(p_j: np.ndarray, Y_ij: np.ndarray) -> float:
    return np.linalg.norm(p_j - Y_ij)


def dual_residual(Y_new: np.ndarray, Y_old: np.ndarray) -> float:
    return np.linalg.norm(Y_new - Y_old)


def update_rho_admm(rho: float, primal_res: float, dual_res: float,
                    mu: float = 10.0, tau_inc: float = 2.0, tau_dec: float = 2.0) -> float:
    if primal_res > mu * dual_res:
        return rho * tau_inc
    elif dual_res > mu * primal_res:
        return rho / tau_dec
    else:
        return rho
