"""Host-side ADMM scalars of the reference's public API (SCvx/optimization/admm_utils.py:5-58): the collision-slack weight,
the two residual norms and the residual-balancing rule for rho.

Only the NAMES, argument order and returned values are the reference's.  The batched consensus round never calls these: it
computes the same two norms for every agent on the device (`scvx_consensus_update`, csrc/linearize.cu); these functions
exist for user code that post-processes trajectories on the host, and as the known-answer anchor of
SCvx/multi_agent_tests/test_admm_utils.py:7-45 (reproduced in tests/test_oracle_golden.py).
"""
import numpy as np

WEIGHT_COLLISION_SLACK = 1e5


def _gap_norm(a, b) -> float:
    # Frobenius norm of a - b, accumulated the way numpy.linalg.norm does for real input (one dot product of the
    # flattened difference), so that values agree with the reference to the last bit
    gap = np.subtract(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)).ravel(order="K")
    return float(np.sqrt(np.dot(gap, gap)))


def primal_residual(p_j, Y_ij):
    """||p_j - Y_ij||_F: how far agent j's own position trajectory is from the copy its neighbours hold."""
    return _gap_norm(p_j, Y_ij)


def dual_residual(Y_new, Y_old):
    """||Y_new - Y_old||_F: how far one consensus update moved the shared copy."""
    return _gap_norm(Y_new, Y_old)


def update_rho_admm(rho, primal_res, dual_res, mu=10.0, tau_inc=2.0, tau_dec=2.0):
    """Residual balancing: rho grows by tau_inc when the primal residual dominates the dual one by more than the factor mu,
    shrinks by tau_dec in the opposite case, and is left alone in between."""
    grow = primal_res > mu * dual_res
    shrink = (not grow) and dual_res > mu * primal_res
    return rho * tau_inc if grow else (rho / tau_dec if shrink else rho)
