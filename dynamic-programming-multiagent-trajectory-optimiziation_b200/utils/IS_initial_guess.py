"""Warm start for the 3-D single integrator -- same call as SCvx/utils/IS_initial_guess.py:87-126 (warm_start_kernel<3>).

`initial_guess(p0, p1, obstacles, clearance, K) -> (X0 (3, K), U0 (3, K))`; `initial_guess_batch` for many agents."""
from .. import _lib
from .initial_guess import _run

_MODEL = _lib.MODEL_SINGLE_INTEGRATOR
_D = 3


def initial_guess_batch(p0s, p1s, obstacle_lists, clearance, K):
    return _run(_MODEL, _D, p0s, p1s, obstacle_lists, clearance, K)


def initial_guess(p0, p1, obstacles, clearance, K):
    X0, U0 = _run(_MODEL, _D, [p0], [p1], [list(obstacles)], clearance, K)
    return X0[0].cpu().numpy(), U0[0].cpu().numpy()
