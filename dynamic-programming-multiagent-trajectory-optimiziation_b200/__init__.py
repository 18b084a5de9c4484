"""scvx_b200 -- B200-native SCvx inner loop (FOH discretisation -> constraint linearisation -> convex
sub-problem) behind the reference's own Python interface.

Drop-in mirror of the hot-path classes of
shiivashaakeri/Dynamic-Programming-MultiAgent-Trajectory-Optimiziation (package `SCvx`):

    SCvx.discretization.first_order_hold.FirstOrderHold   -> scvx_b200.discretization.first_order_hold
    SCvx.models.{unicycle,single_integrator}_model        -> scvx_b200.models.*
    SCvx.models.{multi_agent_model,SI_multi_agent_model}  -> scvx_b200.models.*
    SCvx.optimization.{sc_problem,scvx_solver}            -> scvx_b200.optimization.*
    SCvx.optimization.{agent_solver,admm_coordinator,...} -> scvx_b200.optimization.*

All numerics run in hand-written sm_100a CUDA kernels reached through the C-ABI in
include/scvx_b200.h (ctypes, `_lib.py`).  There is no CPU fallback.

The directory is named after the upstream repository
(`dynamic-programming-multiagent-trajectory-optimiziation_b200`), which is not a valid Python
identifier; import it as `scvx_b200` (a two-line alias package at the repo root).
"""
__version__ = "0.1.0"
