"""ctypes binding of libscvx_b200.so -- the only way the Python mirror reaches the GPU.

There is deliberately NO CPU fallback: if the shared object is missing or a call fails, this module
raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported from here.)
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SCVX_LIB selects an instrumented build of the same sources (e.g. -DSCVX_PHASE_TIMING, tools/phase_timing.py); never a fallback
LIB_PATH = os.environ.get("SCVX_LIB") or os.path.join(HERE, "libscvx_b200.so")

MODEL_UNICYCLE = 0
MODEL_SINGLE_INTEGRATOR = 1
MODEL_USER_BASE = 16          # ids of models registered at run time (codegen.register)

ST_OPTIMAL, ST_MAXITER, ST_NUMERICAL = 0, 1, 2

_c_dp = ctypes.c_void_p
_c_int = ctypes.c_int
_c_dbl = ctypes.c_double


class SolveArgs(ctypes.Structure):
    """Mirror of `scvx_solve_args` (include/scvx_b200.h)."""
    _fields_ = [
        ("model_id", _c_int), ("n_agents", _c_int), ("K", _c_int), ("M", _c_int), ("n_nbr", _c_int),
        ("max_iter", _c_int), ("norm1_induced", _c_int),
        ("A_bar", _c_dp), ("B_bar", _c_dp), ("C_bar", _c_dp), ("S_bar", _c_dp), ("z_bar", _c_dp),
        ("X_ref", _c_dp), ("U_ref", _c_dp), ("sigma_ref", _c_dp), ("tr_radius", _c_dp),
        ("x_init", _c_dp), ("x_final", _c_dp), ("pos_lo", _c_dp), ("pos_hi", _c_dp), ("v_max", _c_dp), ("w_max", _c_dp),
        ("obs_a", _c_dp), ("obs_b", _c_dp), ("col_a", _c_dp), ("col_b", _c_dp), ("col_mask", _c_dp),
        ("quad_rho", _c_dp), ("lin_p", _c_dp),
        ("weight_nu", _c_dbl), ("weight_slack", _c_dbl), ("weight_sigma", _c_dbl), ("weight_col", _c_dbl),
        ("X", _c_dp), ("U", _c_dp), ("nu", _c_dp), ("sigma", _c_dp), ("s_prime", _c_dp), ("col_slack", _c_dp),
        ("objective", _c_dp), ("status", _c_dp), ("iters", _c_dp),
        ("workspace", _c_dp), ("workspace_bytes", ctypes.c_ulonglong),
        ("quad_diag", _c_dp), ("lin_w", _c_dp), ("quad_pair", _c_dp), ("fix_sigma", _c_int), ("block_order", _c_dp), ("mu0", _c_dp),
        ("active", _c_dp), ("retry_failed", _c_int), ("retry_list", _c_dp),
    ]


class LtiArgs(ctypes.Structure):
    """Mirror of `scvx_lti_args` (include/scvx_b200.h)."""
    _fields_ = [
        ("n_robots", _c_int), ("T", _c_int), ("n", _c_int), ("m", _c_int), ("nq", _c_int), ("max_iter", _c_int),
        ("Ad", _c_dp), ("Bd", _c_dp), ("x", _c_dp), ("u", _c_dp), ("x_des", _c_dp),
        ("tr", _c_dbl), ("c_w", _c_dbl), ("rho", _c_dbl), ("c_S", _c_dbl),
        ("box_lo0", _c_dbl), ("box_hi0", _c_dbl), ("box_lo1", _c_dbl), ("box_hi1", _c_dbl),
        ("lin", _c_dp), ("sbar", _c_dp), ("col_h", _c_dp), ("col_g", _c_dp),
        ("d", _c_dp), ("w", _c_dp), ("S", _c_dp), ("objective", _c_dp), ("status", _c_dp), ("iters", _c_dp),
        ("workspace", _c_dp), ("workspace_bytes", ctypes.c_ulonglong),
    ]


# name -> (restype, argtypes); must list EVERY symbol include/scvx_b200.h declares
# (tests/test_capi_symbols.py parses the header and checks).
SIGNATURES = {
    "scvx_abi_version": (_c_int, []),
    "scvx_last_error": (ctypes.c_char_p, []),
    "scvx_model_dims": (_c_int, [_c_int, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int), ctypes.POINTER(_c_int)]),
    "scvx_user_model_register": (_c_int, [ctypes.c_char_p, _c_int, _c_int, _c_int, ctypes.c_char_p, ctypes.POINTER(_c_int)]),
    "scvx_user_model_log": (ctypes.c_char_p, []),
    "scvx_foh_batched": (_c_int, [_c_int, _c_int, _c_int, _c_int] + [_c_dp] * 8 + [_c_dp]),
    "scvx_integrate_piecewise_batched": (_c_int, [_c_int, _c_int, _c_int, _c_int] + [_c_dp] * 4 + [_c_dp]),
    "scvx_integrate_full_batched": (_c_int, [_c_int, _c_int, _c_int, _c_int] + [_c_dp] * 4 + [_c_dp]),
    "scvx_linearize_obstacles_batched": (_c_int, [_c_int, _c_int, _c_int, _c_int] + [_c_dp] * 5 + [_c_dp]),
    "scvx_linearize_collision_batched": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_dbl] + [_c_dp] * 4 + [_c_dp]),
    "scvx_cross_min_dist2": (_c_int, [_c_int, _c_int, _c_int, _c_int] + [_c_dp] * 3 + [_c_dp]),
    "scvx_linearize_collision_indexed": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_dbl] + [_c_dp] * 5 + [_c_dp]),
    "scvx_slab_normals_batched": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int] + [_c_dp] * 7 + [_c_dp]),
    "scvx_order_by_iters": (_c_int, [_c_int, _c_dp, _c_dp, _c_dp]),
    "scvx_mu0_from_iters": (_c_int, [_c_int, _c_dp, _c_int, ctypes.c_double, ctypes.c_double, _c_dp, _c_dp]),
    "scvx_solve_workspace_bytes": (ctypes.c_ulonglong, [_c_int, _c_int, _c_int, _c_int, _c_int]),
    "scvx_solve_batched": (_c_int, [ctypes.POINTER(SolveArgs), _c_dp]),
    "scvx_consensus_update": (_c_int, [_c_int, _c_int, _c_int, _c_dbl] + [_c_dp] * 5 + [_c_dp]),
    "scvx_consensus_update_x": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_dbl] + [_c_dp] * 5 + [_c_dp]),
    "scvx_admm_round_prep": (_c_int, [_c_int] * 6 + [_c_dbl, _c_dbl] + [_c_dp] * 12 + [_c_dp]),
    "scvx_knn_select": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_dbl, _c_dp, _c_dp, _c_dp]),
    "scvx_radius_mask": (_c_int, [_c_int, _c_int, _c_dbl, _c_dp, _c_dp, _c_dp]),
    "scvx_outer_update": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_dbl] + [_c_dp] * 11 + [_c_dp]),
    "scvx_lti_qp_workspace_bytes": (ctypes.c_ulonglong, [_c_int, _c_int, _c_int, _c_int]),
    "scvx_lti_qp_batched": (_c_int, [ctypes.POINTER(LtiArgs), _c_dp]),
    "scvx_sbar_qp_batched": (_c_int, [_c_int, _c_int, _c_int, _c_dbl, _c_dbl] + [_c_dp] * 6 + [_c_dp, ctypes.c_ulonglong, _c_dp]),
    "scvx_warm_start_batched": (_c_int, [_c_int, _c_int, _c_int, _c_int] + [_c_dp] * 5 + [_c_dbl] + [_c_dp] * 3 + [_c_dp]),
    "scvx_min_inter_agent_distance": (_c_int, [_c_int, _c_int, _c_int, _c_int] + [_c_dp] * 3 + [_c_dp]),
    "scvx_min_agent_obstacle_distance": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int] + [_c_dp] * 3 + [_c_dbl] + [_c_dp] * 2 + [_c_dp]),
    "scvx_intersample_batched": (_c_int, [_c_int] * 5 + [_c_dp] * 5 + [_c_dbl, _c_int, _c_dbl, _c_dbl, _c_int] + [_c_dp] * 4 + [_c_dp]),
    "scvx_clearance_samples_batched": (_c_int, [_c_int] * 5 + [_c_dp] * 6 + [_c_dp]),
    "scvx_probe_fp64": (_c_int, [_c_int, _c_int, _c_dp, ctypes.POINTER(_c_dbl), _c_dp]),
    "scvx_l2_flush": (_c_int, [_c_dp, ctypes.c_ulonglong, _c_dp]),
    "scvx_debug_phase_cycles": (_c_int, [ctypes.POINTER(ctypes.c_ulonglong), _c_int]),
}

_lib = None


class ScvxError(RuntimeError):
    pass


def load():
    """Load the shared object (once).  Raises ScvxError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScvxError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.scvx_abi_version() != 1:
        raise ScvxError("libscvx_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().scvx_last_error().decode("utf-8", "replace")
        raise ScvxError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch CUDA tensor (contiguous) as a void* (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ScvxError("expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise ScvxError("expected a contiguous tensor")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
