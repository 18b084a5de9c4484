"""Import the upstream reference's own Python modules (stages 1-2) -- TEST INFRASTRUCTURE ONLY.

Only usable where the reference checkout exists (the build container); never on the GPU box.
Used by tests/golden/make_golden.py and by the optional cross-checks in tests/ (skipped when the
checkout is absent).
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("SCVX_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "SCvx"))


def load(K: int):
    """Return the reference's `SCvx` package with global_parameters.K patched to K *before* any
    other SCvx module binds it (15 modules copy K at import time)."""
    if not available():
        raise RuntimeError("reference checkout not present")
    stub = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cvxpy_stub")
    for p in (REFERENCE_ROOT, stub):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in [m for m in sys.modules if m == "SCvx" or m.startswith("SCvx.")]:
        del sys.modules[name]
    gp = importlib.import_module("SCvx.global_parameters")
    gp.K = K
    return importlib.import_module("SCvx")
