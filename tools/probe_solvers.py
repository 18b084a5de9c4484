"""Probe which of the reference's stage-3 solver packages import on this box.
Writes one JSON line per package; run under gpurun and commit the log under profiles/."""
import importlib
import json
import sys

out = {}
for name in ("cvxpy", "ecos", "clarabel", "osqp", "scs", "cvxopt", "qpsolvers", "jax", "matplotlib", "highspy", "scipy"):
    try:
        m = importlib.import_module(name)
        out[name] = {"ok": True, "version": getattr(m, "__version__", "?")}
    except Exception as e:  # noqa: BLE001
        out[name] = {"ok": False, "error": f"{type(e).__name__}: {e}"}
print(json.dumps({"python": sys.version.split()[0], "packages": out}, indent=1))
