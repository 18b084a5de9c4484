"""Golden vectors for the pure-numpy pieces of the two Distributed_opt scripts, produced by the UNMODIFIED reference code.

The scripts cannot be imported (they import cvxpy / jax and run their optimisation at import time), so the function
definitions `descete_f`, `x_initial`, `cost_fcn` and the module-level scenario assignments are lifted out of the reference
source with `ast` and executed as they stand, in a namespace holding numpy, numpy.linalg (LA) and scipy.signal.
    python tests/golden/make_golden_distopt.py      -> tests/golden/distopt_golden.npz"""
import ast
import os

import numpy as np
from numpy import linalg as LA
from scipy import signal

REF = os.environ.get("SCVX_REFERENCE_ROOT", "/root/reference")
WANT_FUNCS = {"descete_f", "x_initial", "cost_fcn"}
WANT_GLOBALS = {"Tf", "T0", "T", "t_traj", "dt", "n", "m", "N_agents", "robots_name", "R", "x_ini", "x_des", "trust_region", "max_iter",
                "count", "name"}


def lift(path):
    tree = ast.parse(open(path).read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANT_FUNCS:
            keep.append(node)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) and t.id in WANT_GLOBALS for t in node.targets):
            keep.append(node)
        elif isinstance(node, ast.For) and isinstance(node.target, (ast.Tuple, ast.Name)):
            names = {n.id for n in ast.walk(node) if isinstance(n, ast.Name)}
            if {"x_ini", "x_des"} & names and not ({"x_traj_opt", "cp", "plt", "X_traj"} & names):
                keep.append(node)
    ns = {"np": np, "LA": LA, "signal": signal}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


def main():
    out = {}
    rng = np.random.default_rng(20261021)
    for tag, fname in (("ad", "ADMM_decentralized.py"), ("d3", "dist_scvx_3d.py")):
        ns = lift(os.path.join(REF, "Distributed_opt", fname))
        Ad, Bd = ns["descete_f"](ns["dt"])
        X0 = ns["x_initial"](ns["x_ini"], ns["x_des"])
        out[f"{tag}_dt"] = ns["dt"]; out[f"{tag}_T"] = ns["T"]; out[f"{tag}_n"] = ns["n"]; out[f"{tag}_m"] = ns["m"]; out[f"{tag}_R"] = ns["R"]
        out[f"{tag}_names"] = np.array(ns["robots_name"])
        out[f"{tag}_Ad"] = np.asarray(Ad); out[f"{tag}_Bd"] = np.asarray(Bd)
        out[f"{tag}_x_ini"] = np.stack([ns["x_ini"][k] for k in ns["robots_name"]]).astype(float)
        out[f"{tag}_x_des"] = np.stack([ns["x_des"][k] for k in ns["robots_name"]]).astype(float)
        out[f"{tag}_X0"] = np.stack([X0[k] for k in ns["robots_name"]])
        Ad2, Bd2 = ns["descete_f"](0.37)
        out[f"{tag}_Ad_037"] = np.asarray(Ad2); out[f"{tag}_Bd_037"] = np.asarray(Bd2)
        if "cost_fcn" in ns:
            Xr = {k: rng.normal(size=X0[k].shape) for k in ns["robots_name"]}
            out[f"{tag}_cost_X"] = np.stack([Xr[k] for k in ns["robots_name"]]); out[f"{tag}_cost"] = ns["cost_fcn"](Xr)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "distopt_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, sorted(out))


if __name__ == "__main__":
    main()
