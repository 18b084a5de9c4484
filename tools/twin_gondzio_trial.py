"""Negative result behind DESIGN 4.10: Gondzio centrality correctors on the numpy twin of the interior-point kernel
(oracle/ipm_struct.py), over sub-problems of the bench scenes (outer iterations 1..5 of the first N scenes, CPU oracle loop).
Prints, per variant, mean / max interior-point iterations, extra solves per problem and a cycle-cost model
(187 k cycles per iteration + 50 k per corrector: row pass + one-rhs solve + step-length pass, profiles/r01g).

    python tools/twin_gondzio_trial.py [n_scenes=40]        # ~3 min on 8 cores
"""
import os, sys
from multiprocessing import Pool
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
from oracle import foh as ofoh, subproblem as ospb, ipm_struct

K = 100
VARIANTS = [(0, 0.5, 0.3), (1, 0.5, 0.1), (1, 0.5, 0.3), (2, 0.5, 0.3), (1, 0.9, 0.3), (1, 0.3, 0.3)]


def problems_of(om):
    F = ofoh.OracleFOH(om, K)
    X, U = om.initialize_trajectory(K)
    sig, tr, out = 1.0, 100.0, []
    for it in range(6):
        p = ospb.Params(om, K, F.calculate_discretization(X, U, sig), X, U, sig, tr)
        r = ospb.solve(p)
        if it >= 1:
            out.append(p)
        X, U, sig, tr = r["X"], r["U"], r["sigma"], 50.0
    return out


def run(p):
    res = []
    for g, th, dl in VARIANTS:
        s = ipm_struct.StructIPM(p, max_iter=80, gondzio=g, g_thresh=th, g_delta=dl)
        r = s.solve()
        res.append((r["iters"], s.n_solves - 2 * r["iters"], r["status"]))
    return res


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(0)
    scenes = [helpers.random_unicycle_scene(rng, 8) for _ in range(n)]
    with Pool(os.cpu_count()) as pool:
        probs = [p for ps in pool.map(problems_of, scenes) for p in ps]
        res = np.array(pool.map(run, probs))
    print(f"{len(probs)} sub-problems")
    for v, (g, th, dl) in enumerate(VARIANTS):
        it, ex = res[:, v, 0], res[:, v, 1]
        print(f"correctors <= {g}, when min(alpha) < {th}, look-ahead {dl}: iterations mean {it.mean():.2f} max {it.max()}, "
              f"extra solves {ex.mean():.2f}/problem, cost {np.mean(it * 187 + ex * 50):.0f} k cycles, not optimal: {(res[:, v, 2] != 0).sum()}")
