"""Golden vectors for the inter-sample clearance routines from the UNMODIFIED reference
(SCvx/utils/intersample_collision.py + FirstOrderHold._dx).   python tests/golden/make_golden_intersample.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402


def main():
    K = 20
    refshim.load(K)
    from SCvx.discretization.first_order_hold import FirstOrderHold
    from SCvx.models.single_integrator_model import SingleIntegratorModel
    from SCvx.models.unicycle_model import UnicycleModel
    from SCvx.utils import intersample_collision as ref
    rng = np.random.default_rng(20261020)
    out = {"K": K}
    idx = 0
    for kind, Model in (("uni", UnicycleModel), ("si", SingleIntegratorModel)):
        m = Model()
        foh = FirstOrderHold(m, K)
        n_u = m.n_u
        for sigma in (1.0, 30.0):
            for _ in range(6):
                xk = rng.uniform(-1, 1, 3)
                if kind == "uni":
                    u0 = np.array([rng.uniform(0.3, 1.0), rng.uniform(-0.5, 0.5)]); u1 = np.array([rng.uniform(0.3, 1.0), rng.uniform(-0.5, 0.5)])
                    mdim = 2
                else:
                    u0 = rng.uniform(-1, 1, 3); u1 = rng.uniform(-1, 1, 3)
                    mdim = 3
                f_seg, dt_phys = ref.make_segment_f(foh, u0, u1, sigma)
                # put the obstacle beside the middle of the segment so that the clearance has an interior minimum
                mid = f_seg(xk, None, 0.5)[:mdim]
                end = f_seg(xk, None, 1.0)[:mdim]
                d = end - xk[:mdim]
                nrm = np.linalg.norm(d)
                perp = np.cross(np.append(d, 0.0)[:3], [0.0, 0.0, 1.0])[:mdim] if mdim == 2 else np.cross(d, rng.normal(size=3))
                perp = perp / (np.linalg.norm(perp) + 1e-12)
                p_c = mid + perp * max(0.3 * nrm, 0.02) + 0.1 * d * rng.uniform(-1, 1)
                obstacle = (p_c, 0.01)
                T = np.eye(3)[:mdim]
                ts = ref.find_critical_times(xk, u0, f_seg, T, obstacle, dt=1.0)
                out[f"c{idx}_kind"] = kind; out[f"c{idx}_sigma"] = sigma; out[f"c{idx}_xk"] = xk
                out[f"c{idx}_u0"] = u0; out[f"c{idx}_u1"] = u1; out[f"c{idx}_pc"] = p_c; out[f"c{idx}_r"] = 0.01
                out[f"c{idx}_ts"] = np.array(ts)
                lin = [ref.linearize_h(xk, u0, t, f_seg, T, obstacle) for t in ts]
                out[f"c{idx}_h0"] = np.array([l[0] for l in lin]); out[f"c{idx}_gx"] = np.array([l[1] for l in lin]).reshape(len(ts), 3)
                out[f"c{idx}_gu"] = np.array([l[2] for l in lin]).reshape(len(ts), n_u)
                out[f"c{idx}_flow"] = np.array([f_seg(xk, None, t) for t in (0.0, 0.25, 0.7, 1.0)])
                idx += 1
    out["n"] = idx
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "intersample_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, idx, "cases; roots per case:", [len(out[f"c{i}_ts"]) for i in range(idx)])


if __name__ == "__main__":
    main()
