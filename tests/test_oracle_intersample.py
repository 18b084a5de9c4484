"""oracle/intersample.py against golden outputs of the UNMODIFIED reference (tests/golden/intersample_golden.npz)."""
import os

import numpy as np

from oracle import foh as ofoh, intersample as oi, models as omodels

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "intersample_golden.npz"))
K = int(G["K"])


def _case(i):
    kind = str(G[f"c{i}_kind"])
    m = omodels.unicycle() if kind == "uni" else omodels.single_integrator()
    mdim = 2 if kind == "uni" else 3
    return m, mdim, ofoh.OracleFOH(m, K)


def test_reference_tolerance_reproduces_the_reference():
    """Same algorithm, same odeint defaults: identical roots and linearisations."""
    for i in range(0, int(G["n"]), 3):
        m, mdim, foh = _case(i)
        f, _ = oi.make_segment_f(foh, G[f"c{i}_u0"], G[f"c{i}_u1"], float(G[f"c{i}_sigma"]), tol="reference")
        obstacle = (G[f"c{i}_pc"], float(G[f"c{i}_r"]))
        T = np.eye(3)[:mdim]
        ts = oi.find_critical_times(G[f"c{i}_xk"], G[f"c{i}_u0"], f, T, obstacle, dt=1.0)
        assert len(ts) == len(G[f"c{i}_ts"])
        np.testing.assert_allclose(ts, G[f"c{i}_ts"], rtol=0, atol=1e-12)
        for q, t in enumerate(ts):
            h0, gx, gu = oi.linearize_h(G[f"c{i}_xk"], G[f"c{i}_u0"], t, f, T, obstacle)
            assert abs(h0 - G[f"c{i}_h0"][q]) < 1e-12
            np.testing.assert_allclose(gx, G[f"c{i}_gx"][q], rtol=0, atol=1e-9)
            assert not gu.any() and not G[f"c{i}_gu"][q].any()      # the segment flow ignores its control argument


def test_tight_flow_agrees_with_the_reference_flow_to_odeint_accuracy():
    for i in range(int(G["n"])):
        m, mdim, foh = _case(i)
        f, _ = oi.make_segment_f(foh, G[f"c{i}_u0"], G[f"c{i}_u1"], float(G[f"c{i}_sigma"]), tol="tight")
        flow = np.array([f(G[f"c{i}_xk"], None, t) for t in (0.0, 0.25, 0.7, 1.0)])
        assert np.abs(flow - G[f"c{i}_flow"]).max() < 5e-7
