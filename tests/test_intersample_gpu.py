"""Inter-sample clearance on the GPU (csrc/intersample.cu) against the oracle at tight integration tolerance and against
the reference's golden outputs.  The reference differentiates a default-tolerance odeint numerically, so its own t* / gradients
carry noise: the golden comparison is loose (1e-3), the oracle comparison tight."""
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import foh as ofoh, intersample as oi, models as omodels

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "intersample_golden.npz"))
K = int(G["K"])


def _mirror_case(i):
    from scvx_b200.discretization.first_order_hold import FirstOrderHold
    from scvx_b200.models.single_integrator_model import SingleIntegratorModel
    from scvx_b200.models.unicycle_model import UnicycleModel
    kind = str(G[f"c{i}_kind"])
    m = UnicycleModel() if kind == "uni" else SingleIntegratorModel()
    return FirstOrderHold(m, K), (2 if kind == "uni" else 3)


def test_segment_flow_vs_golden_and_tight_oracle(cuda):
    from scvx_b200.utils.intersample_collision import make_segment_f
    for i in range(int(G["n"])):
        foh, mdim = _mirror_case(i)
        f, dt_phys = make_segment_f(foh, G[f"c{i}_u0"], G[f"c{i}_u1"], float(G[f"c{i}_sigma"]))
        assert dt_phys == foh.dt * float(G[f"c{i}_sigma"])
        flow = np.array([f(G[f"c{i}_xk"], None, t) for t in (0.0, 0.25, 0.7, 1.0)])
        assert np.abs(flow - G[f"c{i}_flow"]).max() < 5e-7            # the reference's own odeint accuracy
        om = omodels.unicycle() if mdim == 2 else omodels.single_integrator()
        ft, _ = oi.make_segment_f(ofoh.OracleFOH(om, K), G[f"c{i}_u0"], G[f"c{i}_u1"], float(G[f"c{i}_sigma"]), tol="tight")
        tight = np.array([ft(G[f"c{i}_xk"], None, t) for t in (0.0, 0.25, 0.7, 1.0)])
        assert np.abs(flow - tight).max() < 1e-9 * max(1.0, np.abs(tight).max())


def test_critical_times_and_linearisation(cuda):
    from scvx_b200.utils.intersample_collision import find_critical_times, linearize_h, make_segment_f
    for i in range(int(G["n"])):
        foh, mdim = _mirror_case(i)
        xk, u0, u1, sigma = G[f"c{i}_xk"], G[f"c{i}_u0"], G[f"c{i}_u1"], float(G[f"c{i}_sigma"])
        obstacle = (G[f"c{i}_pc"], float(G[f"c{i}_r"]))
        T = np.eye(3)[:mdim]
        f, _ = make_segment_f(foh, u0, u1, sigma)
        ts = find_critical_times(xk, u0, f, T, obstacle, dt=1.0)
        assert len(ts) == len(G[f"c{i}_ts"]), (i, ts, G[f"c{i}_ts"])
        np.testing.assert_allclose(ts, G[f"c{i}_ts"], rtol=0, atol=1e-3)          # vs the reference (noisy)
        om = omodels.unicycle() if mdim == 2 else omodels.single_integrator()
        ft, _ = oi.make_segment_f(ofoh.OracleFOH(om, K), u0, u1, sigma, tol="tight")
        to = oi.find_critical_times(xk, u0, ft, T, obstacle, dt=1.0)
        np.testing.assert_allclose(ts, to, rtol=0, atol=2e-6)                     # vs the oracle: the bisection tolerance
        for q, t in enumerate(ts):
            h0, gx, gu = linearize_h(xk, u0, t, f, T, obstacle)
            ho, gxo, _ = oi.linearize_h(xk, u0, t, ft, T, obstacle)
            assert abs(h0 - ho) < 1e-9 and np.abs(gx - gxo).max() < 1e-5 and not gu.any()
            assert abs(h0 - G[f"c{i}_h0"][q]) < 1e-5 and np.abs(gx - G[f"c{i}_gx"][q]).max() < 5e-3


def test_batched_kernel_matches_single_calls(cuda):
    """All (agent, segment, obstacle) of a batch in one launch == the per-segment calls."""
    from scvx_b200 import _device, _lib
    from scvx_b200.discretization.first_order_hold import FirstOrderHold
    from scvx_b200.models.unicycle_model import UnicycleModel
    from scvx_b200.utils.intersample_collision import find_critical_times, make_segment_f
    rng = np.random.default_rng(9)
    n, Kb, M = 3, 12, 2
    X = rng.uniform(-1, 1, (n, 3, Kb)); U = np.stack([rng.uniform(0.3, 1.0, (n, Kb)), rng.uniform(-0.5, 0.5, (n, Kb))], axis=1)
    C = rng.uniform(-1.5, 1.5, (n, M, 2)); R = np.full((n, M), 0.05)
    sig = np.full(n, 25.0)
    nr, ts, h0, gx = _device.intersample(_lib.MODEL_UNICYCLE, helpers.to_dev(X, cuda), helpers.to_dev(U, cuda), helpers.to_dev(sig, cuda),
                                         helpers.to_dev(C, cuda), helpers.to_dev(R, cuda))
    nr, ts = nr.cpu().numpy(), ts.cpu().numpy()
    foh = FirstOrderHold(UnicycleModel(), Kb)
    total = 0
    for a in range(n):
        for k in range(0, Kb - 1, 3):
            for j in range(M):
                f, _ = make_segment_f(foh, U[a, :, k], U[a, :, k + 1], 25.0)
                single = find_critical_times(X[a, :, k], U[a, :, k], f, np.eye(3)[:2], (C[a, j], 0.05), dt=1.0)
                assert nr[a, k, j] == len(single)
                np.testing.assert_array_equal(ts[a, k, j, :len(single)], single)
                total += len(single)
    assert total > 0
    hs = _device.clearance_samples(_lib.MODEL_UNICYCLE, helpers.to_dev(X, cuda), helpers.to_dev(U, cuda), helpers.to_dev(sig, cuda),
                                   helpers.to_dev(C[:, 0], cuda), helpers.to_dev(R[:, 0] + 0.5, cuda), resolution=10).cpu().numpy()
    # t = 0 samples are the knot clearances
    knot = np.linalg.norm(X[:, :2, :-1] - C[:, 0, :, None], axis=1) - (R[:, 0, None] + 0.5)
    np.testing.assert_allclose(hs[:, :, 0], knot, rtol=0, atol=1e-14)
