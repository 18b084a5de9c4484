"""Warm-start generators and analysis metrics on the GPU (csrc/utils.cu through the C-ABI) against the oracle and the
reference's golden outputs.  Tolerances: 1e-12 absolute on warm starts (libm vs CUDA asin/atan2/sincos differ in the last
bits), bit-exact on the analysis metrics (sqrt and the no-FMA sums are correctly rounded on both sides)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utils_golden.npz"))


def _obs(tag):
    return [(list(c), float(r)) for c, r in zip(G[f"{tag}_obs_c"], G[f"{tag}_obs_r"])]


def test_unicycle_warm_start_vs_golden():
    from scvx_b200.utils.initial_guess import initial_guess
    worst = 0.0
    for i in range(int(G["n_uni"])):
        t = f"uni{i}"
        X0, U0 = initial_guess(G[f"{t}_p0"], G[f"{t}_p1"], _obs(t), 0.3, int(G[f"{t}_K"]))
        assert X0.shape == G[f"{t}_X0"].shape and U0.shape == G[f"{t}_U0"].shape
        worst = max(worst, np.abs(X0 - G[f"{t}_X0"]).max())
        assert not U0.any()
    assert worst < 1e-12, worst
    X0, _ = initial_guess([-8.0, -8.0, 0.0], [8.0, 8.0, 0.0], [([0.0, 0.0], 2.0)], 0.5, 50)
    assert np.abs(X0 - G["uni_sym_X0"]).max() < 1e-12


def test_si_warm_start_vs_golden():
    from scvx_b200.utils.IS_initial_guess import initial_guess
    worst = 0.0
    for i in range(int(G["n_si"])):
        t = f"si{i}"
        K = int(G[f"{t}_K"])
        X0, U0 = initial_guess(G[f"{t}_p0"], G[f"{t}_p1"], _obs(t), 0.3, K)
        worst = max(worst, np.abs(X0 - G[f"{t}_X0"]).max(), np.abs(U0 - G[f"{t}_U0"]).max() / (K - 1))
    assert worst < 1e-12, worst


def test_batched_warm_start_vs_oracle_ragged():
    """One launch, agents with 0..5 obstacles each (ragged), K = 100: against the oracle on the same seeded inputs."""
    from oracle import utils as ou
    from scvx_b200.utils.initial_guess import initial_guess_batch
    from scvx_b200.utils.IS_initial_guess import initial_guess_batch as si_batch
    rng = np.random.default_rng(5)
    for d, batch, orc in ((2, initial_guess_batch, ou.initial_guess_unicycle), (3, si_batch, ou.initial_guess_si)):
        p0s, p1s, obs = [], [], []
        for i in range(64):
            a = np.zeros(3); b = np.zeros(3)
            a[:d] = rng.uniform(-9, -7, d); b[:d] = rng.uniform(7, 9, d)
            o = [(rng.uniform(-5, 5, d), float(rng.uniform(0.4, 1.5))) for _ in range(i % 6)]
            p0s.append(a); p1s.append(b); obs.append(o)
        X0, U0 = batch(p0s, p1s, obs, 0.25, 100)
        X0 = X0.cpu().numpy(); U0 = U0.cpu().numpy()
        for i in range(64):
            Xo, Uo = orc(p0s[i], p1s[i], obs[i], 0.25, 100)
            assert np.abs(X0[i] - Xo).max() < 1e-12, (d, i)
            assert np.abs(U0[i] - Uo).max() < 1e-10, (d, i)
            assert np.array_equal(X0[i][:d, 0], p0s[i][:d]) and np.array_equal(X0[i][:d, -1], p1s[i][:d])


def test_warm_start_errors_like_the_reference():
    from scvx_b200.utils.initial_guess import initial_guess
    from scvx_b200.utils.IS_initial_guess import initial_guess as si_guess
    with pytest.raises(ValueError, match="inside/on circle"):
        initial_guess([0.5, 0.0, 0.0], [8.0, 0.0, 0.0], [([0.0, 0.0], 1.0)], 0.3, 50)
    with pytest.raises(ValueError, match="too close"):
        si_guess(np.array([0.0, 0.0, 0.0]), np.array([0.0, 0.0, 1e-7]), [(np.array([0.0, 0.0, 5e-8]), 1e-8)], 0.0, 50)


def test_analysis_metrics_bit_exact():
    from scvx_b200.utils.analysis import min_agent_obstacle_distance, min_inter_agent_distance
    X = list(G["an_X"])
    dmin, dmat = min_inter_agent_distance(X)
    omin, omat = min_agent_obstacle_distance(X, _obs("an"), float(G["an_robot_radius"]))
    assert dmin == float(G["an_dmin"]) and np.array_equal(dmat, G["an_dmat"])
    assert omin == float(G["an_omin"]) and np.array_equal(omat, G["an_omat"])


def test_analysis_metrics_large_vs_oracle():
    import torch
    from oracle import utils as ou
    from scvx_b200 import _device
    rng = np.random.default_rng(11)
    N, K, M = 257, 200, 33
    X = rng.normal(size=(N, 3, K)) * 4.0
    C = rng.uniform(-3, 3, (M, 3)); R = rng.uniform(0.2, 1.0, M)
    dmin, dmat = _device.min_inter_agent_distance(torch.as_tensor(X, device="cuda"))
    omin, omat = _device.min_agent_obstacle_distance(torch.as_tensor(X, device="cuda"), torch.as_tensor(C, device="cuda"),
                                                     torch.as_tensor(R, device="cuda"), 0.4)
    rd, rm = ou.min_inter_agent_distance(list(X))
    od, om = ou.min_agent_obstacle_distance(list(X), [(list(c), float(r)) for c, r in zip(C, R)], 0.4)
    assert np.array_equal(dmat.cpu().numpy(), rm) and float(dmin.item()) == rd
    assert np.array_equal(omat.cpu().numpy(), om) and float(omin.item()) == od
    assert np.allclose(dmat.cpu().numpy(), dmat.cpu().numpy().T) and not np.diag(dmat.cpu().numpy()).any()


def test_empty_batches_and_bad_arguments():
    """n = 0 is a no-op for every new entry point; argument errors come back as SCVX_E_BADARG with a message."""
    import ctypes
    import torch
    from scvx_b200 import _device, _lib
    dev = torch.device("cuda")
    z = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)   # noqa: E731
    X0, U0, st = _device.warm_start(_lib.MODEL_UNICYCLE, z(0, 3), z(0, 3), None, None, 0.3, 20)
    assert X0.shape == (0, 3, 20) and U0.shape == (0, 2, 20) and st.numel() == 0
    dmin, dmat = _device.min_inter_agent_distance(z(1, 3, 7).zero_())          # one agent: no pair, +inf minimum
    assert dmat.shape == (1, 1) and np.isinf(dmin.item())
    nr, ts, h0, gx = _device.intersample(_lib.MODEL_UNICYCLE, z(0, 3, 9), z(0, 2, 9), z(0), z(0, 2, 2), z(0, 2))
    assert nr.shape == (0, 8, 2)
    lib = _lib.load()
    assert lib.scvx_warm_start_batched(99, 1, 20, 0, None, None, None, None, None, 0.0, None, None, None, None) == -1
    assert b"null pointer" in lib.scvx_last_error() or b"model_id" in lib.scvx_last_error()
    assert lib.scvx_min_inter_agent_distance(2, 5, 3, 4, None, None, None, None) == -1      # n_rows > n_x
    assert lib.scvx_intersample_batched(0, 1, 5, 1, 2, None, None, None, None, None, 1.0, 1, 1e-4, 1e-6, 4, None, None, None,
                                        None, None) == -1                                     # num_samples < 2
    with pytest.raises(ValueError):                                                          # all distances zero -> no positive entry
        from scvx_b200.utils.analysis import min_inter_agent_distance
        min_inter_agent_distance([np.zeros((3, 4)), np.zeros((3, 4))])
