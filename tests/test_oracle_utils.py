"""oracle/utils.py against golden outputs of the UNMODIFIED reference (tests/golden/utils_golden.npz)."""
import os

import numpy as np
import pytest

from oracle import utils as ou

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utils_golden.npz"))


def _obs(tag):
    return [(list(c), float(r)) for c, r in zip(G[f"{tag}_obs_c"], G[f"{tag}_obs_r"])]


def test_unicycle_warm_start_matches_reference():
    worst = 0.0
    for i in range(int(G["n_uni"])):
        t = f"uni{i}"
        X0, U0 = ou.initial_guess_unicycle(G[f"{t}_p0"], G[f"{t}_p1"], _obs(t), 0.3, int(G[f"{t}_K"]))
        assert X0.shape == G[f"{t}_X0"].shape and U0.shape == G[f"{t}_U0"].shape
        worst = max(worst, np.abs(X0 - G[f"{t}_X0"]).max())
        assert np.array_equal(U0, G[f"{t}_U0"])
    # numpy's norm / dot go through BLAS (possibly FMA), the restatement sums left to right: last-bit differences only
    assert worst < 1e-13, worst


def test_unicycle_symmetric_scene_picks_the_same_side():
    X0, _ = ou.initial_guess_unicycle([-8.0, -8.0, 0.0], [8.0, 8.0, 0.0], [([0.0, 0.0], 2.0)], 0.5, 50)
    assert np.abs(X0 - G["uni_sym_X0"]).max() < 1e-13


def test_si_warm_start_matches_reference():
    worst = 0.0
    for i in range(int(G["n_si"])):
        t = f"si{i}"
        X0, U0 = ou.initial_guess_si(G[f"{t}_p0"], G[f"{t}_p1"], _obs(t), 0.3, int(G[f"{t}_K"]))
        worst = max(worst, np.abs(X0 - G[f"{t}_X0"]).max(), np.abs(U0 - G[f"{t}_U0"]).max() / (int(G[f"{t}_K"]) - 1))
    assert worst < 1e-13, worst


def test_warm_start_errors():
    with pytest.raises(ValueError):      # start inside the inflated disc that the segment crosses
        ou.initial_guess_unicycle([0.5, 0.0, 0.0], [8.0, 0.0, 0.0], [([0.0, 0.0], 1.0)], 0.3, 50)
    with pytest.raises(ValueError):
        ou.detour_waypoints([0.0, 0.0, 0.0], [0.0, 0.0, 1e-9], [0.0, 0.0, 0.0], 1.0)


def test_analysis_metrics_match_reference_bit_for_bit():
    X = list(G["an_X"])
    obstacles = _obs("an")
    dmin, dmat = ou.min_inter_agent_distance(X)
    omin, omat = ou.min_agent_obstacle_distance(X, obstacles, float(G["an_robot_radius"]))
    assert dmin == float(G["an_dmin"]) and np.array_equal(dmat, G["an_dmat"])
    assert omin == float(G["an_omin"]) and np.array_equal(omat, G["an_omat"])
    assert dmat[1, 3] == 0.0 and dmin > 0.0          # the coincident pair is ignored
