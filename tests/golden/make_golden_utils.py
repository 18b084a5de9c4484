"""Golden vectors for the warm-start generators and analysis metrics, from the UNMODIFIED reference functions.

Run once in the build container (where /root/reference exists):   python tests/golden/make_golden_utils.py
Writes tests/golden/utils_golden.npz.  Inputs are stored next to the outputs so the tests need nothing else.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402


def scenes(rng, n, d, M):
    out = []
    while len(out) < n:
        p0 = np.zeros(3); p1 = np.zeros(3)
        p0[:d] = rng.uniform(-9, -6, d); p1[:d] = rng.uniform(6, 9, d)
        if d == 2:
            p0[2] = rng.uniform(-1, 1); p1[2] = rng.uniform(-1, 1)
        obs = []
        for _ in range(M):
            c = rng.uniform(-5, 5, d); r = rng.uniform(0.5, 2.0)
            obs.append((c, r))
        # keep start/goal strictly outside every inflated obstacle (otherwise the reference raises)
        if all(np.linalg.norm(p0[:d] - c) > r + 0.6 and np.linalg.norm(p1[:d] - c) > r + 0.6 for c, r in obs):
            out.append((p0, p1, obs))
    return out


def main():
    refshim.load(50)
    from SCvx.utils import IS_initial_guess as ref_si
    from SCvx.utils import analysis as ref_an
    from SCvx.utils import initial_guess as ref_uni

    rng = np.random.default_rng(20261019)
    out = {}
    # ---- unicycle warm starts: K in {50, 100}, M in {0, 1, 3, 6}
    idx = 0
    for K in (50, 100):
        for M in (0, 1, 3, 6):
            for p0, p1, obs in scenes(rng, 4, 2, M):
                X0, U0 = ref_uni.initial_guess(p0, p1, [(list(c), r) for c, r in obs], 0.3, K)
                out[f"uni{idx}_p0"] = p0; out[f"uni{idx}_p1"] = p1
                out[f"uni{idx}_obs_c"] = np.array([c for c, _ in obs]).reshape(M, 2)
                out[f"uni{idx}_obs_r"] = np.array([r for _, r in obs])
                out[f"uni{idx}_K"] = K; out[f"uni{idx}_X0"] = X0; out[f"uni{idx}_U0"] = U0
                idx += 1
    out["n_uni"] = idx
    # the shipped example layout (run_multi_agent_admm.py): straight through a centred disc -- symmetric candidates
    p0 = np.array([-8.0, -8.0, 0.0]); p1 = np.array([8.0, 8.0, 0.0])
    X0, U0 = ref_uni.initial_guess(p0, p1, [([0.0, 0.0], 2.0)], 0.5, 50)
    out["uni_sym_X0"] = X0
    # ---- single-integrator warm starts
    idx = 0
    for K in (50, 100):
        for M in (0, 1, 2, 4):
            for p0, p1, obs in scenes(rng, 4, 3, M):
                X0, U0 = ref_si.initial_guess(p0, p1, [(c, r) for c, r in obs], 0.3, K)
                out[f"si{idx}_p0"] = p0; out[f"si{idx}_p1"] = p1
                out[f"si{idx}_obs_c"] = np.array([c for c, _ in obs]).reshape(M, 3)
                out[f"si{idx}_obs_r"] = np.array([r for _, r in obs])
                out[f"si{idx}_K"] = K; out[f"si{idx}_X0"] = X0; out[f"si{idx}_U0"] = U0
                idx += 1
    out["n_si"] = idx
    # ---- analysis metrics
    N, K = 7, 60
    X_list = [rng.normal(size=(3, K)) * 3.0 for _ in range(N)]
    X_list[3] = X_list[1].copy()            # a coincident pair: distance 0 must be ignored by the global minimum
    obstacles = [(list(rng.uniform(-2, 2, 3)), float(rng.uniform(0.3, 1.0))) for _ in range(4)]
    dmin, dmat = ref_an.min_inter_agent_distance(X_list)
    omin, omat = ref_an.min_agent_obstacle_distance(X_list, obstacles, 0.5)
    out["an_X"] = np.stack(X_list); out["an_obs_c"] = np.array([c for c, _ in obstacles])
    out["an_obs_r"] = np.array([r for _, r in obstacles]); out["an_robot_radius"] = 0.5
    out["an_dmin"] = dmin; out["an_dmat"] = dmat; out["an_omin"] = omin; out["an_omat"] = omat
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "utils_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
