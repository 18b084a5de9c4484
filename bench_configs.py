"""Synthetic scenes of the five BASELINE.json configs.  Plain numpy + the package's model classes; shared by
bench.py, the tools/ runners and the tests so that every number is quoted on the same workload."""
import numpy as np

K2, M2 = 100, 8


def random_unicycle_scene(rng, M=M2):
    """Config 2 scene: start on the left edge band, goal mirrored, M random discs clear of both."""
    y0 = rng.uniform(-9, 9)
    start = np.array([rng.uniform(-9, -8), y0, 0.0])
    goal = np.array([-start[0], -y0, 0.0])
    obs = []
    while len(obs) < M:
        c = rng.uniform(-7, 7, 2); r = rng.uniform(0.5, 2.0)
        if min(np.linalg.norm(c - start[:2]), np.linalg.norm(c - goal[:2])) > r + 0.5 + 0.5:
            obs.append((list(c), float(r)))
    return start, goal, obs


def config2_scenes(n, seed):
    rng = np.random.default_rng(seed)
    return [random_unicycle_scene(rng) for _ in range(n)]


def config1_model():
    """Config 1: the shipped single agent (unicycle_model.py:27-49), K = 50."""
    from scvx_b200.models.unicycle_model import UnicycleModel
    return UnicycleModel(), 50


def config2_models(n, seed):
    from scvx_b200.models.unicycle_model import UnicycleModel
    return [UnicycleModel(r_init=s, r_final=g, obstacles=o) for s, g, o in config2_scenes(n, seed)], K2


def config3_models():
    """Config 3: 16 unicycle agents on a circle of radius 8 swapping to antipodes, one disc ([0,0],1), K=100, d_min=0.5."""
    from scvx_b200.models.unicycle_model import UnicycleModel
    ang = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    models = [UnicycleModel(r_init=np.array([8 * np.cos(a), 8 * np.sin(a), 0.0]), r_final=np.array([-8 * np.cos(a), -8 * np.sin(a), 0.0]),
                            obstacles=[([0.0, 0.0], 1.0)]) for a in ang]
    return {"name": "config3: 16 unicycle agents on a circle r=8 -> antipodes, K=100, d_min=0.5, rho=1, all pairs", "models": models,
            "K": 100, "d_min": 0.5, "sigma": 20.0, "si": False, "d": 2, "kw": {}}


def config4_models(N=256):
    """Config 4 (i): N single-integrator agents on a Fibonacci sphere of radius 8 -> antipodes, obstacle ([0,0,0],1), K=100."""
    from scvx_b200.models.single_integrator_model import SingleIntegratorModel
    i = np.arange(N) + 0.5
    phi = np.arccos(1 - 2 * i / N); th = np.pi * (1 + 5 ** 0.5) * i
    pts = 8 * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
    models = [SingleIntegratorModel(r_init=q, r_final=-q, obstacles=[([0.0, 0.0, 0.0], 1.0)]) for q in pts]
    return {"name": f"config4: {N} single-integrator agents on a Fibonacci sphere r=8 -> antipodes, K=100, d_min=0.5, rho=1, all pairs",
            "models": models, "K": 100, "d_min": 0.5, "sigma": 20.0, "si": True, "d": 3, "kw": {}}


def config5_models(N=8192, knn=16):
    """Config 5 (decentralised): N unicycle agents x K=200, M=32 discs per agent from a shared field of 512 (seed 1), each agent
    coupled to its `knn` nearest neighbours within 1.0 (documented deviation from all-pairs, needed at this N)."""
    from scvx_b200.models.unicycle_model import UnicycleModel
    K, M = 200, 32
    rng = np.random.default_rng(1)
    field_c = rng.uniform(-7, 7, (512, 2)); field_r = rng.uniform(0.3, 1.0, 512)
    models = []
    for _ in range(N):
        y0 = rng.uniform(-9, 9); start = np.array([rng.uniform(-9, -8), y0, 0.0]); goal = np.array([-start[0], -y0, 0.0])
        d = np.minimum(np.linalg.norm(field_c - start[:2], axis=1), np.linalg.norm(field_c - goal[:2], axis=1)) - field_r
        pick = rng.choice(np.where(d > 1.0)[0], M, replace=False)
        models.append(UnicycleModel(r_init=start, r_final=goal, obstacles=[(list(field_c[j]), float(field_r[j])) for j in pick],
                                    robot_radius=0.05))
    return {"name": f"config5: {N} unicycle agents x K={K}, M={M} of a shared field of 512 (seed 1), ADMM consensus with the {knn} nearest "
                    "neighbours within 1.0, d_min=0.1, rho=1", "models": models, "K": K, "d_min": 0.1, "sigma": 20.0, "si": False, "d": 2,
            "kw": {"neighbor_k": knn, "neighbor_radius": 1.0}}
