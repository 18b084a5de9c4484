"""Test helpers: build seeded sub-problems with the oracle and feed the same numbers to the GPU solver."""
import numpy as np
import torch

from oracle import foh as ofoh, models as omodels, subproblem as ospb


def model_id_of(om):
    from scvx_b200 import _lib
    return _lib.MODEL_UNICYCLE if om.kind == "unicycle" else _lib.MODEL_SINGLE_INTEGRATOR


def to_dev(a, dev, dtype=np.float64):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype)).to(dev)


def random_unicycle_scene(rng, M=8):
    """start on the left edge band, goal mirrored, M random discs."""
    y0 = rng.uniform(-9, 9)
    start = np.array([rng.uniform(-9, -8), y0, 0.0])
    goal = np.array([-start[0], -y0, 0.0])
    obs = []
    while len(obs) < M:
        c = rng.uniform(-7, 7, 2); r = rng.uniform(0.5, 2.0)
        if min(np.linalg.norm(c - start[:2]), np.linalg.norm(c - goal[:2])) > r + 0.5 + 0.5:
            obs.append((list(c), float(r)))
    return omodels.unicycle(r_init=start, r_final=goal, obstacles=obs)


def solve_batch_on_gpu(params, dev, max_iter=0, slab_penalty=1e8):
    """params: list of oracle Params with identical (model kind, K, M, number of neighbours).
    Returns the workspace (outputs as torch tensors)."""
    from scvx_b200 import _device
    p0 = params[0]
    om, K = p0.model, p0.K
    mid = model_id_of(om)
    n, M, NB, d = len(params), len(om.obstacles), len(p0.neighbors), om.d
    st = lambda f: to_dev(np.stack([f(p) for p in params]), dev)   # noqa: E731
    mats = tuple(st(lambda p, i=i: (p.A_bar, p.B_bar, p.C_bar, p.S_bar, p.z_bar)[i]) for i in range(5))
    obs_a = st(lambda p: p.obs_a.reshape(M, d, K)) if M else None
    obs_b = st(lambda p: p.obs_rhs[:, None] + np.einsum("mdk,md->mk", p.obs_a, p.obs_c)) if M else None
    col_a = col_b = quad = lin = None
    if NB:
        col_a = st(lambda p: np.stack([nb["a"] for nb in p.neighbors]))
        col_b = st(lambda p: np.stack([p.d_min + np.einsum("dk,dk->k", nb["a"], nb["Y"]) for nb in p.neighbors]))
        quad = st(lambda p: np.array(p.rho * len(p.neighbors)))
        lin = st(lambda p: sum(nb["Lam"] - p.rho * nb["Y"] for nb in p.neighbors))
    game_kw = {}
    if p0.game is not None:        # Nash best response: slab rows ride in the neighbour slots with an exact penalty
        NB = len(p0.game["rows"])
        if NB:
            col_a = st(lambda p: np.stack([a for a, _ in p.game["rows"]]))
            col_b = st(lambda p: np.stack([b for _, b in p.game["rows"]]))
        game_kw = dict(quad_diag=st(lambda p: p.game["qd"]), lin_w=st(lambda p: p.game["lw"]),
                       quad_pair=st(lambda p: p.game["qp"]), fix_sigma=p0.game["fix_sigma"])
    ws = _device.SubproblemWorkspace(mid, n, K, M, NB, dev)
    _device.solve_subproblem(
        ws, mats, st(lambda p: p.X_ref), st(lambda p: p.U_ref), st(lambda p: np.array(p.sigma_ref)),
        st(lambda p: np.array(p.tr_radius)), st(lambda p: p.model.x_init), st(lambda p: p.model.x_final),
        st(lambda p: np.array(p.model.lower_bound + p.model.robot_radius)),
        st(lambda p: np.array(p.model.upper_bound - p.model.robot_radius)),
        st(lambda p: np.array(p.model.v_max)), st(lambda p: np.array(p.model.w_max)),
        obs_a, obs_b, p0.weight_nu, p0.weight_slack, p0.weight_sigma,
        col_a=col_a, col_b=col_b, quad_rho=quad, lin_p=lin, weight_col=(slab_penalty if p0.game is not None else p0.weight_col),
        max_iter=max_iter, **game_kw)
    torch.cuda.synchronize()
    return ws


def make_problem_sequence(om, K, n_iter, tr0=100.0, **kw):
    """First n_iter sub-problems of the oracle's SCvx loop (each about the previous HiGHS solution)."""
    F = ofoh.OracleFOH(om, K)
    X, U = om.initialize_trajectory(K)
    sig, tr = 1.0, tr0
    out = []
    for _ in range(n_iter):
        mats = F.calculate_discretization(X, U, sig)
        p = ospb.Params(om, K, mats, X, U, sig, tr, **kw)
        r = ospb.solve(p)
        out.append((p, r))
        ev = ospb.evaluate(p, r["X"], r["U"], r["sigma"])
        tr = min(tr * (1.5 if ev["nu_norm"] < 1e-2 and ev["slack_sum"] < 1e-2 else 1.2), 50.0)
        X, U, sig = r["X"], r["U"], r["sigma"]
    return out


def config5_style_problem(rng, K=200, M=32, NB=16, sigma=25.0, d_min=0.5, rho=1.0):
    """One agent's ADMM sub-problem at BASELINE config-5 size: unicycle, K=200 nodes, M=32 discs drawn from a shared field,
    NB=16 neighbours (the compact k-nearest-neighbour coupling of BatchedADMM at N=8192)."""
    from oracle.models import linearize_collision
    field = [(list(rng.uniform(-7, 7, 2)), float(rng.uniform(0.3, 0.9))) for _ in range(2 * M)]
    ms = []
    for _ in range(NB + 1):
        y0 = rng.uniform(-9, 9)
        start = np.array([rng.uniform(-9, -8), y0, 0.0]); goal = np.array([-start[0], -y0, 0.0])
        obs = [o for o in field if min(np.linalg.norm(np.array(o[0]) - start[:2]), np.linalg.norm(np.array(o[0]) - goal[:2])) > o[1] + 1.0]
        assert len(obs) >= M
        ms.append(omodels.unicycle(r_init=start, r_final=goal, obstacles=obs[:M]))
    XU = [m.initialize_trajectory(K) for m in ms]
    Xs = [x + (0.05 * rng.normal(size=x.shape) if j else 0) for j, (x, _) in enumerate(XU)]
    mats = ofoh.OracleFOH(ms[0], K).calculate_discretization(Xs[0], XU[0][1], sigma)
    nbrs = []
    for j in range(1, NB + 1):
        a, _ = linearize_collision(2, d_min, Xs[0], Xs[j])
        nbrs.append({"a": a, "Y": Xs[j][:2] + 0.01 * rng.normal(size=(2, K)), "Lam": 0.1 * rng.normal(size=(2, K))})
    return ospb.Params(ms[0], K, mats, Xs[0], XU[0][1], sigma, 100.0, neighbors=nbrs, rho=rho, d_min=d_min)
