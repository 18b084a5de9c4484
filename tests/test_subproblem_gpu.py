"""GPU parity, stage 3: the batched interior-point kernel (scvx_solve_batched, through the C-ABI) against
the exact HiGHS oracle on identical parameters.

Gate (north_star): per-sub-problem optimal value within 1e-4 relative, constraint violation <= 1e-6.
The kernel is held to much tighter numbers here (1e-7 / 1e-8) so that regressions show early.
Minimisers of these LPs are not unique, so trajectories are never compared entry-wise;
the candidate (X, U, sigma) returned by the GPU is scored by the ORACLE's objective/violation evaluator.
"""
import numpy as np
import pytest
import torch

import helpers
from oracle import foh as ofoh, models as omodels, subproblem as ospb
from oracle.models import linearize_collision

pytestmark = pytest.mark.gpu

OBJ_RTOL = 1e-7      # north_star gate: 1e-4
VIOL_TOL = 1e-8      # north_star gate: 1e-6


def check_against_oracle(ws, pairs, obj_rtol=OBJ_RTOL):
    for i, (p, r) in enumerate(pairs):
        X, U, s = ws.X[i].cpu().numpy(), ws.U[i].cpu().numpy(), ws.sigma[i].item()
        e = ospb.evaluate(p, X, U, s)
        assert ws.status[i].item() == 0, (i, ws.status[i].item(), ws.iters[i].item())
        assert e["viol"] <= VIOL_TOL, (i, e["viol"])
        assert abs(e["obj"] - r["obj"]) <= obj_rtol * abs(r["obj"]), (i, e["obj"], r["obj"])
        # the objective the kernel reports is the oracle's evaluation of its own point
        assert ws.objective[i].item() == pytest.approx(e["obj"], rel=1e-12)
        # outputs are consistent: nu = defect of (X, U, sigma); s' = hinge value
        np.testing.assert_allclose(ws.nu[i].cpu().numpy(), e["nu"], rtol=0, atol=1e-12)
        if len(p.model.obstacles):
            np.testing.assert_allclose(ws.s_prime[i].cpu().numpy(), e["s_prime"], rtol=0, atol=1e-12)


def test_config1_sequence_unicycle_K50(cuda):
    """BASELINE config 1: shipped single agent, K=50; the first 6 sub-problems of the outer loop."""
    pairs = helpers.make_problem_sequence(omodels.unicycle(), 50, 6)
    assert pairs[0][1]["obj"] == pytest.approx(70506.902736, rel=1e-9)     # 7.050690e+03 at w_nu=1e3
    check_against_oracle(helpers.solve_batch_on_gpu([p for p, _ in pairs], cuda), pairs)


def test_config2_random_scenes_K100(cuda):
    """BASELINE config 2 scenes (K=100, M=8 random discs), two outer iterations each, one batch."""
    rng = np.random.default_rng(0)
    pairs = [pr for _ in range(4) for pr in helpers.make_problem_sequence(helpers.random_unicycle_scene(rng), 100, 2)]
    check_against_oracle(helpers.solve_batch_on_gpu([p for p, _ in pairs], cuda), pairs)


def test_single_integrator_socp(cuda):
    """||u_k||_2 <= v_max rows (single_integrator_model.py:103-104): oracle = Kelley cuts on the exact LP."""
    pairs = helpers.make_problem_sequence(omodels.single_integrator(), 40, 3)
    ws = helpers.solve_batch_on_gpu([p for p, _ in pairs], cuda)
    check_against_oracle(ws, pairs, obj_rtol=1e-6)
    assert (torch.linalg.norm(ws.U, dim=1) <= 1.0 + 1e-9).all()


def _admm_problem(kind, N, K, i, rng, sigma=20.0, d_min=0.5, rho=1.0):
    if kind == "unicycle":
        ang = np.linspace(0, 2 * np.pi, N, endpoint=False)
        ms = [omodels.unicycle(r_init=[8 * np.cos(a), 8 * np.sin(a), 0], r_final=[-8 * np.cos(a), -8 * np.sin(a), 0],
                               obstacles=[([0.0, 0.0], 1.0)]) for a in ang]
        d = 2
    else:
        pts = rng.normal(size=(N, 3)); pts = 4 * pts / np.linalg.norm(pts, axis=1, keepdims=True)
        ms = [omodels.single_integrator(r_init=q, r_final=-q, obstacles=[([0.0, 0.0, 0.0], 1.0)]) for q in pts]
        d = 3
    XU = [m.initialize_trajectory(K) for m in ms]
    Xs = [x + (0.05 * rng.normal(size=x.shape) if 0 < j else 0) for j, (x, _) in enumerate(XU)]
    mats = ofoh.OracleFOH(ms[i], K).calculate_discretization(Xs[i], XU[i][1], sigma)
    nbrs = []
    for j in range(N):
        if j == i:
            continue
        a, _ = linearize_collision(d, d_min, Xs[i], Xs[j])
        nbrs.append({"a": a, "Y": Xs[j][:d] + 0.01 * rng.normal(size=(d, K)), "Lam": 0.1 * rng.normal(size=(d, K))})
    return ospb.Params(ms[i], K, mats, Xs[i], XU[i][1], sigma, 100.0, neighbors=nbrs, rho=rho, d_min=d_min)


@pytest.mark.parametrize("kind,N,K", [("unicycle", 4, 30), ("unicycle", 16, 100), ("single_integrator", 5, 24)])
def test_admm_variant_qp(cuda, kind, N, K):
    """agent_solver.py:79-102: inter-agent rows + augmented Lagrangian.  HiGHS' QP solver does not finish on
    these, so optimality is certified with the exact LP solver: LB <= f_opt <= f(z_gpu) (Frank-Wolfe bracket) -- for the
    single-integrator SOCP variant too (Kelley cuts + bracket; round 1 compared it with the numpy twin of the kernel only)."""
    rng = np.random.default_rng(N * 100 + K)
    ps = [_admm_problem(kind, N, K, i, rng) for i in range(min(N, 3))]
    ws = helpers.solve_batch_on_gpu(ps, cuda)
    for i, p in enumerate(ps):
        X, U, s = ws.X[i].cpu().numpy(), ws.U[i].cpu().numpy(), ws.sigma[i].item()
        assert ws.status[i].item() == 0
        # both model kinds are certified by the exact LP solver, independently of the kernel's algorithm: the single
        # integrator's SOC rows are enforced by Kelley cuts inside ospb.solve (1e-9), the quadratic is linearised about the
        # candidate, and the LP optimum is a lower bound: LB <= f_opt <= f(z_gpu)
        f0, lb, viol, ok = ospb.qp_bracket(p, X, U, s)
        assert ok and viol <= VIOL_TOL
        assert 0 <= f0 - lb + 1e-9 * abs(f0) and (f0 - lb) <= 1e-6 * abs(f0), (f0, lb)
        const = sum(-(nb["Lam"] * nb["Y"]).sum() + 0.5 * p.rho * (nb["Y"] ** 2).sum() for nb in p.neighbors)
        assert ws.objective[i].item() + const == pytest.approx(f0, rel=1e-10)
        e = ospb.evaluate(p, X, U, s)
        np.testing.assert_allclose(ws.col_slack[i].cpu().numpy(), np.stack(e["S"]), rtol=0, atol=1e-12)


def _assert_bracketed(ws, ps, gap_rtol=1e-6):
    """Every candidate of the batch certified by the exact-LP bracket; the LPs run on host threads (HiGHS releases the GIL:
    four config-4-size brackets take the wall time of one, ~50 s)."""
    from concurrent.futures import ThreadPoolExecutor
    cands = [(p, ws.X[i].cpu().numpy(), ws.U[i].cpu().numpy(), ws.sigma[i].item()) for i, p in enumerate(ps)]
    assert (ws.status[:len(ps)] == 0).all(), (ws.status.tolist(), ws.iters.tolist())
    with ThreadPoolExecutor(len(ps)) as pool:
        out = list(pool.map(lambda c: ospb.qp_bracket(*c), cands))
    for i, (f0, lb, viol, ok) in enumerate(out):
        assert ok and viol <= VIOL_TOL, (i, viol)
        assert 0 <= f0 - lb + 1e-9 * abs(f0) and (f0 - lb) <= gap_rtol * abs(f0), (i, f0, lb)


def test_config4_size_single_integrator_admm(cuda):
    """BASELINE config 4 size: 256 single-integrator agents, K=100, all pairs -> 255 inter-agent rows per stage (25 500 hinge
    pairs per agent).  Four agents of one ADMM round, each certified by Kelley cuts + the exact-LP bracket."""
    rng = np.random.default_rng(4)
    ps = [_admm_problem("single_integrator", 256, 100, i, rng) for i in range(4)]
    assert all(len(p.neighbors) == 255 for p in ps)
    ws = helpers.solve_batch_on_gpu(ps, cuda)
    assert (torch.linalg.norm(ws.U, dim=1) <= 1.0 + 1e-9).all()
    _assert_bracketed(ws, ps)


def test_config5_size_unicycle_admm(cuda):
    """BASELINE config 5 size: K=200 nodes, M=32 discs, 16 neighbours per agent (48 hinge pairs per stage), unicycle QP."""
    rng = np.random.default_rng(55)
    ps = [helpers.config5_style_problem(rng) for _ in range(4)]
    assert ps[0].K == 200 and len(ps[0].model.obstacles) == 32 and len(ps[0].neighbors) == 16
    _assert_bracketed(helpers.solve_batch_on_gpu(ps, cuda), ps)


def test_edge_cases(cuda):
    """No obstacles; start inside an inflated obstacle (slack > 0 forced at the fixed node and nearby);
    trust region at its 1e-3 floor; the smallest horizon with a free stage (K=3)."""
    K = 30
    cases = []
    m0 = omodels.unicycle(obstacles=[])
    cases.append((m0, K, 100.0))
    m1 = omodels.unicycle(r_init=[-4.0, -4.0, 0.0], r_final=[8.0, 8.0, 0.0], obstacles=[([-5.0, -4.0], 3.0)])
    cases.append((m1, K, 100.0))
    cases.append((omodels.unicycle(), K, 1e-3))
    cases.append((omodels.unicycle(), 3, 100.0))
    for m, k, tr in cases:
        X, U = m.initialize_trajectory(k)
        mats = ofoh.OracleFOH(m, k).calculate_discretization(X, U, 5.0)
        p = ospb.Params(m, k, mats, X, U, 5.0, tr)
        r = ospb.solve(p)
        assert r["ok"]
        ws = helpers.solve_batch_on_gpu([p], cuda)
        check_against_oracle(ws, [(p, r)])
    # the slack case really has slack
    X, U = m1.initialize_trajectory(K)
    p = ospb.Params(m1, K, ofoh.OracleFOH(m1, K).calculate_discretization(X, U, 5.0), X, U, 5.0, 100.0)
    assert helpers.solve_batch_on_gpu([p], cuda).s_prime.sum().item() > 1e-2


def test_bad_args_and_workspace(cuda):
    from scvx_b200 import _device, _lib
    lib = _lib.load()
    assert lib.scvx_solve_batched(None, None) == -1
    assert lib.scvx_solve_workspace_bytes(0, 4, 50, 3, 0) == 4 * 50 * 8 * (28 + 6 * 3)      # row state + the hinge pairs' pending step
    assert lib.scvx_solve_workspace_bytes(1, 1, 10, 2, 3) == 10 * 8 * (31 + 6 * 5 + 1)
    assert lib.scvx_solve_workspace_bytes(9, 1, 10, 0, 0) == 0
    m = omodels.unicycle()
    X, U = m.initialize_trajectory(10)
    p = ospb.Params(m, 10, ofoh.OracleFOH(m, 10).calculate_discretization(X, U, 1.0), X, U, 1.0, 100.0)
    orig = _device.SubproblemWorkspace.__init__

    def small(self, *a, **k):
        orig(self, *a, **k)
        self.nbytes = 16
    _device.SubproblemWorkspace.__init__ = small
    try:
        with pytest.raises(_lib.ScvxError, match="workspace"):
            helpers.solve_batch_on_gpu([p], cuda)
    finally:
        _device.SubproblemWorkspace.__init__ = orig
    # K too large for the shared-memory-resident factor is refused loudly, not silently degraded
    big = torch.zeros
    with pytest.raises(_lib.ScvxError):
        ws = _device.SubproblemWorkspace(_lib.MODEL_UNICYCLE, 1, 400, 0, 0, cuda)
        z = lambda *s: big(s, dtype=torch.float64, device=cuda)   # noqa: E731
        _device.solve_subproblem(ws, (z(1, 9, 399), z(1, 6, 399), z(1, 6, 399), z(1, 3, 399), z(1, 3, 399)), z(1, 3, 400),
                                 z(1, 2, 400), z(1), z(1) + 1, z(1, 3), z(1, 3), z(1) - 9, z(1) + 9, z(1) + 1, z(1) + 1,
                                 None, None, 1e4, 1e6, 100.0)


def test_full_size_batch_properties(cuda):
    """BASELINE config 2 size: 1024 agents x K=100, M=8.  Size-independent properties:
    every agent optimal; the optimum is no worse than the (feasible) reference point; batched == solo
    (bit-identical: agents are independent blocks); agent-permutation equivariance; and a sample of agents
    is checked against the HiGHS oracle."""
    from scvx_b200.batch import AgentBatch, BatchedSCvx
    from scvx_b200.models.unicycle_model import UnicycleModel
    n, K = 1024, 100
    rng = np.random.default_rng(0)
    oms = [helpers.random_unicycle_scene(rng) for _ in range(n)]
    models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in oms]
    eng = BatchedSCvx(models, K, max_iter=2)
    b = eng.batch
    X, U = b.initial_trajectories()
    sig = torch.ones(n, dtype=torch.float64, device=cuda); tr = torch.full((n,), 100.0, dtype=torch.float64, device=cuda)
    act = torch.ones(n, dtype=torch.int32, device=cuda); met = torch.zeros((n, 6), dtype=torch.float64, device=cuda)
    X0, U0 = X.clone(), U.clone()
    eng.iterate(X, U, sig, tr, act, met)
    torch.cuda.synchronize()
    ws = eng.ws
    assert (ws.status == 0).all(), torch.bincount(ws.status)
    obj = ws.objective.cpu().numpy()
    # oracle on a sample, with the oracle's own FOH (tight) so that only stage 3 is being compared
    for i in (0, 17, 511, 1023):
        om = oms[i]
        Xi, Ui = om.initialize_trajectory(K)
        np.testing.assert_allclose(X0[i].cpu().numpy(), Xi, atol=1e-14)
        mats = tuple(m_[i].cpu().numpy() for m_ in eng.mats)
        p = ospb.Params(om, K, mats, X0[i].cpu().numpy(), Ui, 1.0, 100.0)
        r = ospb.solve(p)
        assert abs(obj[i] - r["obj"]) <= OBJ_RTOL * abs(r["obj"])
        ref = ospb.evaluate(p, p.X_ref, p.U_ref, 1.0)
        assert obj[i] <= ref["obj"] * (1 + 1e-12)
    # batched == solo for one agent
    solo = BatchedSCvx([models[17]], K, max_iter=1)
    Xs, Us = solo.batch.initial_trajectories()
    solo.iterate(Xs, Us, torch.ones(1, dtype=torch.float64, device=cuda), torch.full((1,), 100.0, dtype=torch.float64, device=cuda),
                 torch.ones(1, dtype=torch.int32, device=cuda), torch.zeros((1, 6), dtype=torch.float64, device=cuda))
    assert torch.equal(solo.ws.X[0], ws.X[17]) and torch.equal(solo.ws.sigma[0], ws.sigma[17])
    # metrics written by the bookkeeping kernel agree with the solver outputs
    np.testing.assert_allclose(met[:, 5].cpu().numpy(), ws.sigma.cpu().numpy())
    assert torch.equal(X, ws.X)          # nobody converged in iteration 0: new iterate accepted everywhere


def test_accuracy_distribution_over_64_scenes(cuda):
    """Objective error against the exact LP over 64 seeded scenes x 2 outer iterations (128 sub-problems): the
    north_star gate is 1e-4 relative; the kernel must hold 1e-7 on EVERY one of them."""
    rng = np.random.default_rng(2026)
    pairs = [pr for _ in range(64) for pr in helpers.make_problem_sequence(helpers.random_unicycle_scene(rng), 40, 2)]
    ws = helpers.solve_batch_on_gpu([p for p, _ in pairs], cuda)
    assert (ws.status == 0).all()
    rel = []
    for i, (p, r) in enumerate(pairs):
        e = ospb.evaluate(p, ws.X[i].cpu().numpy(), ws.U[i].cpu().numpy(), ws.sigma[i].item())
        assert e["viol"] <= VIOL_TOL
        rel.append((e["obj"] - r["obj"]) / abs(r["obj"]))
    rel = np.array(rel)
    assert np.abs(rel).max() <= OBJ_RTOL, (np.abs(rel).max(), int(np.argmax(np.abs(rel))))
    assert ws.iters.max().item() <= 45


def test_long_horizons_K200_and_K220(cuda):
    """BASELINE config 5 horizon (K=200: factor + Jacobians still fit one block's 227 KB) and K=220, where the FOH
    Jacobians move to the global workspace (same code path otherwise; 256 threads per agent, strided over stages)."""
    from scvx_b200 import _lib
    lib = _lib.load()
    assert lib.scvx_solve_workspace_bytes(0, 1, 200, 4, 0) == 200 * 8 * (28 + 6 * 4)
    assert lib.scvx_solve_workspace_bytes(0, 1, 220, 4, 0) == 220 * 8 * (28 + 6 * 4 + 27)     # + K*NJ for the Jacobians
    rng = np.random.default_rng(5)
    om = helpers.random_unicycle_scene(rng, 4)
    for K in (200, 220):
        pairs = helpers.make_problem_sequence(om, K, 2)
        check_against_oracle(helpers.solve_batch_on_gpu([p for p, _ in pairs], cuda), pairs)
