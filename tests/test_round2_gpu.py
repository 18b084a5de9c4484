"""GPU tests of the round-2 additions: CUDA-graph replay of the lane pipeline, the adaptive barrier start, hinge groups and
thread-block clusters of the sub-problem kernel, and the fused ADMM-round kernel with the device neighbour selection."""
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import models as omodels, subproblem as ospb

pytestmark = pytest.mark.gpu


def _scene_models(n, seed, M=4):
    from scvx_b200.models.unicycle_model import UnicycleModel
    rng = np.random.default_rng(seed)
    oms = [helpers.random_unicycle_scene(rng, M) for _ in range(n)]
    return oms, [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in oms]


def test_graph_replay_is_identical_to_eager_lanes(cuda):
    """One CUDA graph per lane (replayed on the lane's stream) and one graph over all lanes launch exactly what the eager lane
    pipeline launches: iterates, sigma and metrics are bit-identical."""
    from scvx_b200.batch import PipelinedSCvx
    _, models = _scene_models(10, 5)
    K, n_it = 30, 6
    ref = PipelinedSCvx(models, K, n_lanes=3, max_iter=n_it).start()
    ref.run(n_it)
    want = ref.gather()
    for mode in ("lane", "all"):
        P = PipelinedSCvx(models, K, n_lanes=3, max_iter=n_it).start()
        P.run(2)
        if mode == "lane":
            P.build_lane_graphs(); P.run_lane_graphs(n_it - 2)
        else:
            P.build_graph(steps_per_graph=2); P.run_graph(n_it - 2)
        got = P.gather()
        torch.cuda.synchronize()
        for key in ("X", "U", "sigma", "tr_radius", "metrics"):
            assert torch.equal(got[key], want[key]), (mode, key)
        assert P.launches == ref.launches


def test_host_lane_graphs_are_identical_to_eager_host_steps(cuda):
    """`build_lane_host_graphs` captures, per lane, the H2D copies + the step's kernels + the D2H copies; replaying them gives the
    same host arrays as the eager `run_host`, bit for bit (adaptive barrier start and retry pass included)."""
    from scvx_b200.batch import PipelinedSCvx
    _, models = _scene_models(10, 6)
    K, n_it = 30, 6

    def fresh():
        P = PipelinedSCvx(models, K, n_lanes=3, max_iter=n_it, adaptive_mu0=True).start()
        host = P.make_host_buffers()
        host["X"].copy_(torch.cat([st[0] for st in P.state]).cpu()); host["U"].copy_(torch.cat([st[1] for st in P.state]).cpu())
        host["sigma"].fill_(1.0); host["tr"].fill_(100.0)
        return P, host

    P, want = fresh()
    P.run_host(want, n_it)
    Q, got = fresh()
    Q.run_host(got, 2)
    Q.build_lane_host_graphs(got)
    Q.run_host_lane_graphs(got, n_it - 2)
    for key in ("X", "U", "sigma", "tr", "active", "metrics"):
        assert torch.equal(got[key], want[key]), key
    assert Q.launches == P.launches


def test_adaptive_barrier_start_keeps_every_subproblem_optimal(cuda):
    """adaptive_mu0: agents whose previous solve was short start the next one at mu = 0.1.  Every sub-problem of 6 outer
    iterations still matches the exact LP on the loop's own parameters to 1e-7 (the start changes WHICH minimiser of a
    degenerate LP is returned, never its value), with fewer interior-point iterations in total."""
    from scvx_b200.batch import BatchedSCvx
    oms, models = _scene_models(3, 21, M=6)
    K = 40
    totals = {}
    for adaptive in (False, True):
        eng = BatchedSCvx(models, K, max_iter=6, adaptive_mu0=adaptive)
        b = eng.batch
        X, U = b.initial_trajectories()
        n = b.n
        sig = torch.ones(n, dtype=torch.float64, device=cuda); tr = torch.full((n,), 100.0, dtype=torch.float64, device=cuda)
        act = torch.ones(n, dtype=torch.int32, device=cuda); met = torch.zeros((n, 6), dtype=torch.float64, device=cuda)
        total = 0
        for _ in range(6):
            Xr, Ur, sr, trr = X.cpu().numpy().copy(), U.cpu().numpy().copy(), sig.cpu().numpy().copy(), tr.cpu().numpy().copy()
            eng.iterate(X, U, sig, tr, act, met)
            torch.cuda.synchronize()
            assert (eng.ws.status == 0).all()
            total += int(eng.ws.iters.sum().item())
            if adaptive:
                for i in range(n):
                    mats = tuple(m[i].cpu().numpy() for m in eng.mats)
                    p = ospb.Params(oms[i], K, mats, Xr[i], Ur[i], float(sr[i]), float(trr[i]))
                    e = ospb.evaluate(p, eng.ws.X[i].cpu().numpy(), eng.ws.U[i].cpu().numpy(), float(eng.ws.sigma[i].item()))
                    r = ospb.solve(p)
                    assert e["viol"] <= 1e-8 and abs(e["obj"] - r["obj"]) <= 1e-7 * abs(r["obj"]), (i, e["obj"], r["obj"])
        totals[adaptive] = total
    assert totals[True] <= totals[False]


def test_fixed_size_kernels_agree_with_the_generic_one(cuda):
    """`ipm_kernel` instantiated with K = 100 (and, for the plain sub-problem, M = 8) as compile-time constants -- the launch picks
    these for the BASELINE sizes -- solves the same problems as the generic kernel (SCVX_NO_FIXED=1): same statuses, iteration
    counts within one, optimal values equal to 1e-9 (the compiler contracts other multiply-add pairs: round-off, not bitwise)."""
    import test_subproblem_gpu as T
    rng = np.random.default_rng(5)
    oms = [helpers.random_unicycle_scene(rng, 8) for _ in range(6)]
    plain = []
    for om in oms:
        plain += [p for p, _ in helpers.make_problem_sequence(om, 100, 2)]
    coupled = [T._admm_problem("unicycle", 5, 100, i, rng) for i in range(3)]           # K = 100 with inter-agent rows and quadratic terms
    crowd = [T._admm_problem("single_integrator", 81, 100, i, rng) for i in range(2)]    # two hinge groups, cluster launch
    for ps in (plain, coupled, crowd):
        got = helpers.solve_batch_on_gpu(ps, cuda)
        g = (got.objective.cpu().numpy().copy(), got.status.cpu().numpy().copy(), got.iters.cpu().numpy().copy())
        os.environ["SCVX_NO_FIXED"] = "1"
        try:
            ref = helpers.solve_batch_on_gpu(ps, cuda)
        finally:
            os.environ.pop("SCVX_NO_FIXED", None)
        assert (g[1] == 0).all() and (ref.status == 0).all()
        np.testing.assert_allclose(g[0], ref.objective.cpu().numpy(), rtol=1e-9)
        assert np.abs(g[2] - ref.iters.cpu().numpy()).max() <= 1


def test_retry_pass_solves_only_the_failed_agents(cuda):
    """scvx_solve_args.retry_failed: a second launch of the same arguments re-solves, from the cold start, exactly the agents whose
    status is not optimal; every other block returns at once and leaves all of its outputs (iters included) alone.  A solve from a
    small barrier start gives up after 40 iterations instead of running to the cap."""
    from scvx_b200 import _device
    from scvx_b200.batch import BatchedSCvx
    _, models = _scene_models(5, 33, M=6)
    K = 40
    eng = BatchedSCvx(models, K)
    b, ws = eng.batch, eng.ws
    X, U = b.initial_trajectories()
    n = b.n
    sig = torch.ones(n, dtype=torch.float64, device=cuda); tr = torch.full((n,), 100.0, dtype=torch.float64, device=cuda)
    _device.foh(b.model_id, X, U, sig, 0, out=eng.mats)
    _device.linearize_obstacles(b.model_id, X, b.obs_c, b.obs_clear, out=(eng.obs_a, eng.obs_b))

    def solve(**kw):
        _device.solve_subproblem(ws, eng.mats, X, U, sig, tr, b.x_init, b.x_final, b.pos_lo, b.pos_hi, b.v_max, b.w_max,
                                 eng.obs_a, eng.obs_b, eng.weight_nu, eng.weight_slack, eng.weight_sigma, **kw)
        torch.cuda.synchronize()
        return {k: getattr(ws, k).clone() for k in ("X", "U", "nu", "sigma", "objective", "status", "iters")}

    cold = solve()
    assert (cold["status"] == 0).all()
    cut = solve(max_iter=2)                       # every agent stops at the cap: status MAXITER, outputs of a half-done solve
    assert (cut["status"] == 1).all() and (cut["iters"] == 2).all()
    ws.status[0] = 0                              # pretend agent 0 had converged
    again = solve(retry_failed=True)
    for key in cold:
        assert torch.equal(again[key][1:], cold[key][1:]), key          # re-solved from the cold start: the cold solve, bit for bit
        if key != "status":
            assert torch.equal(again[key][0], cut[key][0]), key         # agent 0: untouched
    # small start: either converged, or gave up at 40 iterations (never the full cap of 80)
    small = solve(mu0=torch.full((n,), 1e-3, dtype=torch.float64, device=cuda))
    assert ((small["status"] == 0) | (small["iters"] <= 40)).all()
    np.testing.assert_allclose(small["objective"].cpu().numpy(), cold["objective"].cpu().numpy(), rtol=1e-7)


def _si_crowd(n_nbr, K, n_problems, seed):
    """Single-integrator ADMM sub-problems with n_nbr inter-agent rows per stage (enough for hinge groups / clusters)."""
    import test_subproblem_gpu as T
    rng = np.random.default_rng(seed)
    return [T._admm_problem("single_integrator", n_nbr + 1, K, i, rng) for i in range(n_problems)]


def test_hinge_groups_and_clusters_solve_the_same_problems(cuda):
    """The hinge rows of a stage split over 1, 2 or 4 thread groups of a block and over clusters of 1, 2 or 4 blocks: same
    optimal values (the partial sums are added in a different order: agreement to round-off, not bitwise), same statuses, and
    one instance certified against the exact LP bracket."""
    ps = _si_crowd(80, 24, 3, 77)
    vals = {}
    try:
        for g, c in ((1, 1), (2, 1), (4, 1), (2, 2), (2, 4)):
            os.environ["SCVX_HINGE_GROUPS"], os.environ["SCVX_CLUSTER"] = str(g), str(c)
            ws = helpers.solve_batch_on_gpu(ps, cuda)
            assert (ws.status == 0).all(), (g, c, ws.status.tolist())
            vals[(g, c)] = (ws.objective.cpu().numpy().copy(), ws.iters.cpu().numpy().copy(), ws.X.cpu().numpy().copy())
    finally:
        os.environ.pop("SCVX_HINGE_GROUPS", None); os.environ.pop("SCVX_CLUSTER", None)
    base = vals[(1, 1)]
    for key, (obj, its, X) in vals.items():
        np.testing.assert_allclose(obj, base[0], rtol=1e-9, err_msg=str(key))
        assert np.abs(its - base[1]).max() <= 1, (key, its, base[1])
        np.testing.assert_allclose(X, base[2], atol=1e-6, err_msg=str(key))          # the QP is strictly convex in the positions
    ws = helpers.solve_batch_on_gpu(ps[:1], cuda)          # default choice for one agent: a cluster of four blocks
    f0, lb, viol, ok = ospb.qp_bracket(ps[0], ws.X[0].cpu().numpy(), ws.U[0].cpu().numpy(), ws.sigma[0].item())
    assert ok and viol <= 1e-8 and 0 <= f0 - lb + 1e-9 * abs(f0) and f0 - lb <= 1e-6 * abs(f0)


@pytest.mark.parametrize("d,n_x", [(2, 3), (3, 3)])
def test_admm_round_prep_against_numpy(cuda, d, n_x):
    """scvx_admm_round_prep: normals of linearize_collision (multi_agent_model.py:61-79), right-hand sides d_min + a.Y_j and the
    collapsed augmented-Lagrangian terms (agent_solver.py:79-95), all pairs / masked / indexed, against a direct numpy
    evaluation; known answer of the reference's own test (test_multi_agent_model.py:42-59): a = [0, 1], b = d_min."""
    from scvx_b200 import _device, _lib
    mid = _lib.MODEL_UNICYCLE if d == 2 else _lib.MODEL_SINGLE_INTEGRATOR
    rng = np.random.default_rng(3 + d)
    N, K, i0, nl, d_min, rho = 7, 13, 2, 3, 0.5, 1.7
    X = rng.normal(size=(N, n_x, K)); Y = rng.normal(size=(N, d, K)); L = rng.normal(size=(N, d, K))
    Xo = X[i0:i0 + nl] + 0.1 * rng.normal(size=(nl, n_x, K))
    to = lambda a, dt=None: torch.as_tensor(np.ascontiguousarray(a)).to(cuda)      # noqa: E731

    def numpy_ref(slots):          # slots[i][q] = neighbour id or -1
        nq = len(slots[0])
        a = np.zeros((nl, nq, d, K)); b = np.zeros((nl, nq, K)); lin = np.zeros((nl, d, K)); nact = np.zeros(nl); const = np.zeros(nl)
        for i in range(nl):
            for q, j in enumerate(slots[i]):
                if j < 0 or j == i0 + i:
                    continue
                diff = Xo[i, :d] - X[j, :d]
                a[i, q] = diff / (np.linalg.norm(diff, axis=0) + 1e-6)
                b[i, q] = d_min + (a[i, q] * Y[j]).sum(axis=0)
                lin[i] += L[j] - rho * Y[j]; nact[i] += 1
                const[i] += 0.5 * rho * (Y[j] ** 2).sum() - (L[j] * Y[j]).sum()
        return a, b, lin, rho * nact, const

    def check(tab, ref, mask_ref):
        a, b, lin, quad, const = ref
        np.testing.assert_allclose(tab.col_a.cpu().numpy(), a, atol=1e-15); np.testing.assert_allclose(tab.col_b.cpu().numpy(), b, atol=1e-13)
        np.testing.assert_allclose(tab.lin_p.cpu().numpy(), lin, atol=1e-12); np.testing.assert_allclose(tab.quad_rho.cpu().numpy(), quad)
        np.testing.assert_allclose(tab.aug_const.cpu().numpy(), const, rtol=1e-12, atol=1e-12)
        assert tab.mask.cpu().numpy().tolist() == mask_ref

    # all pairs
    tab = _device.AdmmRoundTables(mid, nl, N, K, cuda)
    _device.admm_round_prep(mid, to(Xo), to(X), to(Y), to(L), d_min, rho, i0, tab)
    slots = [list(range(N)) for _ in range(nl)]
    check(tab, numpy_ref(slots), [[0 if j == i0 + i else 1 for j in range(N)] for i in range(nl)])
    # masked (neighbour culling of the all-pairs tables)
    mask = (rng.uniform(size=(nl, N)) < 0.6).astype(np.uint8)
    _device.admm_round_prep(mid, to(Xo), to(X), to(Y), to(L), d_min, rho, i0, tab, mask_in=to(mask))
    slots = [[j if mask[i, j] else -1 for j in range(N)] for i in range(nl)]
    check(tab, numpy_ref(slots), [[int(mask[i, j] and j != i0 + i) for j in range(N)] for i in range(nl)])
    # indexed slots with empty entries
    idx = np.array([[4, -1, 0], [1, 6, -1], [-1, -1, 3]], dtype=np.int32)
    tab3 = _device.AdmmRoundTables(mid, nl, 3, K, cuda)
    _device.admm_round_prep(mid, to(Xo), to(X), to(Y), to(L), d_min, rho, i0, tab3, nbr_idx=to(idx))
    check(tab3, numpy_ref(idx.tolist()), [[int(j >= 0 and j != i0 + i) for j in row] for i, row in enumerate(idx.tolist())])
    if d == 2:        # SCvx/multi_agent_tests/test_multi_agent_model.py:42-59: agent above its neighbour -> a = [0, 1], b = d_min
        Xa = np.zeros((2, 3, 4)); Xa[0, 1] = 1.0
        t2 = _device.AdmmRoundTables(mid, 1, 2, 4, cuda)
        _device.admm_round_prep(mid, to(Xa[:1]), to(Xa), to(Xa[:, :2].copy()), to(np.zeros((2, 2, 4))), 0.5, 1.0, 0, t2)
        np.testing.assert_allclose(t2.col_a[0, 1].cpu().numpy(), np.tile([[0.0], [1.0]], (1, 4)), atol=1e-6)
        np.testing.assert_allclose(t2.col_b[0, 1].cpu().numpy(), 0.5, atol=1e-12)


def test_device_neighbour_selection(cuda):
    """scvx_knn_select / scvx_radius_mask against torch.topk / a comparison on the same table."""
    from scvx_b200 import _device
    rng = np.random.default_rng(8)
    nl, N, i0, k = 5, 300, 40, 16
    d2 = rng.uniform(0.0, 4.0, size=(nl, N))
    t = torch.as_tensor(d2).to(cuda)
    want = t.clone()
    want[torch.arange(nl), torch.arange(i0, i0 + nl)] = float("inf")
    val, idx = torch.topk(want, k, dim=1, largest=False)
    got = _device.knn_select(t.clone(), i0, k)
    assert torch.equal(got.long(), idx)
    got_r = _device.knn_select(t.clone(), i0, k, radius=0.5)
    assert torch.equal(got_r.long(), torch.where(val <= 0.25, idx, torch.full_like(idx, -1)))
    few = _device.knn_select(t[:3, :4].contiguous().clone(), 0, 6)         # fewer candidates than slots: the rest are -1
    assert (few[:, 3:] == -1).all() and (few[:, :3] >= 0).all()
    m = _device.radius_mask(t, 1.0)
    assert torch.equal(m.bool(), t <= 1.0)
