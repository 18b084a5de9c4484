"""One launch of EVERY kernel of libscvx_b200.so at a representative size, for `ncu --set full` (north_star: "ncu-backed
roofline fractions for every kernel").  The plain run prints one JSON line per kernel family with its CUDA-event time and
the ALGORITHMIC bytes / fp64 flops of the launch (DESIGN.md section 4 states the per-unit figures); under ncu the same
launches are captured (tools/ncu_all_report.py turns the report into profiles/*_all_kernels_ncu.md).

    python tools/prof_all_kernels.py > gpurun_out/all_plain.log
    ncu --metrics $(python tools/prof_all_kernels.py --metrics) --clock-control none --profile-from-start off \
        -k "regex:$(python tools/prof_all_kernels.py --regex)" --csv --log-file gpurun_out/all_metrics.csv python tools/prof_all_kernels.py --once
(a targeted metric list, a few replay passes per launch: `--set full` over all ~40 launches costs 6+ GPU-minutes and a 90 MB report)
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)

KERNEL_REGEX = ("foh_rk4|integrate_|linearize_|ipm_kernel|outer_update|order_by_iters|consensus_kernel|lti_qp|sbar_|"
                "slab_normals|warm_start|min_pair|min_obstacle|intersample|clearance_samples|cross_min|admm_prep|knn_select|radius_mask|mu0_from_iters|retry_list")
METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,"
           "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic")
if "--regex" in sys.argv:
    print(KERNEL_REGEX); sys.exit(0)
if "--metrics" in sys.argv:
    print(METRICS); sys.exit(0)

import numpy as np, torch
from scvx_b200 import _device, _lib
from scvx_b200.batch import BatchedSCvx, BatchedADMM
from scvx_b200.models.unicycle_model import UnicycleModel
from scvx_b200.models.single_integrator_model import SingleIntegratorModel
from scvx_b200.Distributed_opt import _engine
import scvx_b200.Distributed_opt.ADMM_decentralized as M2
import scvx_b200.Distributed_opt.dist_scvx_3d as M3

ONCE = "--once" in sys.argv          # under ncu: no warm-up repeat, one launch per kernel
dev = torch.device("cuda")
F64 = torch.float64
rng = np.random.default_rng(0)


def timed(name, fn, alg_bytes=None, alg_flops=None, note=""):
    if not ONCE:
        fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if ONCE:
        torch.cuda.synchronize(); torch.cuda.profiler.start()      # ncu --profile-from-start off: only these launches are captured
    a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    if ONCE:
        torch.cuda.profiler.stop()
    ms = a.elapsed_time(b)
    rec = {"kernel": name, "ms": ms, "note": note}
    if alg_bytes is not None:
        rec["alg_bytes"] = alg_bytes; rec["alg_GBps"] = alg_bytes / ms / 1e6
    if alg_flops is not None:
        rec["alg_flops"] = alg_flops; rec["alg_TFLOPs"] = alg_flops / ms / 1e9
    print(json.dumps(rec), flush=True)
    return r


# ---------------------------------------------------------------- config 2: 1024 unicycle agents x K = 100, M = 8
n, K, M = 1024, 100, 8


def scene(M):
    """start on the left edge band, goal mirrored, M random discs clear of both."""
    y0 = rng.uniform(-9, 9)
    start = np.array([rng.uniform(-9, -8), y0, 0.0]); goal = np.array([-start[0], -y0, 0.0])
    obs = []
    while len(obs) < M:
        c = rng.uniform(-7, 7, 2); r = rng.uniform(0.5, 2.0)
        if min(np.linalg.norm(c - start[:2]), np.linalg.norm(c - goal[:2])) > r + 1.0:
            obs.append((list(c), float(r)))
    return start, goal, obs


scenes = [scene(M) for _ in range(n)]
models = [UnicycleModel(r_init=s0, r_final=s1, obstacles=ob) for s0, s1, ob in scenes]
eng = BatchedSCvx(models, K, max_iter=4, adaptive_mu0=True)      # with the warm barrier start: mu0 map, retry list, retry pass
b = eng.batch
X, U = b.initial_trajectories()
sig = torch.ones(n, dtype=F64, device=dev); tr = torch.full((n,), 100.0, dtype=F64, device=dev)
act = torch.ones(n, dtype=torch.int32, device=dev); met = torch.zeros((4, n, 6), dtype=F64, device=dev)
eng.iterate(X, U, sig, tr, act, met[0]); eng.iterate(X, U, sig, tr, act, met[1]); torch.cuda.synchronize()   # leave the idle start
Km1 = K - 1
timed("foh_rk4_kernel<Unicycle>", lambda: _device.foh(b.model_id, X, U, sig, 0, out=eng.mats),
      alg_bytes=n * 8 * (5 * K + 1 + 27 * Km1), note=f"{n} agents x {Km1} intervals; flops = n_sub x 1.34 kflop per interval (n_sub picked on the device)")
timed("linearize_obstacles_kernel<2>", lambda: _device.linearize_obstacles(b.model_id, X, b.obs_c, b.obs_clear, out=(eng.obs_a, eng.obs_b)),
      alg_bytes=n * 8 * (2 * K + M * 3 + M * 3 * K), note=f"{n} agents x {M} discs x {K} nodes")
timed("ipm_kernel<Unicycle> + order + outer_update (one SCvx iteration)", lambda: eng.iterate(X, U, sig, tr, act, met[2]),
      note="the bench step; see bench.py for the roofline of ipm_kernel")
Xp = timed("integrate_piecewise_kernel<Unicycle>", lambda: _device.integrate_piecewise(b.model_id, X, U, sig),
           alg_bytes=n * 8 * (5 * K + 1 + 3 * K))
timed("integrate_full_kernel<Unicycle>", lambda: _device.integrate_full(b.model_id, X[:, :, 0].contiguous(), U, sig),
      alg_bytes=n * 8 * (3 + 2 * K + 1 + 3 * K), note="sequential in k per agent (one thread per agent)")

# ---------------------------------------------------------------- single integrator: 256 agents x K = 100 (SOCP)
ns = 256
i = np.arange(ns) + 0.5; phi = np.arccos(1 - 2 * i / ns); th = np.pi * (1 + 5 ** 0.5) * i
pts = 8.0 * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
si = [SingleIntegratorModel(r_init=p, r_final=-p, obstacles=[([0.0, 0.0, 0.0], 1.0)]) for p in pts]
es = BatchedSCvx(si, K, max_iter=3)
Xs, Us = es.batch.initial_trajectories()
sg = torch.ones(ns, dtype=F64, device=dev); trs = torch.full((ns,), 100.0, dtype=F64, device=dev)
acs = torch.ones(ns, dtype=torch.int32, device=dev); mes = torch.zeros((3, ns, 6), dtype=F64, device=dev)
es.iterate(Xs, Us, sg, trs, acs, mes[0]); torch.cuda.synchronize()
timed("foh_rk4_kernel<SI> + linearize<3> + ipm_kernel<SingleIntegrator> (one SCvx iteration)", lambda: es.iterate(Xs, Us, sg, trs, acs, mes[1]),
      note=f"{ns} single-integrator agents x K={K}, second-order-cone velocity rows")

# ---------------------------------------------------------------- config 3: 16 unicycle agents, all pairs, one ADMM round
N3 = 16
ang = np.linspace(0, 2 * np.pi, N3, endpoint=False)
ms3 = [UnicycleModel(r_init=np.array([8 * np.cos(a), 8 * np.sin(a), 0.0]), r_final=np.array([-8 * np.cos(a), -8 * np.sin(a), 0.0]),
                     obstacles=[([0.0, 0.0], 1.0)]) for a in ang]
XU = [m.initialize_trajectory(np.zeros((3, K)), np.zeros((2, K))) for m in ms3]
X0 = torch.as_tensor(np.stack([x for x, _ in XU])).to(dev); U0 = torch.as_tensor(np.stack([u for _, u in XU])).to(dev)
adm = BatchedADMM(ms3, 0.5, K, max_iter=1)
timed("ADMM round: linearize_collision + ipm_kernel (n_nbr = 16) + consensus_kernel", lambda: adm.solve(X0, U0, 20.0),
      note=f"{N3} agents, all pairs, K={K}: latency-bound (16 blocks)")

# ---------------------------------------------------------------- config 4: 256 single-integrator agents, all pairs, one ADMM round
XU4 = [m.initialize_trajectory(np.zeros((3, K)), np.zeros((3, K))) for m in si]
X4 = torch.as_tensor(np.stack([x for x, _ in XU4])).to(dev); U4 = torch.as_tensor(np.stack([u for _, u in XU4])).to(dev)
adm4 = BatchedADMM(si, 0.5, K, max_iter=1, si_variant=True)
timed("ADMM round, config 4: admm_prep_kernel<3> + ipm_kernel<SingleIntegrator, G=2> (n_nbr = 256) + consensus_kernel", lambda: adm4.solve(X4, U4, 20.0),
      alg_bytes=ns * 30 * 8 * 255 * K * 36, note=f"{ns} agents, all pairs, K={K}: 255 hinge pairs per stage, two hinge groups per block; "
      "alg_bytes assumes 36 interior-point iterations")
tab4 = _device.AdmmRoundTables(_lib.MODEL_SINGLE_INTEGRATOR, ns, ns, K, dev)
Y4 = X4[:, :3].clone(); L4 = torch.zeros_like(Y4)
timed("admm_prep_kernel<3>", lambda: _device.admm_round_prep(_lib.MODEL_SINGLE_INTEGRATOR, X4, X4, Y4, L4, 0.5, 1.0, 0, tab4),
      alg_bytes=ns * ns * K * 8 * (3 + 3 + 3 + 4), note=f"{ns} x {ns} slots x K={K}: reads X, Y, Lambda of every neighbour, writes a (3) and b")

# ---------------------------------------------------------------- neighbour tables at config-5 scale (2048 agents x K = 200)
N5, K5 = 2048, 200
X5 = torch.as_tensor(rng.normal(size=(N5, 3, K5)) * 5.0, device=dev)
d2 = timed("cross_min_dist2_kernel", lambda: _device.cross_min_dist2(_lib.MODEL_UNICYCLE, X5, X5), alg_bytes=N5 * 2 * K5 * 8 * 2 + N5 * N5 * 8,
           alg_flops=N5 * N5 * K5 * 5.0, note=f"{N5} x {N5} pairs x K={K5}; the N^2 K re-reads are served from L2 / shared memory")
idx = timed("knn_select_kernel", lambda: _device.knn_select(d2.clone(), 0, 16), alg_bytes=N5 * N5 * 8 + N5 * 16 * 4,
            note=f"{N5} rows x {N5} candidates, 16 nearest (one warp per row, 16 rounds of warp arg-min over the row)")
timed("radius_mask_kernel", lambda: _device.radius_mask(d2, 1.0), alg_bytes=N5 * N5 * 9, note=f"{N5} x {N5} table")
timed("linearize_collision_indexed_kernel", lambda: _device.linearize_collision_indexed(_lib.MODEL_UNICYCLE, X5, X5, idx, 0.5),
      alg_bytes=N5 * 16 * K5 * 8 * (2 + 3), note=f"{N5} agents x 16 neighbours x K={K5}")
timed("min_pair_distance_kernel", lambda: _device.min_inter_agent_distance(X5), alg_bytes=N5 * 3 * K5 * 8 + N5 * N5 * 8,
      alg_flops=N5 * (N5 - 1) / 2 * K5 * 9.0, note=f"{N5 * (N5 - 1) // 2} pairs x K={K5}")
Co = torch.as_tensor(rng.uniform(-5, 5, (64, 3)), device=dev); Ro = torch.as_tensor(rng.uniform(0.2, 1, 64), device=dev)
timed("min_obstacle_distance_kernel", lambda: _device.min_agent_obstacle_distance(X5, Co, Ro, 0.5), alg_bytes=N5 * 3 * K5 * 8 + N5 * 64 * 8,
      note=f"{N5} agents x 64 obstacles x K={K5}")
Pd = torch.as_tensor(rng.normal(size=(256, 3, K)) * 5.0, device=dev)
timed("slab_normals_kernel", lambda: _device.slab_normals(_lib.MODEL_UNICYCLE, Pd, Pd, Pd + 0.1, torch.full((256,), 0.3, dtype=F64, device=dev)),
      alg_bytes=256 * 256 * K * 8 * 3, note=f"256 x 256 pairs x K={K}")

# ---------------------------------------------------------------- warm starts: 8192 agents x K = 200 x M = 32
nw, Kw, Mw = 8192, 200, 32
p0 = np.zeros((nw, 3)); p1 = np.zeros((nw, 3))
p0[:, 0] = rng.uniform(-9.5, -9, nw); p0[:, 1] = rng.uniform(-9, 9, nw); p1[:, 0] = -p0[:, 0]; p1[:, 1] = -p0[:, 1]
Cw = rng.uniform(-7, 7, (nw, Mw, 2)); Rw = rng.uniform(0.2, 0.8, (nw, Mw))
tw = [torch.as_tensor(a, device=dev) for a in (p0, p1, Cw, Rw)]
timed("warm_start_kernel<2>", lambda: _device.warm_start(_lib.MODEL_UNICYCLE, tw[0], tw[1], tw[2], tw[3], 0.3, Kw),
      alg_bytes=nw * 8 * (6 + 3 * Mw + 5 * Kw), note=f"{nw} agents x K={Kw} x M={Mw}")

# ---------------------------------------------------------------- inter-sample clearance: 256 agents x K = 50 x 3 discs
ni, Ki = 64, 50
eng_i = BatchedSCvx(models[:ni], Ki, max_iter=10)       # run the outer loop far enough that the trajectories really move (sigma >> 0)
out_i = eng_i.solve(early_exit=False)
Xi, Ui, si_ = out_i["X"], out_i["U"], out_i["sigma"]
oc = torch.as_tensor(np.array([[c for c, _ in ob][:3] for _, _, ob in scenes[:ni]]), device=dev)
orad = torch.as_tensor(np.array([[r + 0.5 for _, r in ob][:3] for _, _, ob in scenes[:ni]]), device=dev)
timed("intersample_kernel<Unicycle>", lambda: _device.intersample(_lib.MODEL_UNICYCLE, Xi, Ui, si_, oc, orad),
      note=f"{ni} agents x {Ki - 1} segments x 3 discs, 100 grid samples each (FP64 pipe: >= 200 RK4 flows per item)")
timed("clearance_samples_kernel<Unicycle>", lambda: _device.clearance_samples(_lib.MODEL_UNICYCLE, Xi, Ui, si_, oc[:, 0].contiguous(), orad[:, 0].contiguous()),
      note=f"{ni} agents x {Ki - 1} segments x 50 samples")

# ---------------------------------------------------------------- Distributed_opt rows
X2 = M2.x_initial(M2.x_ini, M2.x_des)
names = M2.robots_name
Xd2 = torch.as_tensor(np.stack([X2[r] for r in names]), device=dev)
xdes2 = np.stack([M2.x_des[r][:M2.n] for r in names])
R2 = 256
Xb2 = Xd2[torch.arange(R2, device=dev) % Xd2.shape[0]].contiguous()
xb2 = xdes2[np.arange(R2) % xdes2.shape[0]]
timed("lti_qp_kernel<4,2>", lambda: _engine.solve_robot_qps(M2.Ad, M2.Bd, Xb2, xb2, 0.25, 100.0, ((-1.0, 22.0), (-1.0, 20.0)), rho=1.0,
                                                            lin=torch.full((R2, M2.T, 2), 10.0, dtype=F64, device=dev),
                                                            sbar=torch.zeros((R2, M2.T, 2), dtype=F64, device=dev)),
      note=f"{R2} robot QPs (the script's 4 robots replicated), T={M2.T}, 2-D double integrator")
X3 = M3.x_initial(M3.x_ini, M3.x_des)
Xd3 = torch.as_tensor(np.stack([X3[r] for r in M3.robots_name]), device=dev)
xdes3 = np.stack([M3.x_des[r][:M3.n] for r in M3.robots_name])
h3, g3 = _engine.collision_tables(Xd3[:, :, :3].contiguous(), M3.R)
r3 = timed("lti_qp_kernel<6,3>", lambda: _engine.solve_robot_qps(M3.Ad, M3.Bd, Xd3, xdes3, 0.25, 1.0, ((-1.0, 22.0), (-1.0, 20.0)), col_h=h3, col_g=g3, c_S=1e4),
           note=f"the script's {Xd3.shape[0]} robots, T={M3.T}, collision rows with one slack per step")
print(json.dumps({"lti_qp<6,3> status": r3[2].tolist(), "iters": r3[3].tolist()}))
timed("ADMM_decentralized.x_traj_opt, one sweep: lti_qp_kernel<4,2> (4 robots) + sbar_enum_kernel",
      lambda: M2.x_traj_opt({k: v.copy() for k, v in X2.items()}, 0.25, n_admm=1))
print(json.dumps({"intersample sigma (min, median, max)": [float(si_.min()), float(si_.median()), float(si_.max())]}))
print(json.dumps({"done": True}))
