"""Multi-agent ADMM consensus runs (BASELINE configs 3 and 4) on 1..N GPUs:  torchrun --nproc-per-node N tools/run_admm_multi.py
Prints one JSON line per config on rank 0: per-round device time (max over ranks), residual histories, min separation,
and a checksum of the final trajectories (must not depend on the number of ranks)."""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from scvx_b200.batch import BatchedADMM
from scvx_b200.models.unicycle_model import UnicycleModel
from scvx_b200.models.single_integrator_model import SingleIntegratorModel

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def config3():
    N, K = 16, 100
    ang = np.linspace(0, 2 * np.pi, N, endpoint=False)
    models = [UnicycleModel(r_init=np.array([8 * np.cos(a), 8 * np.sin(a), 0.0]), r_final=np.array([-8 * np.cos(a), -8 * np.sin(a), 0.0]),
                            obstacles=[([0.0, 0.0], 1.0)]) for a in ang]
    return "config3: 16 unicycle agents on a circle r=8 -> antipodes, K=100, d_min=0.5, rho=1, 10 rounds, all pairs", models, K, 0.5, 20.0, False, 2


def config4(N=256):
    K = 100
    i = np.arange(N) + 0.5
    phi = np.arccos(1 - 2 * i / N); th = np.pi * (1 + 5 ** 0.5) * i
    pts = 8 * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
    models = [SingleIntegratorModel(r_init=q, r_final=-q, obstacles=[([0.0, 0.0, 0.0], 1.0)]) for q in pts]
    return f"config4: {N} single-integrator agents on a Fibonacci sphere r=8 -> antipodes, K=100, d_min=0.5, rho=1, 10 rounds, all pairs", models, K, 0.5, 20.0, True, 3


def run(name, models, K, d_min, sigma, si, d, rounds=10):
    eng = BatchedADMM(models, d_min, K, rho_admm=1.0, max_iter=rounds, si_variant=si)
    N = len(models)
    XU = [m.initialize_trajectory(np.zeros((3, K)), np.zeros((m.n_u, K))) for m in models]
    X0 = torch.as_tensor(np.stack([x for x, _ in XU])).to(dev); U0 = torch.as_tensor(np.stack([u for _, u in XU])).to(dev)
    eng.solve(X0, U0, sigma)                           # warm-up (allocations, NCCL channels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = eng.solve(X0, U0, sigma); b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    X = out["X"].cpu().numpy()
    P = X[:, :d, :]
    diff = P[:, None] - P[None]
    dist_ij = np.sqrt((diff ** 2).sum(axis=2)) + np.eye(N)[:, :, None] * 1e9
    if rank == 0:
        print(json.dumps({"config": name, "n_gpus": world, "ms_total": float(ms.item()), "ms_per_round": float(ms.item()) / rounds,
                          "agent_iterations_per_sec": N * rounds / (float(ms.item()) * 1e-3),
                          "primal_hist": [round(v, 6) for v in out["primal_hist"]], "dual_hist": [round(v, 6) for v in out["dual_hist"]],
                          "min_separation": float(dist_ij.min()), "allgather_bytes_per_round": int(N * (3 + models[0].n_u) * K * 8),
                          "sha1_X": hashlib.sha1(np.ascontiguousarray(X).tobytes()).hexdigest()[:16]}))


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "3"):
    run(*config3())
if which in ("all", "4"):
    run(*config4(int(sys.argv[2]) if len(sys.argv) > 2 else 256))
if world > 1:
    dist.destroy_process_group()
