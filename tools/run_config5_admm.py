"""BASELINE config 5 in its decentralised form: N unicycle agents x K=200, M=32 discs per agent from a shared field of 512
(seed 1), ADMM consensus rounds with each agent coupled to its k nearest neighbours, agents sharded over the ranks and ONE
all-gather of (X, U) per round.  torchrun --nproc-per-node G tools/run_config5_admm.py [N] [rounds] [k]"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from scvx_b200.batch import BatchedADMM
from scvx_b200.models.unicycle_model import UnicycleModel
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
knn = int(sys.argv[3]) if len(sys.argv) > 3 else 16
K, M = 200, 32
rng = np.random.default_rng(1)
field_c = rng.uniform(-7, 7, (512, 2)); field_r = rng.uniform(0.3, 1.0, 512)
models = []
for a in range(N):
    y0 = rng.uniform(-9, 9); start = np.array([rng.uniform(-9, -8), y0, 0.0]); goal = np.array([-start[0], -y0, 0.0])
    d = np.minimum(np.linalg.norm(field_c - start[:2], axis=1), np.linalg.norm(field_c - goal[:2], axis=1)) - field_r
    pick = rng.choice(np.where(d > 1.0)[0], M, replace=False)
    models.append(UnicycleModel(r_init=start, r_final=goal, obstacles=[(list(field_c[j]), float(field_r[j])) for j in pick], robot_radius=0.05))
eng = BatchedADMM(models, 0.1, K, rho_admm=1.0, max_iter=rounds, neighbor_k=knn, neighbor_radius=1.0)
XU = [m.initialize_trajectory(np.zeros((3, K)), np.zeros((2, K))) for m in models]
X0 = torch.as_tensor(np.stack([x for x, _ in XU])).to(dev); U0 = torch.as_tensor(np.stack([u for _, u in XU])).to(dev)
eng.max_iter = 1; eng.solve(X0, U0, 20.0); eng.max_iter = rounds          # warm-up (allocations, NCCL channels)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = eng.solve(X0, U0, 20.0); b.record(); torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
ok = torch.tensor([float((eng.ws.status == 0).sum().item()) if eng.nl else 0.0], dtype=torch.float64, device=dev)
nact = torch.tensor([float((eng.last_nbr_idx >= 0).sum().item()) if eng.nl else 0.0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(ok); dist.all_reduce(nact)
if rank == 0:
    X = out["X"].cpu().numpy()
    print(json.dumps({"workload": f"config5 (decentralised): {N} unicycle agents x K={K}, M={M} of a shared field of 512, ADMM consensus with the "
                      f"{knn} nearest neighbours within 1.0, d_min=0.1", "n_gpus": world, "rounds": rounds, "ms_per_round": ms.item() / rounds,
                      "agent_iterations_per_sec": N * rounds / (ms.item() * 1e-3), "optimal_frac_last": ok.item() / N,
                      "mean_active_neighbours": nact.item() / N, "allgather_bytes_per_round": int(N * 5 * K * 8),
                      "primal_hist": [round(v, 6) for v in out["primal_hist"]], "sha1_X": hashlib.sha1(np.ascontiguousarray(X).tobytes()).hexdigest()[:16]}))
if world > 1:
    dist.destroy_process_group()
