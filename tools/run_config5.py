"""BASELINE config 5 scale, independent-agent form: 8192 unicycle agents x K=200, M=32 discs per agent drawn from a shared
field of 512 (seed 1).  Times a few outer iterations on one GPU (or one shard per rank under torchrun: strong scaling)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from scvx_b200.batch import BatchedSCvx, shard_bounds
from scvx_b200.models.unicycle_model import UnicycleModel
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N, K, M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 200, 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rng = np.random.default_rng(1)
field_c = rng.uniform(-7, 7, (512, 2)); field_r = rng.uniform(0.3, 1.0, 512)
models = []
for a in range(N):
    y0 = rng.uniform(-9, 9); start = np.array([rng.uniform(-9, -8), y0, 0.0]); goal = np.array([-start[0], -y0, 0.0])
    d = np.minimum(np.linalg.norm(field_c - start[:2], axis=1), np.linalg.norm(field_c - goal[:2], axis=1)) - field_r
    ok = np.where(d > 1.0)[0]
    pick = rng.choice(ok, M, replace=False)
    models.append(UnicycleModel(r_init=start, r_final=goal, obstacles=[(list(field_c[j]), float(field_r[j])) for j in pick]))
per, i0, i1 = shard_bounds(N, world, rank)
eng = BatchedSCvx(models[i0:i1], K, max_iter=iters + 1)
b = eng.batch
X, U = b.initial_trajectories()
n = b.n
F64 = torch.float64
sig = torch.ones(n, dtype=F64, device=b.device); tr = torch.full((n,), 100.0, dtype=F64, device=b.device)
act = torch.ones(n, dtype=torch.int32, device=b.device); met = torch.zeros((iters + 1, n, 6), dtype=F64, device=b.device)
eng.iterate(X, U, sig, tr, act, met[0]); torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a_.record()
its = []
for it in range(iters):
    eng.iterate(X, U, sig, tr, act, met[it + 1]); its.append(eng.ws.iters.double().mean().item())
b_.record(); torch.cuda.synchronize()
ms = torch.tensor([a_.elapsed_time(b_)], dtype=F64, device=b.device)
ok = torch.tensor([(eng.ws.status == 0).double().sum().item()], dtype=F64, device=b.device)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(ok)
if rank == 0:
    print(json.dumps({"workload": f"config5 (independent-agent form): {N} unicycle agents x K={K}, M={M} of a shared field of 512",
                      "n_gpus": world, "ms_per_outer_iteration": ms.item() / iters, "agent_iterations_per_sec": N * iters / (ms.item() * 1e-3),
                      "ipm_iters_mean": its, "optimal_frac_last": ok.item() / N}))
if world > 1:
    dist.destroy_process_group()
