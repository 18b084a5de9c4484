"""Experiment: does an ipm_kernel block run faster ALONE on its SM than paired with a second block?
n identical copies of one hard bench scene; n = 148 (one block per SM) vs n = 296 (two per SM); CUDA-event time of the
sub-problem kernel per outer iteration (all blocks do identical work, so kernel time = one block's solve time)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from scvx_b200.batch import BatchedSCvx
from scvx_b200.models.unicycle_model import UnicycleModel
which = int(sys.argv[1]) if len(sys.argv) > 1 else 5
o = bench.make_scenes(which + 1, 0)[which]
dev = torch.device("cuda"); F64 = torch.float64
for n in (148, 296, 74, 148, 296):
    models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for _ in range(n)]
    eng = BatchedSCvx(models, 100, max_iter=8)
    b = eng.batch
    X, U = b.initial_trajectories()
    sig = torch.ones(n, dtype=F64, device=dev); tr = torch.full((n,), 100.0, dtype=F64, device=dev)
    act = torch.ones(n, dtype=torch.int32, device=dev); met = torch.zeros((8, n, 6), dtype=F64, device=dev)
    rows = []
    for it in range(8):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        eng.iterate(X, U, sig, tr, act, met[it], solver_events=ev)
        torch.cuda.synchronize()
        rows.append((int(eng.ws.iters.max().item()), round(ev[0].elapsed_time(ev[1]), 3)))
    print(n, "blocks: (ipm iterations, kernel ms) per outer iteration:", rows, flush=True)
