"""Profiling driver: N agents of the bench workload, a few outer iterations (used under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from scvx_b200.batch import BatchedSCvx
from scvx_b200.models.unicycle_model import UnicycleModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scenes = bench.make_scenes(n, 0)
models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in scenes]
eng = BatchedSCvx(models, 100, max_iter=iters)
out = eng.solve(early_exit=False)
torch.cuda.synchronize()
print("ok", out["ipm_iters"].double().mean(dim=1).tolist(), (out["status"] == 0).double().mean().item())
