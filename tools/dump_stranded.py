"""Find warm sub-problems on which the small barrier start strands (status != optimal with the retry pass off) and dump them
for the numpy twin: python tools/dump_stranded.py [n=4096] [iters=12] [seed=3] -> gpurun_out/stranded_<seed>.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from scvx_b200.batch import BatchedSCvx
from scvx_b200.models.unicycle_model import UnicycleModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scenes = bench.make_scenes(n, seed)
models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in scenes]
eng = BatchedSCvx(models, 100, max_iter=iters, adaptive_mu0=True)
eng.retry_stranded = False
b = eng.batch
dev = b.device
X, U = b.initial_trajectories()
sig = torch.ones(n, dtype=torch.float64, device=dev); tr = torch.full((n,), 100.0, dtype=torch.float64, device=dev)
act = torch.ones(n, dtype=torch.int32, device=dev); met = torch.zeros((n, 6), dtype=torch.float64, device=dev)
dump = {}
prev_iters = None
for o in range(iters):
    Xr, Ur, sr, trr = X.clone(), U.clone(), sig.clone(), tr.clone()
    mu0 = None if eng._mu0 is None else eng._mu0.clone()
    eng.iterate(X, U, sig, tr, act, met)
    torch.cuda.synchronize()
    bad = torch.nonzero(eng.ws.status != 0).flatten().tolist()
    for i in bad:
        key = f"o{o}_a{i}"
        print("stranded", key, "iters", int(eng.ws.iters[i]), "prev iters", None if prev_iters is None else int(prev_iters[i]), "mu0", None if mu0 is None else float(mu0[i]))
        dump[key + "_mats"] = np.array([m[i].cpu().numpy().ravel() for m in eng.mats], dtype=object)
        for nm, t in (("X", Xr), ("U", Ur)):
            dump[key + "_" + nm] = t[i].cpu().numpy()
        dump[key + "_scal"] = np.array([float(sr[i]), float(trr[i]), float(mu0[i]) if mu0 is not None else 10.0, i, seed])
    prev_iters = eng.ws.iters.clone()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", f"stranded_{seed}.npz"), **dump, allow_pickle=True)
print("dumped", len(dump) // 4)
