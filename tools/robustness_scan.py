"""Run the bench workload for N outer iterations and report solver status / IPM iteration statistics per outer iteration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from scvx_b200 import batch as _B
from scvx_b200.batch import BatchedSCvx
if os.environ.get('SCVX_MU0_POLICY'):
    _B.MU0_POLICY = eval(os.environ['SCVX_MU0_POLICY'])
from scvx_b200.models.unicycle_model import UnicycleModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 25
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
scenes = bench.make_scenes(n, seed)
models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in scenes]
eng = BatchedSCvx(models, 100, max_iter=iters, adaptive_mu0=(os.environ.get('SCVX_ADAPTIVE_MU0', '1') != '0'))
eng.retry_stranded = os.environ.get('SCVX_RETRY', '1') != '0'
out = eng.solve(early_exit=False)
torch.cuda.synchronize()
st = out["status"].cpu().numpy(); it = out["ipm_iters"].cpu().numpy(); met = out["metrics"].cpu().numpy()
for o in range(iters):
    bad = np.where(st[o] != 0)[0]
    print(f"outer {o:2d}: ipm iters mean {it[o].mean():5.2f} max {it[o].max():3d} | status!=0: {len(bad)} {[(int(b), int(st[o][b]), int(it[o][b])) for b in bad[:6]]} | finite X: {bool(torch.isfinite(out['X']).all())}")
print("total non-optimal", int((st != 0).sum()), "of", st.size, "| active at end", int(out["active"].sum().item()))
print("nu_norm median per outer", np.median(met[:, :, 0], axis=1).round(4).tolist())
print("sigma median per outer", np.median(met[:, :, 5], axis=1).round(2).tolist())
