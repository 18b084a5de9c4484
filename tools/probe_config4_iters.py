"""Which config-4 agents run long?  Per-round interior-point iteration counts (top of the distribution) and status."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench_configs as cfg
from scvx_b200.batch import BatchedADMM
c = cfg.config4_models(256)
dev = torch.device("cuda"); K = c["K"]
eng = BatchedADMM(c["models"], c["d_min"], K, rho_admm=1.0, max_iter=1, si_variant=True, ipm_max_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
XU = [m.initialize_trajectory(np.zeros((3, K)), np.zeros((m.n_u, K))) for m in c["models"]]
X0 = torch.as_tensor(np.stack([x for x, _ in XU])).to(dev); U0 = torch.as_tensor(np.stack([u for _, u in XU])).to(dev)
# replicate solve() round by round to look at iters: run with max_iter = r and take the last
for r in range(1, 6):
    eng.max_iter = r
    out = eng.solve(X0, U0, c["sigma"])
    it = eng.ws.iters.cpu().numpy(); st = eng.ws.status.cpu().numpy()
    top = np.argsort(-it)[:6]
    print(f"round {r}: mean {it.mean():.1f}, sorted top {[(int(a), int(it[a]), int(st[a])) for a in top]}, primal {out['primal_hist'][-1]:.4f}, "
          f"obj of top {[float(out['objective'][-1, a]) for a in top[:2]]}")
