"""Robustness of the Nash best-response path: N random crossing pairs (GameUnicycleModel, slab penalty 1e8), a few Jacobi sweeps;
reports solver status, IPM iterations, residual slab slack and infeasibility flags.  usage: python tools/robustness_scan_game.py [pairs] [K]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from scvx_b200.batch import BatchedNash
from scvx_b200.models.game_model import GameUnicycleModel
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rng = np.random.default_rng(7)
models, Xr = [], np.zeros((2 * pairs, 3, K)); Ur = np.zeros((2 * pairs, 2, K))
s = np.linspace(0, 1, K)
for q in range(pairs):
    ox = 6.0 * q
    w = rng.uniform(1.0, 3.0); hgt = rng.uniform(3.0, 5.0)
    ob = [([ox + rng.uniform(0.2, w - 0.2), rng.uniform(0.5, hgt - 1.5)], rng.uniform(0.1, 0.3))]
    for h, (x0, x1) in enumerate(((0.0, w), (w, 0.0))):
        i = 2 * q + h
        r0 = np.array([ox + x0, -1.0, 0.0]); r1 = np.array([ox + x1, hgt - 1.0, 0.0])
        models.append(GameUnicycleModel(r_init=r0, r_final=r1, obstacles=ob, control_weight=rng.uniform(1, 10), collision_radius=rng.uniform(0.2, 0.4),
                                        control_rate_weight=rng.uniform(1, 10), curvature_weight=rng.uniform(10, 200),
                                        inertia_weight=rng.choice([0.0, 0.3]), bounds=(-5.0, 6.0 * pairs + 5.0), robot_radius=0.1))
        Xr[i, 0] = r0[0] + (r1[0] - r0[0]) * s + (rng.uniform(0.4, 0.8) * np.sin(np.pi * s) if h else 0.0)
        Xr[i, 1] = r0[1] + (r1[1] - r0[1]) * s
        Xr[i, 2, :-1] = np.arctan2(np.diff(Xr[i, 1]), np.diff(Xr[i, 0])); Xr[i, 2, 0] = 0.0; Xr[i, 2, -1] = 0.0
import scvx_b200.batch as _b
if len(sys.argv) > 3:
    _b.SLAB_PENALTY = float(sys.argv[3]); _b.SLAB_PENALTIES = (float(sys.argv[3]),)
bn = BatchedNash(models, K, max_iter=3, neighbor_radius=(float(sys.argv[4]) if len(sys.argv) > 4 else None),
                 n_colors=(int(sys.argv[5]) if len(sys.argv) > 5 else 1))
dev = torch.device("cuda")
# neighbours far away (other lanes) make their slab rows inactive; all 2*pairs - 1 slots are still walked
out = bn.solve(torch.as_tensor(Xr, device=dev), torch.as_tensor(Ur, device=dev), 8.0)
torch.cuda.synchronize()
st = bn.ws.status.cpu().numpy(); it = bn.ws.iters.cpu().numpy()
print(f"{2 * pairs} game agents x K={K}, {len(out['change_hist'])} sweeps, {bn.launches} launches")
print("last launch: status counts", {int(v): int((st == v).sum()) for v in np.unique(st)}, "| IPM iterations mean %.1f max %d" % (it.mean(), it.max()))
print("iteration-cap flags per sweep", out["iteration_cap"].sum(dim=1).tolist())
print("infeasible flags per sweep", out["infeasible"].sum(dim=1).tolist(), "| ACS iterations mean per sweep", out["acs_iters"].double().mean(dim=1).tolist())
X = out["X"].cpu().numpy()
sep = [np.linalg.norm(X[2 * q, :2] - X[2 * q + 1, :2], axis=0).min() for q in range(pairs)]
rad = [max(models[2 * q].collision_radius, models[2 * q + 1].collision_radius) for q in range(pairs)]
print("pairs whose final separation is below the smaller collision radius - 1e-6:", int(sum(sp < min(models[2 * q].collision_radius, models[2 * q + 1].collision_radius) - 1e-6 for q, sp in enumerate(sep))),
      "| min separation %.4f" % min(sep), "| finite:", bool(np.isfinite(X).all()))
