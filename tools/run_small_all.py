"""Small end-to-end run of every kernel (for compute-sanitizer): SCvx batch (unicycle + SI), ADMM round, Distributed_opt."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from scvx_b200.batch import BatchedSCvx, BatchedADMM
from scvx_b200.models.unicycle_model import UnicycleModel
from scvx_b200.models.single_integrator_model import SingleIntegratorModel
K = 24
out = BatchedSCvx([UnicycleModel(), UnicycleModel(obstacles=[])], K, max_iter=2).solve(early_exit=False)
print("scvx unicycle", out["status"].tolist(), out["ipm_iters"].tolist())
out = BatchedSCvx([SingleIntegratorModel()], K, max_iter=2).solve(early_exit=False)
print("scvx SI", out["status"].tolist(), out["ipm_iters"].tolist())
ang = np.linspace(0, 2 * np.pi, 3, endpoint=False)
ms = [UnicycleModel(r_init=np.array([4 * np.cos(a), 4 * np.sin(a), 0.0]), r_final=np.array([-4 * np.cos(a), -4 * np.sin(a), 0.0]), obstacles=[([0.0, 0.0], 0.5)]) for a in ang]
XU = [m.initialize_trajectory(np.zeros((3, K)), np.zeros((2, K))) for m in ms]
X0 = torch.as_tensor(np.stack([x for x, _ in XU])).cuda(); U0 = torch.as_tensor(np.stack([u for _, u in XU])).cuda()
o = BatchedADMM(ms, 0.5, K, max_iter=2).solve(X0, U0, 12.0)
print("admm", o["primal_hist"])
import scvx_b200.Distributed_opt.ADMM_decentralized as M2
import scvx_b200.Distributed_opt.dist_scvx_3d as M3
X = M2.x_initial(M2.x_ini, M2.x_des); M2.x_traj_opt(X, 0.25, n_admm=1); print("distopt 2d", M2.last_log)
X = M3.x_initial(M3.x_ini, M3.x_des); M3.x_traj_opt(X, 0.25); print("distopt 3d", M3.last_objective.tolist())
torch.cuda.synchronize()
