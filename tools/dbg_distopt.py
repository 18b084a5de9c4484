import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import distopt as od
from scvx_b200.Distributed_opt import _engine
import scvx_b200.Distributed_opt.ADMM_decentralized as M2
import scvx_b200.Distributed_opt.dist_scvx_3d as M3
cuda = torch.device("cuda")
# ---- 3D QPs
X = M3.x_initial(M3.x_ini, M3.x_des); names = M3.robots_name
Xd = torch.as_tensor(np.stack([X[k] for k in names])).to(cuda)
h, g = _engine.collision_tables(Xd[:, :, :3].contiguous(), M3.R)
s, obj, status, iters, S = _engine.solve_robot_qps(M3.Ad, M3.Bd, Xd, np.stack([M3.x_des[k][:6] for k in names]), 0.25, 1.0, ((-1.0, 22.0), (-1.0, 20.0)), col_h=h, col_g=g, c_S=1e4)
print('3D status', status.tolist(), iters.tolist(), obj.tolist(), flush=True)
s = s.cpu().numpy()
for i, k in enumerate(names):
    ho, go = od.collision_rows(X, names, k, M3.R, 3, M3.T)
    q = od.RobotQP(M3.Ad, M3.Bd, X[k][:, :6], X[k][:M3.T - 1, 6:], M3.x_des[k], 0.25, 1.0, col_h=ho[:M3.T - 1], col_g=go[:M3.T - 1], c_S=1e4)
    f0, lb, viol, ok = od.qp_bracket(q, s[i][:, :6], s[i][:, 6:])
    print('   ', k, 'f', f0, 'lb', lb, 'gap', f0 - lb, 'viol', viol, flush=True)
# ---- sbar
rng = np.random.default_rng(2)
X = M2.x_initial(M2.x_ini, M2.x_des); names = M2.robots_name
X[names[1]][:, 1] = X[names[0]][:, 1] + 1.0
Xd = torch.as_tensor(np.stack([X[k] for k in names])).to(cuda)
h, g = _engine.collision_tables(Xd[:, :, :2].contiguous(), M2.R)
s_pos = 0.3 * rng.normal(size=(4, M2.T, 2)); r = 10.0 + rng.normal(size=(4, M2.T, 2))
sb, S = _engine.solve_sbar_qps(torch.as_tensor(s_pos).to(cuda), torch.as_tensor(r).to(cuda), 1.0, h, g, 1e6)
sb, S = sb.cpu().numpy(), S.cpu().numpy()
for i, k in enumerate(names[:3]):
    ho, go = od.collision_rows(X, names, k, M2.R, 2, M2.T)
    want, Sw = od.solve_sbar_qp(s_pos[i], r[i], 1.0, ho, go)
    err = np.abs(sb[i] - want).max(axis=1)
    print('sbar', k, 'max err', err.max(), 'n>1e-5', int((err > 1e-5).sum()), 'worst t', int(err.argmax()), sb[i][err.argmax()], want[err.argmax()], flush=True)
# ---- 2D x_traj_opt sweeps
X = M2.x_initial(M2.x_ini, M2.x_des)
Xd = torch.as_tensor(np.stack([X[k] for k in names])).to(cuda)
xd = np.stack([M2.x_des[k][:4] for k in names])
r_all = torch.full((4, M2.T, 2), 10.0, dtype=torch.float64, device=cuda); s_bar = torch.zeros((4, M2.T, 2), dtype=torch.float64, device=cuda)
col_h, col_g = _engine.collision_tables(Xd[:, :, :2].contiguous(), M2.R)
Xo = {k: v.copy() for k, v in X.items()}
ro = {k: np.ones((M2.T, 2)) * 10 for k in names}; sbo = {k: np.zeros((M2.T, 2)) for k in names}
for sweep in range(3):
    s_val, obj, status, iters, _ = _engine.solve_robot_qps(M2.Ad, M2.Bd, Xd, xd, 0.25, 100.0, ((-1.0, 22.0), (-1.0, 20.0)), rho=1.0, lin=r_all, sbar=s_bar)
    sv = s_val.cpu().numpy()
    print('2D sweep', sweep, 'status', status.tolist(), iters.tolist(), flush=True)
    for i, k in enumerate(names):
        q = od.RobotQP(M2.Ad, M2.Bd, Xo[k][:, :4], Xo[k][:M2.T - 1, 4:], M2.x_des[k], 0.25, 100.0, rho=1.0, lin=r_all[i].cpu().numpy(), sbar=s_bar[i].cpu().numpy())
        f0, lb, viol, ok = od.qp_bracket(q, sv[i][:, :4], sv[i][:, 4:])
        print('    ', k, 'f', f0, 'gap', f0 - lb, 'viol', viol, 'gpuobj', obj[i].item(), flush=True)
    s_pos_t = s_val[:, :, :2].contiguous()
    sbn, _S = _engine.solve_sbar_qps(s_pos_t, r_all, 1.0, col_h, col_g, 1e6)
    # oracle sbar on the same inputs
    for i, k in enumerate(names):
        ho, go = od.collision_rows(Xo, names, k, M2.R, 2, M2.T)
        want, Sw = od.solve_sbar_qp(s_pos_t[i].cpu().numpy(), r_all[i].cpu().numpy(), 1.0, ho, go)
        print('     sbar', k, 'max err', np.abs(sbn[i].cpu().numpy() - want).max(), flush=True)
    r_all = r_all + 1.0 * (s_pos_t - sbn); s_bar = sbn
