import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import distopt as od
from scvx_b200.Distributed_opt import _engine
import scvx_b200.Distributed_opt.ADMM_decentralized as M
cuda = torch.device("cuda")
rng = np.random.default_rng(2)
X = M.x_initial(M.x_ini, M.x_des); names = M.robots_name
X[names[1]][:, 1] = X[names[0]][:, 1] + 1.0
Xd = torch.as_tensor(np.stack([X[k] for k in names])).to(cuda)
h, g = _engine.collision_tables(Xd[:, :, :2].contiguous(), M.R)
s_pos = 0.3 * rng.normal(size=(4, M.T, 2)); r = 10.0 + rng.normal(size=(4, M.T, 2))
sb, S = _engine.solve_sbar_qps(torch.as_tensor(s_pos).to(cuda), torch.as_tensor(r).to(cuda), 1.0, h, g, 1e6)
sb, S = sb.cpu().numpy(), S.cpu().numpy()
for i, k in enumerate(names):
    ho, go = od.collision_rows(X, names, k, M.R, 2, M.T)
    want, Sw = od.solve_sbar_qp(s_pos[i], r[i], 1.0, ho, go)
    err = np.abs(sb[i] - want).max(axis=1)
    bad = np.argsort(-err)[:3]
    for t in bad:
        st = s_pos[i][t] + r[i][t]
        print(k, t, 'err', err[t], 'gpu', sb[i][t], S[i][t], 'highs', want[t], Sw[t], 'viol@unc', ho[t] - go[t] @ st, 'viol@gpu', ho[t]-go[t]@sb[i][t], 'viol@highs', ho[t]-go[t]@want[t])
        f = lambda z, S_: -r[i][t] @ z + 0.5 * ((s_pos[i][t] - z) ** 2).sum() + 1e6 * S_
        print('    obj gpu', f(sb[i][t], max(0, (ho[t]-go[t]@sb[i][t]).max())), 'obj highs', f(want[t], max(0,(ho[t]-go[t]@want[t]).max())))
