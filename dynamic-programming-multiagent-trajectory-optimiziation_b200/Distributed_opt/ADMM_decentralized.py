"""GPU mirror of Distributed_opt/ADMM_decentralized.py (2-D double integrator, decentralised SCvx + 3-block ADMM).

Same module-level names as the script -- `descete_f(dt)`, `x_traj_opt(X_traj, trust_region)`, `x_initial(x_ini, x_des)` and
the globals `Tf, T0, T, dt, n, m, N_agents, robots_name, R, x_ini, x_des, Ad, Bd, trust_region, max_iter`
(ADMM_decentralized.py:215-242) -- so code written against the script keeps working; the functions read the globals at call
time exactly as the script's do.  Plotting is not mirrored.  Every cvxpy/CLARABEL solve of one ADMM sweep runs as ONE batched
launch over all robots: `scvx_lti_qp_batched` for the per-robot QP (:52-98) and `scvx_sbar_qp_batched` for the per-time-step
consensus QP (:106-139).
"""
import numpy as np
import torch
from numpy import linalg as LA

from . import _engine

# ---- global constants (ADMM_decentralized.py:215-242) ----------------------------------------------------------------
Tf = 40
T0 = 0
T = 81
t_traj = np.linspace(T0, Tf, T)
dt = t_traj[1] - t_traj[0]
n = 4
m = 2
trust_region = 0.25
max_iter = 1000
N_agents = 4
robots_name = ["robot01", "robot02", "robot03", "robot04"]
R = 2.3
x_ini = {}
x_des = {}
for _count, _name in enumerate(robots_name):
    x_ini[_name] = np.array([0, _count * 5.1, 0, 0, 0, 0], dtype=float)
    x_des[_name] = np.array([14 - _count * 0.2, (N_agents - _count - 1) * 5, 0, 0, 0, 0], dtype=float)


def descete_f(dt):
    """Exact ZOH of the 2-D double integrator (ADMM_decentralized.py:14-29) -> [Ad, Bd]."""
    return _engine.zoh_double_integrator(dt, n, m)


[Ad, Bd] = descete_f(dt)


def x_initial(x_ini, x_des):
    """ADMM_decentralized.py:174-180."""
    return {name: np.linspace(x_ini[name], x_des[name], T) for name in robots_name}


def x_traj_opt(X_traj, trust_region, n_admm=5, rho=1, verbose=False):
    """ADMM_decentralized.py:32-170.  X_traj: dict name -> (T, n+m) array; returns the dict with X_traj[name] += s_val[name]
    (updated in place like the script).  `last_log` holds the per-iteration consensus differences the script prints."""
    global last_log
    dev = torch.device("cuda")
    names = list(robots_name)
    Rn = len(names)
    X = torch.as_tensor(np.stack([np.asarray(X_traj[k], dtype=np.float64) for k in names])).to(dev)      # (R, T, n+m)
    xd = np.stack([x_des[k][:n] for k in names])
    r_all = torch.full((Rn, T, 2), 10.0, dtype=torch.float64, device=dev)                                   # duals (:40)
    s_bar = torch.zeros((Rn, T, 2), dtype=torch.float64, device=dev)
    col_h, col_g = _engine.collision_tables(X[:, :, :2].contiguous(), R)                                   # outer iterate (:126-137)
    s_val = torch.zeros_like(X)
    last_log = []
    for _ in range(n_admm):
        # (1) per-robot QP in s = (d, w)   (:52-98)
        s_val, obj, status, iters, _ = _engine.solve_robot_qps(Ad, Bd, X, xd, trust_region, 100.0,
                                                                ((-1.0, 22.0), (-1.0, 20.0)), rho=float(rho), lin=r_all, sbar=s_bar)
        if int((status == 2).sum().item()):
            raise RuntimeError("x_traj_opt: per-robot QP failed numerically")
        # (2) per-robot, per-time-step consensus QP in s_bar   (:106-139)
        s_pos = s_val[:, :, :2].contiguous()
        s_bar_new, _S = _engine.solve_sbar_qps(s_pos, r_all, float(rho), col_h, col_g, 1e6)
        # (3) dual ascent   (:148-155)
        r_all = r_all + rho * (s_pos - s_bar_new)
        diff = sum(LA.norm((s_bar_new[i] - s_pos[i]).cpu().numpy(), 2) for i in range(Rn))
        last_log.append(float(diff))
        if verbose:
            print("Difference:  ", diff)
        s_bar = s_bar_new
    S = s_val.cpu().numpy()
    for i, k in enumerate(names):
        X_traj[k] += S[i]
    return X_traj


last_log = []
