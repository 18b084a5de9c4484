"""GPU mirror of Distributed_opt/dist_scvx_3d.py (3-D double integrator, Jacobi-style decentralised SCvx).

Same module-level names as the script -- `descete_f(dt)`, `x_traj_opt(X_traj, trust_region)`, `x_initial`, `cost_fcn` and the
globals of dist_scvx_3d.py:200-231.  One call to `x_traj_opt` solves every robot's QP (:51-111) in ONE launch of
`scvx_lti_qp_batched`, all robots linearising about the others' previous trajectories.
"""
import numpy as np
import torch

from . import _engine

Tf = 30
T0 = 0
T = 51
t_traj = np.linspace(T0, Tf, T)
dt = t_traj[1] - t_traj[0]
n = 6
m = 3
trust_region = 0.25
max_iter = 1000
N_agents = 3
robots_name = ["robot01", "robot02", "robot03"]
R = 2.3
x_ini = {}
x_des = {}
for _count, _name in enumerate(robots_name):
    x_ini[_name] = np.array([0, _count * 5.1, 10, 0, 0, 0, 0, 0, 0], dtype=float)
    x_des[_name] = np.array([14, (N_agents - _count - 1) * 5, 10 + _count * 1, 0, 0, 0, 0, 0, 0], dtype=float)


def descete_f(dt):
    """dist_scvx_3d.py:9-28 -> [Ad, Bd]."""
    return _engine.zoh_double_integrator(dt, n, m)


[Ad, Bd] = descete_f(dt)
cost_list = np.zeros(max_iter)


def x_initial(x_ini, x_des):
    """dist_scvx_3d.py:122-128."""
    return {name: np.linspace(x_ini[name], x_des[name], T) for name in robots_name}


def x_traj_opt(X_traj, trust_region):
    """dist_scvx_3d.py:31-118.  Returns the dict with X_traj[name] += s_val[name] (in place, like the script)."""
    global last_objective
    dev = torch.device("cuda")
    names = list(robots_name)
    X = torch.as_tensor(np.stack([np.asarray(X_traj[k], dtype=np.float64) for k in names])).to(dev)
    xd = np.stack([x_des[k][:n] for k in names])
    col_h, col_g = _engine.collision_tables(X[:, :, :3].contiguous(), R)
    s_val, obj, status, iters, _S = _engine.solve_robot_qps(Ad, Bd, X, xd, trust_region, 1.0, ((-1.0, 22.0), (-1.0, 20.0)),
                                                             col_h=col_h, col_g=col_g, c_S=1e4)
    if int((status == 2).sum().item()):
        raise RuntimeError("x_traj_opt: per-robot QP failed numerically")
    last_objective = obj.cpu().numpy()
    S = s_val.cpu().numpy()
    for i, k in enumerate(names):
        X_traj[k] += S[i]
    return X_traj


def cost_fcn(X_traj):
    """dist_scvx_3d.py:131-138: sum over robots and steps of |u_t|^2."""
    cost_iter = 0
    for name in robots_name:
        u_traj_i = X_traj[name][0:T - 1, n:n + m]
        cost_iter += float((u_traj_i ** 2).sum())
    return cost_iter


last_objective = None
