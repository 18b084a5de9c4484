"""Timing of the Distributed_opt mirrors (rows a14/a15): shipped scenarios and a 256-robot synthetic 3-D scene.
Prints JSON lines; the CPU line is the oracle port (HiGHS QP) on ONE call of the shipped scenario."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import scvx_b200.Distributed_opt.ADMM_decentralized as M2
import scvx_b200.Distributed_opt.dist_scvx_3d as M3
from scvx_b200.Distributed_opt import _engine
from oracle import distopt as od

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

X2 = M2.x_initial(M2.x_ini, M2.x_des)
ms2 = timed(lambda: M2.x_traj_opt({k: v.copy() for k, v in X2.items()}, 0.25))
t0 = time.perf_counter(); od.x_traj_opt_2d({k: v.copy() for k, v in X2.items()}, 0.25, M2.robots_name, M2.x_des, M2.Ad, M2.Bd, M2.T); cpu2 = time.perf_counter() - t0
print(json.dumps({"workload": "ADMM_decentralized.x_traj_opt, shipped scenario (4 robots, T=81, 5 ADMM sweeps = 20 robot QPs + 1620 consensus QPs)",
                  "gpu_ms_per_call": ms2, "cpu_oracle_s_per_call": cpu2, "speedup": cpu2 * 1e3 / ms2}))
X3 = M3.x_initial(M3.x_ini, M3.x_des)
ms3 = timed(lambda: M3.x_traj_opt({k: v.copy() for k, v in X3.items()}, 0.25))
print(json.dumps({"workload": "dist_scvx_3d.x_traj_opt, shipped scenario (3 robots, T=51)", "gpu_ms_per_call": ms3}))
# synthetic: R robots on a Fibonacci sphere of radius 30 around (10, 10, 10) -> antipodes, 3-D double integrator, T = 51
R = 256
i = np.arange(R) + 0.5; phi = np.arccos(1 - 2 * i / R); th = np.pi * (1 + 5 ** 0.5) * i
pts = 9.0 * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1) + np.array([10.0, 9.5, 10.0])
# goals: the antipode rotated by 30 degrees about z -- exact antipodes would put all robots AT the centre at mid time
# (coincident positions: the scripts' collision normal has no epsilon in its denominator and is NaN there)
ctr = np.array([10.0, 9.5, 10.0]); ca, sa = np.cos(np.pi / 6), np.sin(np.pi / 6)
Rz = np.array([[ca, -sa, 0.0], [sa, ca, 0.0], [0.0, 0.0, 1.0]])
goals = ctr - (pts - ctr) @ Rz.T
Xs = np.stack([np.linspace(np.concatenate([p, np.zeros(6)]), np.concatenate([g, np.zeros(6)]), 51) for p, g in zip(pts, goals)])
Xd = torch.as_tensor(Xs).cuda()
xdes = np.stack([np.concatenate([g, np.zeros(3)]) for g in goals])
def big():
    h, g = _engine.collision_tables(Xd[:, :, :3].contiguous(), 0.5)
    return _engine.solve_robot_qps(M3.Ad, M3.Bd, Xd, xdes, 0.25, 1.0, ((-1.0, 22.0), (-1.0, 20.0)), col_h=h, col_g=g, c_S=1e4)
msb = timed(big, reps=3)
s, obj, status, iters, S = big()
print(json.dumps({"workload": f"dist_scvx_3d-style QPs, {R} robots x T=51, all-pairs collision rows ({R-1} per step), one launch",
                  "gpu_ms_per_call": msb, "robot_qps_per_sec": R / (msb * 1e-3), "status_optimal_frac": float((status == 0).double().mean().item()),
                  "ipm_iters_mean": float(iters.double().mean().item())}))
