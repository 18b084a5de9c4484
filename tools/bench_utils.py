"""Timing of the section-8(f) rows on one GPU (CUDA events, warm) beside the CPU oracle on a bounded sample.
usage: python tools/bench_utils.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import utils as ou
from scvx_b200 import _device, _lib
from scvx_b200.batch import BatchedNash
from scvx_b200.models.game_model import GameUnicycleModel


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


dev = torch.device("cuda")
rng = np.random.default_rng(0)
# ---- warm start: 8192 unicycle agents, K = 200, 32 discs each (config 5 sizes)
n, K, M = 8192, 200, 32
p0 = np.zeros((n, 3)); p1 = np.zeros((n, 3))
p0[:, 0] = rng.uniform(-9.5, -9, n); p0[:, 1] = rng.uniform(-9, 9, n); p1[:, 0] = -p0[:, 0]; p1[:, 1] = -p0[:, 1]
C = rng.uniform(-7, 7, (n, M, 2)); R = rng.uniform(0.2, 0.8, (n, M))
t = [torch.as_tensor(a, device=dev) for a in (p0, p1, C, R)]
ms = timed(lambda: _device.warm_start(_lib.MODEL_UNICYCLE, t[0], t[1], t[2], t[3], 0.3, K))
X0, U0, st = _device.warm_start(_lib.MODEL_UNICYCLE, t[0], t[1], t[2], t[3], 0.3, K)
ok = int((st == 0).sum().item())
c0 = time.perf_counter()
for i in range(64):
    try:
        ou.initial_guess_unicycle(p0[i], p1[i], [(list(c), float(r)) for c, r in zip(C[i], R[i])], 0.3, K)
    except ValueError:
        pass
cpu = (time.perf_counter() - c0) / 64
print(f"warm start   : {n} agents x K={K} x M={M}: {ms:.3f} ms/launch = {n / ms * 1e3:.3e} agents/s ({ok} ok); CPU oracle {cpu * 1e3:.2f} ms/agent = {1 / cpu:.1f} agents/s")
# ---- analysis metrics: N = 2048 agents, K = 200
N = 2048
X = torch.as_tensor(rng.normal(size=(N, 3, K)) * 5.0, device=dev)
ms = timed(lambda: _device.min_inter_agent_distance(X))
pairs = N * (N - 1) // 2
Xs = list(X[:64].cpu().numpy())
c0 = time.perf_counter(); ou.min_inter_agent_distance(Xs); cpu = (time.perf_counter() - c0) / (64 * 63 // 2)
print(f"pair distance: N={N}, K={K}: {ms:.3f} ms = {pairs / ms * 1e3:.3e} pairs/s, {pairs * 2 * 3 * K * 8 / ms / 1e6:.1f} GB/s algorithmic; CPU oracle {1 / cpu:.3e} pairs/s")
Co = torch.as_tensor(rng.uniform(-5, 5, (64, 3)), device=dev); Ro = torch.as_tensor(rng.uniform(0.2, 1, 64), device=dev)
ms = timed(lambda: _device.min_agent_obstacle_distance(X, Co, Ro, 0.5))
print(f"obstacle dist: N={N}, M=64, K={K}: {ms:.3f} ms = {N * 64 / ms * 1e3:.3e} (agent, obstacle)/s")
# ---- Nash best-response sweeps: 8 crossing pairs in parallel lanes (the shipped two-agent game, replicated), K = 50
Ng, Kg = 16, 50
models, Xr = [], np.zeros((Ng, 3, Kg))
Ur = np.zeros((Ng, 2, Kg))
s = np.linspace(0, 1, Kg)
for q in range(Ng // 2):
    for h, (x0, x1) in enumerate(((0.0, 2.0), (2.0, 0.0))):
        i = 2 * q + h
        r0 = np.array([x0 + 4.0 * q, -1.0, 0.0]); r1 = np.array([x1 + 4.0 * q, 3.0, 0.0])
        models.append(GameUnicycleModel(r_init=r0, r_final=r1, obstacles=[([1.0 + 4.0 * q, 1.0], 0.25)], control_weight=5.0,
                                        collision_radius=0.3, control_rate_weight=5.0, curvature_weight=100.0,
                                        bounds=(-5.0, 40.0), robot_radius=0.1))
        Xr[i, 0] = r0[0] + (r1[0] - r0[0]) * s + (0.5 * np.sin(np.pi * s) if h else 0.0)
        Xr[i, 1] = r0[1] + (r1[1] - r0[1]) * s
        Xr[i, 2, :-1] = np.arctan2(np.diff(Xr[i, 1]), np.diff(Xr[i, 0])); Xr[i, 2, -1] = Xr[i, 2, -2]
        Xr[i, 2, 0] = 0.0; Xr[i, 2, -1] = 0.0
bn = BatchedNash(models, Kg, max_iter=3)
Xd, Ud = torch.as_tensor(Xr, device=dev), torch.as_tensor(Ur, device=dev)
out = bn.solve(Xd, Ud, 8.0); torch.cuda.synchronize()
t0 = time.perf_counter(); bn.launches = 0; out = bn.solve(Xd, Ud, 8.0); torch.cuda.synchronize(); wall = time.perf_counter() - t0
sweeps = len(out["change_hist"])
print(f"Nash sweeps  : {Ng} agents, K={Kg}: {sweeps} sweeps, {bn.launches} best-response launches, {wall * 1e3:.1f} ms wall "
      f"({wall / max(bn.launches, 1) * 1e3:.2f} ms per launch of {Ng} best responses); infeasible flags {int(out['infeasible'].sum().item())}; "
      f"change {['%.2e' % c for c in out['change_hist']]}")
