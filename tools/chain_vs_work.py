"""What bounds a pipelined step?  From the interior-point iteration counts of 20 steady-state outer iterations (steps 3..22 of the
bench workload, adaptive barrier start): total work per step (agent-iterations / 296 block slots) against the hardest agent's own
chain and the hardest lane's chain (sum over steps of the lane's maximum), all in interior-point iterations per step.
    python tools/chain_vs_work.py [seed=0] [lanes=16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from scvx_b200.batch import BatchedSCvx, lane_bounds
from scvx_b200.models.unicycle_model import UnicycleModel
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = 1024
scenes = bench.make_scenes(n, seed)
models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in scenes]
eng = BatchedSCvx(models, 100, max_iter=23, adaptive_mu0=True)
out = eng.solve(early_exit=False)
it = out["ipm_iters"].cpu().numpy()[3:23].astype(float) + 0.4        # + the residual pass that detects convergence
steps = it.shape[0]
work = it.sum() / steps / 296.0
agent_chain = it.sum(axis=0).max() / steps
lane_chain = max(it[:, a:b].max(axis=1).sum() for a, b in lane_bounds(n, lanes)) / steps
# dealing the agents to the lanes by hardness measured on EARLIER steps (1..2): lane l gets ranks l, l + lanes, l + 2 lanes, ...
early = out["ipm_iters"].cpu().numpy()[1:3].sum(axis=0)
order = np.argsort(-early, kind="stable")
dealt = max(it[:, order[l::lanes]].max(axis=1).sum() for l in range(lanes)) / steps
print(f"seed {seed}: dealt by early hardness: hardest lane's chain = {dealt:.1f}")
print(f"seed {seed}: mean iterations/solve {it.mean() - 0.4:.2f}; per step: total work / 296 slots = {work:.1f} iterations, hardest agent's chain = "
      f"{agent_chain:.1f}, hardest of {lanes} lanes' chain = {lane_chain:.1f}; agents above 15 iterations/step on average: {(it.mean(axis=0) > 15).sum()}")
