"""Experiment behind DESIGN 4.3 (round 2, session 3): barrier-start policy of the warm sub-problems on the bench workload
(1024 agents, 8 lanes, one CUDA graph per lane), with the second-order weight and the adaptive step fraction in the kernel.
Prints ms/step over 20 timed steps, mean / max interior-point iterations over the timed steps' last one, non-optimal statuses
over all steps and the sigma checksum, per policy (easy_max_iters, mu0_easy, mu0_hard).

    python tools/exp_mu0_policy.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from scvx_b200 import batch as B
from scvx_b200.models.unicycle_model import UnicycleModel

scenes = bench.make_scenes(1024, 0)
models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in scenes]
POLICIES = [None, (10, 0.1, 10.0), (1000, 0.1, 10.0), (1000, 1e-2, 10.0), (1000, 1e-3, 10.0), (1000, 1e-4, 10.0), (20, 1e-3, 10.0), (20, 1e-3, 0.1),
            (1000, 1e-3, 10.0)]
if len(sys.argv) > 1:
    POLICIES = [eval(a) for a in sys.argv[1:]]
for pol in POLICIES:
    if pol is not None:
        B.MU0_POLICY = pol
    P = B.PipelinedSCvx(models, 100, n_lanes=int(os.environ.get('LANES', '8')), max_iter=23, adaptive_mu0=pol is not None).start()
    bad = 0
    for _ in range(3):
        P.run(1); torch.cuda.synchronize()
        bad += int((P.status() != 0).sum().item())
    P.build_lane_graphs()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(); P.run_lane_graphs(19, keep_history=False); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 19
    bad += int((P.status() != 0).sum().item())
    P.run_lane_graphs(1, keep_history=False); torch.cuda.synchronize()
    it = P.ipm_iters().double()
    print(f"policy={pol}: {ms:.3f} ms/step; last step: mean {it.mean().item():.2f} max {int(it.max().item())} IPM iterations; "
          f"non-optimal statuses seen {bad} (+{int((P.status() != 0).sum().item())} last); sigma checksum {sum(s[2].sum().item() for s in P.state):.6f}", flush=True)
