"""Experiment: adaptive barrier start (BatchedSCvx(adaptive_mu0=True)) on the bench workload, pipelined in 4 lanes.
Prints ms/step over 20 timed steps (after 3 warm-up steps), mean / max interior-point iterations and the count of
non-optimal solver statuses over ALL steps, and the sigma checksum, for the default start and the adaptive one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from scvx_b200.batch import PipelinedSCvx
from scvx_b200.models.unicycle_model import UnicycleModel
scenes = bench.make_scenes(1024, 0)
models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in scenes]
for adaptive in (False, True, False, True):
    P = PipelinedSCvx(models, 100, n_lanes=4, max_iter=23, adaptive_mu0=adaptive).start()
    bad, its, mx = 0, 0.0, 0
    for _ in range(3):
        P.run(1); torch.cuda.synchronize()
        bad += int((P.status() != 0).sum().item())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); P.run(19); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 19
    bad += int((P.status() != 0).sum().item())
    P.run(1); torch.cuda.synchronize()
    it = P.ipm_iters().double()
    print(f"adaptive_mu0={adaptive}: {ms:.3f} ms/step; last step: mean {it.mean().item():.2f} max {int(it.max().item())} IPM iterations; "
          f"non-optimal statuses seen {bad}; sigma checksum {sum(s[2].sum().item() for s in P.state):.6f}", flush=True)
