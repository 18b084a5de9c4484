"""Dev script: GPU IPM vs HiGHS oracle on a few seeded sub-problems (run on the GPU box)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import models as omodels, subproblem as ospb, ipm_struct
import helpers

dev = torch.device("cuda:0")
def report(tag, pairs):
    ps = [p for p, _ in pairs]
    t = time.time(); ws = helpers.solve_batch_on_gpu(ps, dev); dt = time.time() - t
    for i, (p, r) in enumerate(pairs):
        X, U, s = ws.X[i].cpu().numpy(), ws.U[i].cpu().numpy(), ws.sigma[i].item()
        e = ospb.evaluate(p, X, U, s)
        print(f"{tag}[{i}] st={ws.status[i].item()} it={ws.iters[i].item()} opt={r['obj']:.6f} gpu_obj={ws.objective[i].item():.6f} "
              f"eval={e['obj']:.6f} rel={(e['obj']-r['obj'])/abs(r['obj']):.2e} viol={e['viol']:.1e} ({dt*1e3:.1f} ms batch)")

report("uni50", helpers.make_problem_sequence(omodels.unicycle(), 50, 6))
rng = np.random.default_rng(0)
report("cfg2", [pr for a in range(3) for pr in helpers.make_problem_sequence(helpers.random_unicycle_scene(rng), 100, 2)])
report("SI40", helpers.make_problem_sequence(omodels.single_integrator(), 40, 3))
