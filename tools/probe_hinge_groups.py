"""One GPU's share of config 4 on 8 GPUs: n agents (default 32), each with 255 inter-agent rows per stage -- one launch of the
sub-problem kernel per setting of SCVX_HINGE_GROUPS (the launch is under-subscribed: latency-bound, not bandwidth-bound)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
import test_subproblem_gpu as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rng = np.random.default_rng(4)
ps = [T._admm_problem("single_integrator", 256, 100, i, rng) for i in range(n)]
dev = torch.device("cuda")
from scvx_b200 import _device
ref = None
_orig = _device.solve_subproblem
_ev = []


def _timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = _orig(*a, **k); e1.record(); _ev.append((e0, e1))
    return out


_device.solve_subproblem = _timed
for g, c in ((1, 1), (2, 1), (4, 1), (2, 2), (2, 4)):
    os.environ["SCVX_HINGE_GROUPS"] = str(g); os.environ["SCVX_CLUSTER"] = str(c)
    helpers.solve_batch_on_gpu(ps[:2], dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ws = helpers.solve_batch_on_gpu(ps, dev)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    it = ws.iters.double()
    obj = ws.objective.cpu().numpy()
    if ref is None:
        ref = obj
    kms = _ev[-1][0].elapsed_time(_ev[-1][1])
    print(f"G={g} C={c}: kernel {kms:.2f} ms ({1e3 * dt:.1f} ms wall incl. upload); iterations mean {it.mean().item():.1f} max {it.max().item():.0f}; status ok "
          f"{(ws.status == 0).double().mean().item():.3f}; max rel objective difference to G=1 {np.abs(obj / ref - 1).max():.2e}", flush=True)
