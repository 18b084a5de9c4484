"""Config 4 probe: one ADMM round of 256 single-integrator agents (all pairs), interior-point iteration counts, round time, and
(with the -DSCVX_PHASE_TIMING build selected by SCVX_LIB) the per-phase cycles of block 0.
usage: python tools/probe_config4.py [N] [rounds]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench_configs as cfg
from scvx_b200 import _lib
from scvx_b200.batch import BatchedADMM
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = cfg.config4_models(N)
dev = torch.device("cuda")
eng = BatchedADMM(c["models"], c["d_min"], c["K"], rho_admm=1.0, max_iter=1, si_variant=True)
K = c["K"]
XU = [m.initialize_trajectory(np.zeros((3, K)), np.zeros((m.n_u, K))) for m in c["models"]]
X0 = torch.as_tensor(np.stack([x for x, _ in XU])).to(dev); U0 = torch.as_tensor(np.stack([u for _, u in XU])).to(dev)
eng.solve(X0, U0, c["sigma"])
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * 32)()
lib.scvx_debug_phase_cycles(buf, 1)
eng.max_iter = rounds
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = eng.solve(X0, U0, c["sigma"]); b.record(); torch.cuda.synchronize()
it = eng.ws.iters.double()
print(f"N={N}: {a.elapsed_time(b) / rounds:.2f} ms/round; last round IPM iterations mean {it.mean().item():.1f} min {it.min().item():.0f} max {it.max().item():.0f}; "
      f"optimal {float((eng.ws.status == 0).double().mean().item()):.4f}; primal_hist {[round(v, 4) for v in out['primal_hist']]}")
lib.scvx_debug_phase_cycles(buf, 1)
tot = sum(buf)
if tot:
    names = {0: "setup", 1: "pass R rows", 2: "assembly+reduce+term", 3: "cr_factor", 4: "cr_forward<5>", 5: "schur glue", 6: "cr_backward<5>", 7: "dWa",
             8: "pass P rows", 18: "pass P tail+rhs+yb", 16: "cr_forward<1>", 17: "cr_backward<1>", 9: "corrector dW", 12: "pass S rows", 13: "pass S tail", 19: "epilogue"}
    for i in sorted(names, key=lambda i: -buf[i]):
        print(f"  {names[i]:24s} {buf[i]:>12d}  {100.0 * buf[i] / tot:5.1f} %")
