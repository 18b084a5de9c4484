"""Per-phase SM cycles of ipm_kernel (block 0).  Build first with
    SCVX_NVCC_EXTRA=-DSCVX_PHASE_TIMING python -m scvx_b200._build --force
usage: SCVX_NO_FIXED=1 python tools/phase_timing.py [n_agents] [outer_iters]      (the counters live in the generic kernels' translation unit)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from scvx_b200 import _lib
from scvx_b200.batch import BatchedSCvx
from scvx_b200.models.unicycle_model import UnicycleModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scenes = bench.make_scenes(n, 0)
models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in scenes]
eng = BatchedSCvx(models, 100, max_iter=iters)
lib = _lib.load()
buf = (ctypes.c_ulonglong * 32)()
lib.scvx_debug_phase_cycles(buf, 1)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
out = eng.solve(early_exit=False)
t1.record()
torch.cuda.synchronize()
print(f"{iters} outer iterations of {n} agents: {t0.elapsed_time(t1):.2f} ms; mean IPM iterations {out['ipm_iters'].double().mean(dim=1).tolist()}; optimal {(out['status'] == 0).double().mean().item():.6f}")
lib.scvx_debug_phase_cycles(buf, 1)
its = int(out["ipm_iters"][:, 0].sum().item())
names = {0: "setup", 1: "pass R rows (+ pending step)", 2: "assembly+reduce+term", 3: "cr_factor", 4: "cr_forward<5>", 5: "schur glue",
         6: "cr_backward<5>", 7: "dWa", 8: "pass P rows", 18: "pass P reduce+tail+rhs+yb", 16: "cr_forward<1>", 17: "cr_backward<1>",
         9: "corrector dW", 12: "pass S rows", 13: "pass S tail + W update", 19: "epilogue"}
tot = sum(buf)
print(f"agent 0: {its} IPM iterations over {iters} launches; total {tot} cycles; {tot / max(its, 1):.0f} cycles / IPM iteration")
for i in sorted(names, key=lambda i: -buf[i]):
    print(f"  {names[i]:24s} {buf[i]:>10d}  {100.0 * buf[i] / max(tot, 1):5.1f} %   {buf[i] / max(its, 1):9.0f} / it")
