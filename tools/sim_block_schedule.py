"""List-scheduling simulation behind DESIGN 4.9 / 4.10: 1024 sub-problem blocks on 296 SM slots (148 SMs x 2 resident blocks),
block time proportional to its interior-point iteration count (recorded on the bench: profiles/r01j_ipm_iters_1024x12.npy,
12 outer iterations x 1024 agents).  Prints, per launch order, the makespan relative to the lower bound
max(total work / slots, longest block), and that bound in iteration times.

    python tools/sim_block_schedule.py [profiles/r01j_ipm_iters_1024x12.npy]
"""
import heapq, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01j_ipm_iters_1024x12.npy")
a = np.load(path).astype(float)
SLOTS = 296


def makespan(order, t):
    h = [0.0] * SLOTS
    heapq.heapify(h)
    for i in order:
        heapq.heappush(h, heapq.heappop(h) + t[i])
    return max(h)


res = {}
for it in range(3, a.shape[0]):
    t = a[it]
    bound = max(t.sum() / SLOTS, t.max())
    preds = {"index order": None, "longest first by the previous solve (built, 4.9)": a[it - 1],
             "longest first by max of the last two solves": np.maximum(a[it - 1], a[it - 2]),
             "longest first, exact (oracle)": t}
    for k, p in preds.items():
        order = np.arange(t.size) if p is None else np.argsort(-p, kind="stable")
        res.setdefault(k, []).append(makespan(order, t) / bound)
    res.setdefault("_bound", []).append((t.sum() / SLOTS, t.max()))
print("lower bound per outer iteration (total work / 296 slots, longest block), in interior-point iteration times:")
print("  " + "  ".join(f"({w:.1f}, {m:.0f})" for w, m in res.pop("_bound")))
for k, v in res.items():
    print(f"{k:55s} makespan / bound: mean {np.mean(v):.3f}   per iteration " + " ".join(f"{x:.2f}" for x in v))
