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

# ---- three ways to run the SAME recorded work (outer iterations 3.. of all agents), in interior-point iteration times -----------
# (a) synchronous: one launch over all agents per step, longest-first by the previous solve, barrier after every step
# (b) lanes (PipelinedSCvx): 4 lanes of 256 agents, a barrier per lane per step, lanes share the 296 slots (event simulation)
# (c) per-agent chains: every agent runs its own steps back to back on one slot (a persistent block per agent; 296 at a time)
steps = list(range(3, a.shape[0]))
sync = sum(makespan(np.argsort(-a[it - 1], kind="stable"), a[it]) for it in steps)


def lanes_time(n_lanes):
    n = a.shape[1]
    per = (n + n_lanes - 1) // n_lanes
    lanes = [np.arange(i, min(i + per, n)) for i in range(0, n, per)]
    # event simulation: each lane submits its step's blocks (longest first by its previous solve) when its previous step is done;
    # free slots take the pending block whose lane submitted earliest (stream order within a lane, FIFO across lanes)
    slot_free = [0.0] * SLOTS
    heapq.heapify(slot_free)
    lane_ready = [0.0] * len(lanes)
    lane_step = [0] * len(lanes)
    finish = 0.0
    pending = True
    while pending:
        pending = False
        # pick the lane whose next step can start earliest
        cand = [(lane_ready[l], l) for l in range(len(lanes)) if lane_step[l] < len(steps)]
        if not cand:
            break
        pending = True
        t0, l = min(cand)
        it = steps[lane_step[l]]
        idx = lanes[l]
        order = idx[np.argsort(-a[it - 1][idx], kind="stable")]
        done = t0
        for i in order:
            s = max(heapq.heappop(slot_free), t0)
            e = s + a[it][i]
            heapq.heappush(slot_free, e)
            done = max(done, e)
        lane_ready[l] = done
        lane_step[l] += 1
        finish = max(finish, done)
    return finish


def chains_time():
    tot = a[steps].sum(axis=0)                      # every agent's own chain over the recorded steps
    return makespan(np.argsort(-tot, kind="stable"), tot)


work = a[steps].sum() / SLOTS
print(f"\n{len(steps)} recorded steps, total work / 296 slots = {work:.0f} iteration times ({work / len(steps):.1f} per step)")
print(f"(a) synchronous steps, longest first          : {sync:.0f}  ({sync / len(steps):.1f} per step, {sync / work:.2f} x total work)")
for nl in (2, 4, 8, 16, 64):
    t = lanes_time(nl)
    print(f"(b) {nl:3d} lanes, barrier per lane per step       : {t:.0f}  ({t / len(steps):.1f} per step, {t / work:.2f} x total work)")
t = chains_time()
print(f"(c) per-agent chains (persistent block per agent): {t:.0f}  ({t / len(steps):.1f} per step, {t / work:.2f} x total work)")
