"""Eager lane pipeline vs CUDA-graph replay, over the number of lanes (bench scenes, config 2).
usage: python tools/exp_graph_lanes.py [steps] [lanes ...]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from scvx_b200.batch import PipelinedSCvx
from scvx_b200.models.unicycle_model import UnicycleModel

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lanes_list = [int(x) for x in sys.argv[2:]] or [4, 8, 16]
warm = 3
scenes = bench.make_scenes(bench.N_AGENTS, 0)
models = [UnicycleModel(r_init=o.x_init, r_final=o.x_final, obstacles=[(list(c), r) for c, r in o.obstacles]) for o in scenes]
stream = torch.cuda.current_stream()
for lanes in lanes_list:
    for mode in ("eager", "lanegraph"):
        P = PipelinedSCvx(models, bench.K_NODES, n_lanes=lanes, max_iter=warm + steps + 4, adaptive_mu0=bool(int(os.environ.get('SCVX_ADAPTIVE_MU0', '0')))).start()
        P.run(warm)
        if mode == "lanegraph":
            P.build_lane_graphs()
        elif mode != "eager":
            P.build_graph(steps_per_graph=4 if mode == "graph4" else 1)
            # build_graph may have run one more eager step only if none had run: not the case here
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        if mode == "eager":
            P.run(steps)
        elif mode == "lanegraph":
            P.run_lane_graphs(steps, keep_history=False)
        else:
            P.run_graph(steps, keep_history=True)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        sig = float(sum(st[2].sum().item() for st in P.state))
        print(json.dumps({"lanes": lanes, "mode": mode, "ms_per_step": ms / steps, "agent_it_per_s": bench.N_AGENTS * steps / (ms * 1e-3),
                          "checksum_sigma": sig, "optimal": float((P.status() == 0).double().mean().item()), "launches": P.launches, "mean_ipm_iters": float(P.ipm_iters().double().mean().item())}), flush=True)
