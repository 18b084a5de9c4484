"""BASELINE config 1 end to end (shipped unicycle, K=50, <= 30 outer iterations): the GPU loop, the oracle loop, and -- because
the sub-problems are degenerate LPs whose minimisers are not unique, so the two loops part ways after a few iterations -- the
oracle's exact LP re-solved ON THE GPU LOOP'S OWN PARAMETERS at every iteration.

usage: python tools/config1_table.py [out.md]      (needs cuda:0; imports oracle/: a checker tool, not a product path)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch


def run(K=50, max_iter=30):
    from oracle import foh as ofoh, models as omodels, scvx as oscvx, subproblem as ospb
    from scvx_b200.batch import BatchedSCvx
    from scvx_b200.models.unicycle_model import UnicycleModel
    om = omodels.unicycle()
    eng = BatchedSCvx([UnicycleModel()], K, max_iter=max_iter)
    b = eng.batch
    dev = b.device
    X, U = b.initial_trajectories()
    sig = torch.ones(1, dtype=torch.float64, device=dev); tr = torch.full((1,), 100.0, dtype=torch.float64, device=dev)
    act = torch.ones(1, dtype=torch.int32, device=dev); met = torch.zeros((1, 6), dtype=torch.float64, device=dev)
    F = ofoh.OracleFOH(om, K)
    rows = []
    for it in range(max_iter):
        if not int(act.item()):
            break
        Xr, Ur, sr, trr = X[0].cpu().numpy().copy(), U[0].cpu().numpy().copy(), float(sig.item()), float(tr.item())
        eng.iterate(X, U, sig, tr, act, met)
        torch.cuda.synchronize()
        mats_gpu = tuple(m[0].cpu().numpy() for m in eng.mats)
        mats_tight = F.calculate_discretization(Xr, Ur, sr, tol="tight")
        foh_err = max(np.abs(g - w).max() / max(np.abs(w).max(), 1e-300) for g, w in zip(mats_gpu, mats_tight))
        p = ospb.Params(om, K, mats_gpu, Xr, Ur, sr, trr)
        r = ospb.solve(p)                                              # exact LP (HiGHS) on the GPU loop's parameters
        Xn, Un, sn = eng.ws.X[0].cpu().numpy(), eng.ws.U[0].cpu().numpy(), float(eng.ws.sigma[0].item())
        e = ospb.evaluate(p, Xn, Un, sn)
        m = met[0].cpu().numpy()
        rows.append({"iter": it, "obj_gpu": e["obj"], "obj_lp": r["obj"], "rel": (e["obj"] - r["obj"]) / abs(r["obj"]), "viol": e["viol"],
                     "nu_norm": float(m[0]), "slack": float(m[1]), "sigma": sn, "sigma_ref": sr, "tr": trr, "foh_err": foh_err,
                     "ipm_iters": int(eng.ws.iters[0].item()), "status": int(eng.ws.status[0].item())})
    _, _, sig_o, rec_o = oscvx.scvx_solve(om, K, max_iter=max_iter)
    return rows, rec_o, sig_o


def render(rows, rec_o):
    L = ["| outer it | GPU sub-problem value | exact LP on the same parameters | rel. diff | hard-constraint viol. | FOH rel. err vs tight oracle | nu-norm GPU loop | nu-norm oracle loop | sigma GPU loop | sigma oracle loop | IPM its |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
    for i, r in enumerate(rows):
        o = rec_o[i] if i < len(rec_o) else None
        L.append(f"| {r['iter']} | {r['obj_gpu']:.6f} | {r['obj_lp']:.6f} | {r['rel']:+.1e} | {r['viol']:.1e} | {r['foh_err']:.1e} | "
                 f"{r['nu_norm']:.3e} | {(o['nu_norm'] if o else float('nan')):.3e} | {r['sigma']:.4f} | {(o['sigma'] if o else float('nan')):.4f} | {r['ipm_iters']} |")
    return "\n".join(L)


if __name__ == "__main__":
    rows, rec_o, sig_o = run()
    txt = render(rows, rec_o)
    last, lo = rows[-1], rec_o[-1]
    summary = (f"\nAfter {len(rows)} outer iterations (GPU loop) / {len(rec_o)} (oracle loop): final sub-problem value {last['obj_gpu']:.4f} vs "
               f"{lo['obj']:.4f}, nu-norm {last['nu_norm']:.3e} vs {lo['nu_norm']:.3e}, obstacle slack {last['slack']:.3e} vs {lo['slack_norm']:.3e}, "
               f"sigma {last['sigma']:.4f} vs {lo['sigma']:.4f}; worst per-iteration value difference on identical parameters "
               f"{max(abs(r['rel']) for r in rows):.1e}, worst hard-constraint violation {max(r['viol'] for r in rows):.1e}, "
               f"worst FOH error {max(r['foh_err'] for r in rows if r['sigma_ref'] >= 1e-6):.1e} over the iterations with sigma_ref >= 1e-6 (the shipped "
               f"loop drives sigma to ~1e-12, where B_bar, C_bar, S_bar are ~1e-14 in magnitude and the ORACLE's absolute tolerance of 1e-14 "
               f"no longer resolves them: the relative figures of those rows measure the oracle, not the kernel).\n")
    out = "# Config 1 end to end: shipped unicycle, K=50 (`python tools/config1_table.py`)\n\n" + txt + "\n" + summary
    print(out)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(out)
