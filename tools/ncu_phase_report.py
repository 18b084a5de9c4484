"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line and by kernel phase.

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_phase_report.py src.csv [solver.cu]
Phases are found from marker comments in solver.cu, so the report follows the file as it changes.
"""
import csv, sys, re, collections

def load(path):
    rows = list(csv.reader(open(path, newline="")))
    out = []  # (file, line, samples, insts, stall dict)
    cur_file = None; hdr = None; cur_line = None
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; continue
        if r[0] != "": cur_line = int(r[0]); continue
        if hdr is None or len(r) < len(hdr) or r[2] in ("...", ""): continue
        d = dict(zip(hdr[2:], r[2:]))
        def f(k):
            try: return float(d.get(k, "0") or 0)
            except ValueError: return 0.0
        st = {k: f(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        out.append((cur_file, cur_line, f("# Samples"), f("Instructions Executed"), st, d.get("Source", "")))
    return out

def phases(src):
    L = open(src).read().split("\n")
    marks = []
    pats = [("cr_factor", r"void cr_factor\("), ("cr_forward", r"void cr_forward\("), ("cr_backward", r"void cr_backward\("),
            ("schur_solve", r"void schur_solve\("), ("assemble_rows", r"void assemble_rows\("), ("kernel_setup", r"^ipm_kernel\("),
            ("passR_rows", r"-+ RESIDUAL PASS"), ("passR_assembly_term", r"second half of the assembly"),
            ("factor_predictor_glue", r"factor \+ predictor solve"), ("passPS_rows", r"-+ PASS P / PASS S"),
            ("pass_epilogues", r"---- pass epilogues"), ("epilogue", r"---- epilogue: outputs"), ("host", r"constexpr size_t SMEM_LIMIT"),
            ("helpers", r"^#include \"common.cuh\""), ("chol/apply", r"void chol_inverse\(")]
    for name, p in pats:
        for i, l in enumerate(L):
            if re.search(p, l): marks.append((i + 1, name)); break
    marks.sort()
    def ph(line):
        name = "?"
        for ln, n in marks:
            if line >= ln: name = n
        return name
    return ph

if __name__ == "__main__":
    data = load(sys.argv[1])
    src = sys.argv[2] if len(sys.argv) > 2 else "dynamic-programming-multiagent-trajectory-optimiziation_b200/csrc/solver_kernel.cuh"
    kfile = src.split("/")[-1]
    ph = phases(src)
    tot_s = sum(d[2] for d in data); tot_i = sum(d[3] for d in data)
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0, collections.Counter()])
    for f, line, s, i, st, _ in data:
        key = ph(line) if f == kfile else f
        a = agg[key]; a[0] += s; a[1] += i; a[2] += 1
        for k, v in st.items(): a[3][k] += v
    print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.3e}, SASS instructions {len(data)}")
    print("| phase | stall samples | instructions | SASS | top stall reasons |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        top = ", ".join(f"{n[6:]} {100*v/max(a[0],1):.0f}%" for n, v in a[3].most_common(4))
        print(f"| {k} | {100*a[0]/tot_s:.1f} % | {100*a[1]/tot_i:.1f} % | {a[2]} | {top} |")
    byline = collections.defaultdict(float)
    for f, line, s, i, st, _ in data: byline[(f, line)] += s
    print("\ntop lines:")
    srcl = open(src).read().split("\n")
    for (f, line), s in sorted(byline.items(), key=lambda kv: -kv[1])[:14]:
        txt = srcl[line - 1].strip()[:110] if f == kfile else ""
        print(f"- {100*s/tot_s:.1f}% {f}:{line} `{txt}`")
