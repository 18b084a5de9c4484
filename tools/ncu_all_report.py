"""Markdown table of roofline fractions for every kernel in an `ncu --set full` report (read HERE with `ncu -i`).

    python tools/ncu_all_report.py gpurun_out/all_r1k.ncu-rep [--all] > profiles/r01k_all_kernels_ncu.md

Per captured launch: duration, DRAM traffic and its fraction of the measured HBM copy peak (MEASURED_PEAKS.json),
EXECUTED fp64 flops (2 x DFMA + DADD + DMUL thread instructions, from the SASS op counters) and their fraction of the
measured DFMA peak (scvx_probe_fp64: 35.0 TFLOP/s), pipe / issue / occupancy figures and the dominant stall reason.
Without --all only the LAST captured launch of every distinct kernel is listed (the earlier ones are set-up iterations).
"""
import csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64_PEAK_TFLOPS = 35.0          # scvx_probe_fp64 DFMA micro-benchmark on the B200 (bench.py measures it live)


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("scvx::", "")


def print_table(recs, show_all, hbm_peak):
    if not show_all:
        last = {}
        for rec in recs:
            last[(rec["name"], rec["grid"], rec["block"])] = rec
        # keep the largest launch of every kernel name (set-up calls use small grids)
        best = {}
        for rec in last.values():
            if rec["name"] not in best or rec["us"] > best[rec["name"]]["us"]:
                best[rec["name"]] = rec
        recs = sorted(best.values(), key=lambda x: -x["us"])
    print(f"| kernel | grid x block | regs | dyn smem | duration | DRAM traffic | HBM GB/s (frac of {hbm_peak:.0f}) | executed fp64 TFLOP/s (frac of {FP64_PEAK_TFLOPS:.0f}) | fp64 pipe active | issue active | warps active | bound | top stalls (warps per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for x in recs:
        fh, ff = x["hbm_gbs"] / hbm_peak, x["tflops"] / FP64_PEAK_TFLOPS
        if x["us"] < 15.0 and max(fh, ff) < 0.25:
            bound = "launch latency (short)"
        elif max(fh, ff, x["fp64_pipe"] / 100.0) < 0.15:
            bound = "latency / launch"
        elif fh >= max(ff, x["fp64_pipe"] / 100.0):
            bound = "HBM"
        else:
            bound = "FP64 pipe"
        print(f"| `{x['name']}` | {x['grid']} x {x['block']} | {x['regs']} | {x['smem_kb']:.1f} KB | {x['us']:.1f} us | {x['dram_mb']:.2f} MB | "
              f"{x['hbm_gbs']:.0f} ({100 * fh:.1f} %) | {x['tflops']:.2f} ({100 * ff:.1f} %) | {x['fp64_pipe']:.1f} % | {x['issue']:.1f} % | "
              f"{x['warps']:.1f} % | {bound} | {x['stall']} |")


def long_csv_report(path, show_all, hbm_peak, SCALE):
    """`ncu --metrics ... --csv --log-file x.csv` output: one line per (launch, metric)."""
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    h = rows[0]
    ci = {x: i for i, x in enumerate(h)}
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(int(r[ci["ID"]]), {"name": short(r[ci["Kernel Name"]]), "block": r[ci["Block Size"]], "grid": r[ci["Grid Size"]], "m": {}})
        try:
            v = float(r[ci["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        d["m"][r[ci["Metric Name"]]] = v * SCALE.get(r[ci["Metric Unit"]], 1.0)
    prod = lambda t: eval("*".join(t.strip("()").split(",")))      # noqa: E731  "(128, 1, 1)" -> 128
    recs = []
    for _id in sorted(launches):
        d = launches[_id]; m = d["m"]
        t = m.get("gpu__time_duration.sum", 0.0)
        dram = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        flops = 2.0 * m.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0) + \
            m.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0) + m.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
        recs.append({"name": d["name"], "grid": prod(d["grid"]), "block": prod(d["block"]), "regs": int(m.get("launch__registers_per_thread", 0)),
                     "smem_kb": m.get("launch__shared_mem_per_block_dynamic", 0.0) / 1e3, "us": t * 1e6, "dram_mb": dram / 1e6,
                     "hbm_gbs": dram / t / 1e9 if t else 0.0, "tflops": flops / t / 1e12 if t else 0.0,
                     "fp64_pipe": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0.0),
                     "issue": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0),
                     "warps": m.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0), "stall": "-"})
    print_table(recs, show_all, hbm_peak)


def main():
    rep = sys.argv[1]
    show_all = "--all" in sys.argv
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm_peak = 6650.0
    SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "s": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "usecond": 1e-6,
             "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0, "Kbyte/block": 1e3, "Mbyte/block": 1e6, "byte/block": 1.0, "Ghz": 1e9,
             "Mhz": 1e6, "cycle/nsecond": 1e9, "cycle/usecond": 1e6, "cycle/second": 1.0}
    if rep.endswith(".csv"):
        return long_csv_report(rep, show_all, hbm_peak, SCALE)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    col = {x: i for i, x in enumerate(h)}

    def val(r, key, default=0.0):
        i = col.get(key)
        if i is None or r[i] == "":
            return default
        return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)

    stalls = [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    recs = []
    for r in data:
        name = short(r[col["Kernel Name"]])
        t = val(r, "gpu__time_duration.sum")
        dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        hz = val(r, "sm__cycles_elapsed.max.per_second")
        flop_cyc = 2.0 * val(r, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") + \
            val(r, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") + \
            val(r, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed")
        tflops = flop_cyc * hz / 1e12
        st = sorted(((val(r, k), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for k in stalls
                     if "selected" not in k), reverse=True)
        recs.append({
            "name": name, "grid": int(val(r, "launch__grid_size")), "block": int(val(r, "launch__block_size")),
            "regs": int(val(r, "launch__registers_per_thread")), "smem_kb": val(r, "launch__shared_mem_per_block_dynamic") / 1e3,
            "us": t * 1e6, "dram_mb": dram / 1e6, "hbm_gbs": dram / t / 1e9 if t else 0.0, "tflops": tflops,
            "fp64_pipe": val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "stall": ", ".join(f"{n} {v:.2f}" for v, n in st[:2]),
        })
    print_table(recs, show_all, hbm_peak)


if __name__ == "__main__":
    main()
