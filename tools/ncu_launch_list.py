"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file x.csv` launch list by (kernel, grid x block): launches, total
and mean duration, share of all captured kernel time.  usage: python tools/ncu_launch_list.py x.csv [top=24]"""
import collections, csv, re, sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("scvx::", "")[:72]


rows = [r for r in csv.reader(open(sys.argv[1], newline="")) if r and r[0].isdigit()]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
agg = collections.OrderedDict()
for r in rows:
    key = (short(r[4]), r[8], r[7])
    d = agg.setdefault(key, [0, 0.0])
    d[0] += 1; d[1] += float(r[14]) / 1e3      # ns -> us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} ms of kernel time\n")
print("| kernel | grid x block | launches | total | per launch | share |")
print("|---|---|---|---|---|---|")
for (name, grid, block), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"| `{name}` | {grid} x {block} | {n} | {us / 1e3:.3f} ms | {us / n:.1f} us | {100.0 * us / tot:.1f} % |")
