"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel.
usage: python tools/launch_summary.py launches.csv"""
import collections, csv, re, sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if r[ui] == "ns":
        v /= 1000.0
    name = r[ki].replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"[<(].*", "", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
print(f"{'kernel':40s} {'launches':>8s} {'us':>12s} {'share':>7s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:40]:40s} {n:8d} {v:12.1f} {100 * v / tot:6.1f}%")
