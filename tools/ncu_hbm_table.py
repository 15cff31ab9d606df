"""Summarise `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` captures per kernel:
mean duration, DRAM traffic and bandwidth against the measured copy peak.  usage: python tools/ncu_hbm_table.py a.csv [b.csv ...]"""
import csv, json, os, re, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6547.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
print("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (cold-cache, serialised launches)")
print(f"# HBM peak (MEASURED_PEAKS.json, copy kernel): {peak:.0f} GB/s")
for path in sys.argv[1:]:
    rows = [l for l in open(path) if l.startswith('"')]
    per = OrderedDict()
    for r in csv.DictReader(rows):
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    agg = OrderedDict()
    for d in per.values():
        m = re.search(r"(\w+_kernel)", d["name"])
        a = agg.setdefault(m.group(1) if m else d["name"][:40], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d["dram__bytes_read.sum"]; a[3] += d["dram__bytes_write.sum"]
    print(f"\n## {os.path.basename(path)}")
    print(f"{'kernel':<28}{'n':>4}{'us/launch':>11}{'rd MB':>10}{'wr MB':>10}{'GB/s':>9}{'% peak':>8}")
    for k, (n, ns, rd, wr) in agg.items():
        gbs = (rd + wr) / ns
        print(f"{k:<28}{n:>4}{ns / n / 1e3:>11.1f}{rd / n / 1e6:>10.1f}{wr / n / 1e6:>10.1f}{gbs:>9.0f}{100 * gbs / peak:>7.1f}%")
