"""Summarise CONVDBG lines (PCB_CONV_DEBUG=1 stderr of tools/profile_layers.py): per layer shape, the share of block 0's
cycles each warp role spent blocked."""
import collections, sys
agg = collections.OrderedDict()
for line in open(sys.argv[1]):
    if not line.startswith("CONVDBG"):
        continue
    desc, vals = line[8:].split(" | ")
    d = dict(kv.split("=") for kv in vals.split())
    a = agg.setdefault(desc, [0, collections.Counter()])
    a[0] += 1
    for k, v in d.items():
        a[1][k] += float(v)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
for desc, (n, c) in sorted(agg.items(), key=lambda kv: -kv[1][1]["cyc"])[:top]:
    print(desc[:100], n, " ".join(f"{k}={v/n:.2f}" if k not in ("mhz", "cyc") else f"{k}={v/n:.0f}" for k, v in c.items()))
