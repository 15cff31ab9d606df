"""Per-layer A/B of two or more PCB_PROFILE_DUMP csv files (tools/profile_layers.py --out): rows keyed by the layer's
shape (plan fields dropped), one ms column per file.  usage: python tools/layers_ab.py a.csv b.csv [c.csv ...]"""
import collections, re, sys


def load(f):
    rows = collections.OrderedDict()
    for line in open(f):
        parts = line.strip().split(",")
        key = ",".join(parts[:-2]); t = float(parts[-2]); fl = float(parts[-1])
        key = re.sub(r",(ntile|mt|tiles|grid|ast|bst|bres|sbox|pair|mc|iss)=[^,]*", "", key)
        a = rows.setdefault(key, [0, 0.0, 0.0]); a[0] += 1; a[1] += t; a[2] += fl
    return rows


tabs = [load(f) for f in sys.argv[1:]]
print(f"{'layer':76s} cnt " + " ".join(f"{'ms['+str(i)+']':>8s}" for i in range(len(tabs))) + "   TF/s[0] -> TF/s[-1]")
for k, (n, t, f) in sorted(tabs[0].items(), key=lambda kv: -kv[1][1]):
    ms = [tb.get(k, [0, float('nan'), 0])[1] for tb in tabs]
    print(f"{k:76s} {n:3d} " + " ".join(f"{m:8.3f}" for m in ms) + f"   {f/ms[0]/1e9:6.0f} -> {f/ms[-1]/1e9:6.0f}")
print(f"{'total':76s}     " + " ".join(f"{sum(v[1] for v in tb.values()):8.3f}" for tb in tabs))
