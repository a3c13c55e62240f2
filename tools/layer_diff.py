"""Compare two per-launch tables of tools/layer_table.py, aggregated by (stage, kernel shape): python tools/layer_diff.py old.csv new.csv"""
import csv, collections, re, sys


def load(p):
    agg = collections.OrderedDict()
    for r in csv.reader(open(p)):
        if not r or r[0] == "idx" or r[0].startswith("#"):
            continue
        d = r[6] if len(r) > 6 else ""
        d = re.sub(r" (BN|st|sets|acc)=\d+", "", d)
        d = re.sub(r" tiles=\S+", "", d).replace(" pair", "").strip()
        a = agg.setdefault(d, [0, 0.0])
        a[0] += 1; a[1] += float(r[2])
    return agg


a, b = load(sys.argv[1]), load(sys.argv[2])
keys = list(dict.fromkeys(list(a) + list(b)))
rows = [(b.get(k, [0, 0.0])[1] - a.get(k, [0, 0.0])[1], k) for k in keys]
ta, tb = sum(v[1] for v in a.values()), sum(v[1] for v in b.values())
print(f"total {ta:.3f} -> {tb:.3f} ms")
for dlt, k in sorted(rows, key=lambda x: -abs(x[0]))[:45]:
    va, vb = a.get(k, [0, 0.0]), b.get(k, [0, 0.0])
    print(f"{dlt:+8.3f} ms  {va[1]:7.3f} ({va[0]:2d}x) -> {vb[1]:7.3f} ({vb[0]:2d}x)  {k[:110]}")
