"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
(one cs_frame step) per kernel: python tools/summarize_launches.py launches.csv [traffic.json] > summary.md; the optional
second argument receives the DRAM bytes per launch of the tcgen05 conv family (bench.py's roofline.traffic)."""
import csv, json, re, sys
from collections import OrderedDict

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
per = OrderedDict()
for r in rows:
    k = per.setdefault(r["ID"], {"name": r["Kernel Name"], "ms": 0.0, "bytes": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        k["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
    else:
        k["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def short(name):
    name = name.replace("cs::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
    m = re.match(r"(\w+)(<[^>]*>)?", name)
    base = m.group(1) if m else name
    targs = (m.group(2) or "") if m else ""
    return base.replace("_kernel", "") + targs.replace("(bool)", "").replace("(int)", "")


agg = OrderedDict()
for k in per.values():
    a = agg.setdefault(short(k["name"]), [0, 0.0, 0.0])
    a[0] += 1; a[1] += k["ms"]; a[2] += k["bytes"]
tot = sum(a[1] for a in agg.values())
print(f"{len(per)} launches, {tot:.1f} ms summed (cold-cache, serialised)\n")
print("| kernel | launches | ms | share | dram GB |\n|---|---|---|---|---|")
for name, (n, ms, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% | {by / 1e9:.2f} |")
fam = [(n, ms, by) for name, (n, ms, by) in agg.items() if name.startswith(("conv_tc", "conv7_tc", "conv3s_tc", "wino_"))]
fn, fms, fby = sum(f[0] for f in fam), sum(f[1] for f in fam), sum(f[2] for f in fam)
print(f"\nconv family (conv_tc + conv7_tc + conv3s_tc + Winograd transforms): {fn} launches, {fms:.1f} ms, {fby / 1e9:.1f} GB DRAM traffic = "
      f"{fby / fn / 1e6:.1f} MB per launch.")
if len(sys.argv) > 2:
    json.dump({"dram_bytes_per_launch": fby / fn, "launches": fn, "family_ms": fms, "source": path}, open(sys.argv[2], "w"), indent=1)
