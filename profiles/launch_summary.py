"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: usage launch_summary.py file.csv"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}.get(r[ui], 1e-6)
    name = r[ki]
    name = name[: name.index("(")] if "(" in name else name
    agg[name[:70]][0] += 1
    agg[name[:70]][1] += v
tot = sum(v[1] for v in agg.values())
print("%-72s %5s %10s %6s" % ("kernel", "n", "ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %5d %10.3f %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("%-72s %5s %10.3f" % ("total (serialised, cold-cache; shares are what matters)", "", tot))
