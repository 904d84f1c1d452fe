"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per kernel launches, total / mean device time and share.  Output goes under profiles/."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        rows.append((name, v, r["Grid Size"], r["Block Size"]))
agg = collections.OrderedDict()
for name, v, g, b in rows:
    a = agg.setdefault(name, [0, 0.0, g, b])
    a[0] += 1
    a[1] += v
total = sum(a[1] for a in agg.values()) or 1.0
print("%-44s %8s %12s %12s %7s  %s" % ("kernel", "launches", "total_us", "mean_us", "share", "grid x block (first)"))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s %8d %12.1f %12.2f %6.1f%%  %s x %s" % (name[:44], a[0], a[1], a[1] / a[0], 100 * a[1] / total, a[2], a[3]))
print("total_us %.1f over %d launches" % (total, len(rows)))
