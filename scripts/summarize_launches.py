"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel (share of the step)."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("capdec::", "").replace("(anonymous namespace)::", "")
    rows.append((name, r.get("Grid Size", ""), r.get("Block Size", ""), ns))
total = sum(r[3] for r in rows)
agg = OrderedDict()
for name, grid, block, ns in rows:
    k = (name, grid)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
byname = {}
for (name, grid), (n, ns) in agg.items():
    b = byname.setdefault(name, [0, 0.0])
    b[0] += n
    b[1] += ns
print("total %.3f ms over %d launches" % (total / 1e6, len(rows)))
print("\n== by kernel ==")
for name, (n, ns) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us  %5.1f%%  x%-4d %s" % (ns / 1e3, 100 * ns / total, n, name))
print("\n== by kernel and grid (top 40) ==")
for (name, grid), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%8.1f us  %5.1f%%  x%-4d avg %7.1f us  grid %-16s %s" % (ns / 1e3, 100 * ns / total, n, ns / n / 1e3, grid, name))
