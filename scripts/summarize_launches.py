"""Summarise an ncu launch list (csv of gpu__time_duration.sum per launch) by kernel name."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((name, v))
agg = collections.OrderedDict()
for n, v in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in rows)
print(f"{len(rows)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
print(f"{'kernel':90s} {'n':>5s} {'us':>10s} {'share':>7s}")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:90]:90s} {c:5d} {v:10.1f} {100 * v / tot:6.1f}%")
