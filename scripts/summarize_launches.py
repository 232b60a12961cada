"""Summarise an ncu launch list (csv: gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum] per
launch) by kernel name.  python scripts/summarize_launches.py launches.csv [gemm_traffic.json]"""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
launch = collections.OrderedDict()      # ID -> [name, us, read, write]
for r in csv.DictReader(lines):
    key = r["ID"]
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))
    e = launch.setdefault(key, [name, 0.0, 0.0, 0.0])
    v = float(r["Metric Value"].replace(",", "") or 0)
    unit = r.get("Metric Unit", "")
    m = r.get("Metric Name")
    if m == "gpu__time_duration.sum":
        e[1] = v * UNIT.get(unit, 1e-3)
    elif m == "dram__bytes_read.sum":
        e[2] = v * BYTES.get(unit, 1.0)
    elif m == "dram__bytes_write.sum":
        e[3] = v * BYTES.get(unit, 1.0)
agg = collections.OrderedDict()
for name, us, rd, wr in launch.values():
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] += rd
    a[3] += wr
tot = sum(v[1] for v in launch.values())
print(f"{len(launch)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
print(f"{'kernel':84s} {'n':>5s} {'us':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
for n, (c, v, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:84]:84s} {c:5d} {v:10.1f} {100 * v / tot:6.1f}% {rd / 1e6:11.1f} {wr / 1e6:11.1f}")
g = [(c, v, rd, wr) for n, (c, v, rd, wr) in agg.items() if n.startswith("act::gemm_bf16")]
if g:
    n = sum(x[0] for x in g)
    us, rd, wr = sum(x[1] for x in g), sum(x[2] for x in g), sum(x[3] for x in g)
    print(f"\ntcgen05 GEMM kernels: {n} launches, {us:.1f} us = {100 * us / tot:.1f}% of kernel time, DRAM traffic "
          f"{(rd + wr) / 1e6:.1f} MB per step = {(rd + wr) / n / 1e6:.2f} MB per launch")
    if len(sys.argv) > 2:
        json.dump({"launches": n, "dram_bytes_per_step": rd + wr, "dram_bytes_per_launch": (rd + wr) / n,
                   "share_of_kernel_time": us / tot, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, one step"},
                  open(sys.argv[2], "w"), indent=1)
