"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
   python tools/launch_summary.py launches.csv [last_n_launches]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
agg = collections.OrderedDict()
tot = 0.0
for row in rows:
    name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1][:70]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}.get(u, 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"{len(rows)} launches, {tot:.3f} ms total")
for k, (c, v) in agg.items():
    print(f"{k:70s} n={c:4d} total={v:10.3f} ms avg={v / c:9.3f} ms share={100 * v / tot:5.1f}%")
