"""Top stall sites of an `ncu --page source --csv` dump:  python tools/ncu_source_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
# find header row
his = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "Line")]
sect = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = his[sect]
end = his[sect + 1] if sect + 1 < len(his) else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
print("sections at", his, "first cols", [rows[h][:2] for h in his])
samp = col["# Samples"]
src = col["Source"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(r[samp] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
ranked = sorted(enumerate(data), key=lambda t: -int(t[1][samp] or 0))[:n]
w = col.get("L1 Wavefronts Shared")
wi = col.get("L1 Wavefronts Shared Ideal")
for idx, r in sorted(ranked):
    print(f"{idx:5d} {int(r[samp] or 0):7d} {100 * int(r[samp] or 0) / tot:5.1f}%  wf={r[w] if w else ''}/{r[wi] if wi else ''}  {r[src][:110]}")
