"""Instruction mix (by SASS opcode) of an `ncu --page source --csv` dump of ONE kernel:
   python tools/ncu_sass_mix.py src.csv [top_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
agg, samp = collections.Counter(), collections.Counter()
for r in data:
    s = re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip())
    op = s.split()[0].split(".")[0]
    agg[op] += int(r[col["Instructions Executed"]] or 0)
    samp[op] += int(r[col["# Samples"]] or 0)
tot, ts = sum(agg.values()), max(1, sum(samp.values()))
print(f"{len(data)} SASS instructions, {tot} warp-instructions executed, {ts} samples")
for op, n in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:12s} {n:12d} {100 * n / tot:5.1f}%   stall samples {100 * samp[op] / ts:5.1f}%")
