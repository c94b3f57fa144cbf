"""Per-CUDA-source-line instruction counts of one kernel from an .ncu-rep (needs -lineinfo):
   python tools/ncu_line_mix.py report.ncu-rep kernel_regex [launch_skip] [top_n]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{pat}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 10 and r[0] not in ("", "Line No"):
        try:
            lines.append((cur_file, int(r[0]), r[1], int(r[hdr["Instructions Executed"]] or 0), int(r[hdr["# Samples"]] or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print(f"total warp-instructions {tot}, samples {ts}")
for f, ln, src, n, s in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100 * n / tot:5.1f}% instr {100 * s / ts:5.1f}% stall  {f}:{ln:<4d} {src.strip()[:100]}")
