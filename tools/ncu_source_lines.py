"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line:
samples and executed warp instructions. Usage: ncu_source_lines.py dump.csv [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, newline="")))
cur_file = None
agg = defaultdict(lambda: [0, 0, ""])
tot_s = tot_i = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "", "Function Name"):
        continue
    try:
        line = int(r[0])
        samples = int(r[6]) if r[6] not in ("-", "") else 0
        inst = int(r[7]) if r[7] not in ("-", "") else 0
    except ValueError:
        continue
    a = agg[(cur_file, line)]
    a[0] += samples
    a[1] += inst
    a[2] = r[1].strip()[:110]
    tot_s += samples
    tot_i += inst
print(f"total samples {tot_s}  total warp inst {tot_i}")
for (f, l), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{l:<4} inst {i:>10} ({100.0*i/max(tot_i,1):5.1f}%)  samples {s:>6} ({100.0*s/max(tot_s,1):5.1f}%)  {src}")
