"""Exclusive per-source-line instruction counts from `ncu --page source --print-source cuda,sass --csv`.
An inlined SASS instruction is listed under every line of its inline stack; it is attributed here to the
line whose group is smallest (the most specific one).  usage: ncu_exclusive.py dump.csv [top]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1], newline="")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur_file, cur = None, None
groups = defaultdict(list)      # (file,line) -> [(addr, inst, samples)]
src = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name"):
        continue
    if r[0] != "":
        try:
            cur = (cur_file, int(r[0])); src[cur] = r[1].strip()[:100]
        except ValueError:
            cur = None
        continue
    if cur is None or not r[2].startswith("0x"):
        continue
    try:
        groups[cur].append((r[2], int(r[7]), int(r[6]) if r[6] not in ("-", "") else 0, r[3].strip()))
    except ValueError:
        pass
best = {}
for k, lst in groups.items():
    for addr, inst, smp, txt in lst:
        if addr not in best or len(lst) < best[addr][0]:
            best[addr] = (len(lst), k, inst, smp, txt)
agg = defaultdict(lambda: [0, 0])
ops = defaultdict(int)
for addr, (_, k, inst, smp, txt) in best.items():
    agg[k][0] += inst; agg[k][1] += smp
    ops[txt.split()[0].split(".")[0] if not txt.startswith("@") else txt.split()[1].split(".")[0]] += inst
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"exclusive totals: warp inst {ti}  samples {ts}  (distinct SASS {len(best)})")
for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<4} inst {i:>10} ({100.0*i/ti:5.1f}%) samples {100.0*s/max(ts,1):5.1f}%  {src.get(k,'')}")
print("opcodes:", ", ".join(f"{o} {100.0*c/ti:.1f}%" for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:24]))

# optional phase buckets: ncu_exclusive.py dump.csv top "file:lo-hi=name,..."
if len(sys.argv) > 3:
    buckets = []
    for spec in sys.argv[3].split(","):
        rng, name = spec.split("=")
        f, lh = rng.split(":")
        lo, hi = lh.split("-")
        buckets.append((f, int(lo), int(hi), name))
    tot = defaultdict(int)
    for (f, l), (i, s) in agg.items():
        for bf, lo, hi, name in buckets:
            if f == bf and lo <= l <= hi:
                tot[name] += i
                break
        else:
            tot["other:" + f] += i
    for name, i in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"  {name:28s} {i:>11} {100.0*i/ti:5.1f}%  ({i/10000:.0f} per keypoint)")
