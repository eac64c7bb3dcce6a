"""Aggregate an ncu source page (ncu -i rep --page source --csv --print-source cuda,sass) per CUDA source line: share of executed
instructions and of stall samples. Usage: python tools/ncu_lines.py page.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and cur:
        try:
            ln = int(r[0])
            ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
            a = agg[(cur, ln)]
            a[0] += int(r[ie]); a[1] += int(r[sm]); a[2] += 1; a[3] = r[1][:120]
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print("total inst", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]}:{k[1]:5d} inst {100 * v[0] / tot:5.1f}% samp {100 * v[1] / ts:5.1f}% nsass {v[2]:5d} | {v[3]}")
