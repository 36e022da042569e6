"""Summarise an ncu source-page CSV: top SASS instructions by stall samples, and sample
share by opcode.  usage: ncu -i X.ncu-rep --page source --csv > f.csv; python scripts/ncu_hot.py f.csv [N]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
body = [r for r in rows[2:] if len(r) >= len(hdr) - 5]
ks = hdr.index("# Samples")
src = hdr.index("Source")
ie = hdr.index("Instructions Executed")
tot = sum(float(r[ks] or 0) for r in body)
print("total samples", tot, "sass lines", len(body))
byop = Counter()
cnt = Counter()
for r in body:
    op = r[src].split()[0] if r[src].split() else "?"
    if op.startswith("@"):
        op = r[src].split()[1]
    byop[op.split(".")[0]] += float(r[ks] or 0)
    cnt[op.split(".")[0]] += float(r[ie] or 0)
print("--- samples by opcode (share of samples | share of executed warp-instructions)")
ti = sum(cnt.values())
for op, v in byop.most_common(14):
    print(f"{100*v/tot:6.2f}%  {100*cnt[op]/ti:6.2f}%  {op}")
print("--- top instructions")
for i, r in sorted(enumerate(body), key=lambda t: -float(t[1][ks] or 0))[:N]:
    print(f"{100*float(r[ks] or 0)/tot:6.2f}%  line {i:5d}  {r[src].strip()[:110]}")
