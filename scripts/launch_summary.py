"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python scripts/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, body = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in body:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0][:72]
    v = float(r[mv].replace(",", ""))
    v *= {"us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(r[mu], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':72s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>6s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n:8d} {t:10.3f} {t / n:9.3f} {100 * t / tot:5.1f}%")
