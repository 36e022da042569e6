"""Attribute the stall samples / executed instructions of one kernel to SOURCE FUNCTIONS.

ncu's CSV export has per-SASS-instruction metrics but no source correlation; nvdisasm -g gives
file:line per SASS offset for the same binary.  Joined by instruction offset.

usage:
  cuobjdump -xelf all libfsgpu.so; nvdisasm -g -c fsgpu_elements.sm_100a.cubin > el.sass
  ncu -i X.ncu-rep --page source --csv > src.csv
  python scripts/ncu_by_function.py src.csv el.sass <mangled-kernel-substring> [more substrings]
"""
import csv
import re
import sys
from collections import Counter, defaultdict

src_csv, sass, *subs = [a for a in sys.argv[1:] if not a.startswith("--")]

# --- nvdisasm listing: offsets -> (file, line) inside the wanted kernel section
sect = None
cur = ("?", 0)
loc = {}
for ln in open(sass, errors="replace"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        name = m.group(1)
        sect = name if all(s in name for s in subs) else None
        continue
    if sect is None:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        loc[int(m.group(1), 16)] = (cur, m.group(2))

# --- function table per file (definitions start in column 0)
ftab = {}


def func_of(path, line):
    if path not in ftab:
        t = []
        try:
            for i, s in enumerate(open(path, errors="replace"), 1):
                if re.match(r"^(FS_HD|__device__|__global__|template|static|inline|int |void |EmitRuns|EmitScatter)", s):
                    m = re.search(r"([A-Za-z_]\w*)\s*\(", s)
                    if m and m.group(1) not in ("__launch_bounds__", "defined", "if"):
                        t.append((i, m.group(1)))
                    elif s.startswith("template"):
                        t.append((i, None))  # name on the next line
                elif t and t[-1][1] is None:
                    m = re.search(r"([A-Za-z_]\w*)\s*\(", s)
                    if m and m.group(1) != "__launch_bounds__":
                        t[-1] = (t[-1][0], m.group(1))
        except OSError:
            pass
        ftab[path] = t
    name = "?"
    for i, n in ftab[path]:
        if i <= line and n:
            name = n
        if i > line:
            break
    return name


rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ks, ie = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
isrc = hdr.index("Source")
body = [r for r in rows[2:] if len(r) > ie and r[ia].startswith("0x")]
base = int(body[0][ia], 16)
samp, inst, fp64 = Counter(), Counter(), Counter()
bad = 0
for r in body:
    off = int(r[ia], 16) - base
    if off not in loc:
        bad += 1
        continue
    (path, line), text = loc[off]
    op_csv = r[isrc].split()
    key = (path.split("/")[-1], func_of(path, line))
    s, n = float(r[ks] or 0), float(r[ie] or 0)
    samp[key] += s
    inst[key] += n
    op = text.split()[1] if text.startswith("@") else text.split()[0]
    if op.split(".")[0] in ("DFMA", "DMUL", "DADD"):
        fp64[key] += n
ts, ti = sum(samp.values()), sum(inst.values())
print(f"kernel section match: {len(loc)} instructions; unmatched csv rows {bad}; samples {ts:.0f}; warp-instructions {ti:.0f}")
print(f"{'samples%':>9} {'instr%':>8} {'fp64 share':>10}  function")
for k, v in samp.most_common(30):
    print(f"{100 * v / ts:8.2f}% {100 * inst[k] / ti:7.2f}% {100 * fp64[k] / max(inst[k], 1):9.1f}%  {k[0]}:{k[1]}")

# per-line table for the kernel's own file (member functions are not in the function table)
if "--lines" in sys.argv:
    pass
byline_s, byline_i = Counter(), Counter()
for r in body:
    off = int(r[ia], 16) - base
    if off in loc:
        (path, line), _ = loc[off]
        if path.endswith("fsgpu_elements.cu") or path.endswith("fsgpu_tile.cu"):
            byline_s[line] += float(r[ks] or 0)
            byline_i[line] += float(r[ie] or 0)
print("--- fsgpu_elements.cu by source line (samples% instr%)")
for line in sorted(byline_s):
    if byline_s[line] / ts > 0.004 or byline_i[line] / ti > 0.004:
        print(f"  line {line:5d}: {100 * byline_s[line] / ts:6.2f}% {100 * byline_i[line] / ti:6.2f}%")

# optional: --ranges=NAME:a-b,NAME:c-d  sums the kernel file's lines by range; everything attributed to
# other files (inlined math, intrinsics) is reported as "callees"
rng = [a for a in sys.argv[1:] if a.startswith("--ranges=")]
if rng:
    spec = [(x.split(":")[0], *map(int, x.split(":")[1].split("-"))) for x in rng[0][9:].split(",")]
    acc_s, acc_i = Counter(), Counter()
    for line in byline_s:
        name = next((n for n, a, b in spec if a <= line <= b), "other-lines")
        acc_s[name] += byline_s[line]
        acc_i[name] += byline_i[line]
    acc_s["callees"] = ts - sum(byline_s.values())
    acc_i["callees"] = ti - sum(byline_i.values())
    print("--- by line range (samples% instr%)")
    for n in acc_s:
        print(f"  {n:>12}: {100 * acc_s[n] / ts:6.2f}% {100 * acc_i[n] / ti:6.2f}%")
