#!/usr/bin/env python
"""Join an ncu SASS source page (csv) with nvdisasm -g line info: instructions executed and
stall samples per CUDA source line. Usage:
  ncu -i rep --page source --csv --kernel-name regex:NAME > src.csv
  cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > all.sass
  python sass_by_line.py src.csv all.sass MANGLED_SUBSTR [top]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, key = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
# 1. nvdisasm: sequence of (line) per instruction for the function
lines = open(sass).read().split("\n")
infn = False
cur = None
seq = []
for l in lines:
    if l.startswith(".text.") and l.endswith(":"):
        infn = key in l
        continue
    if l.startswith("//---------------------") and infn and ".text." in l and key not in l:
        infn = False
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ie, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
data = []
for r in rows[2:]:
    if len(r) < 10:
        break            # a second instance of the kernel follows: keep the first
    data.append(r)
print("sass instr in profile: %d, in disasm: %d" % (len(data), len(seq)))
agg = defaultdict(lambda: [0, 0, 0])
for i, r in enumerate(data):
    ln = seq[i] if i < len(seq) else None
    a = agg[ln]
    a[0] += int(r[ie]); a[1] += int(r[isamp]); a[2] += 1
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
srcs = {}
def srcline(ln):
    if ln is None: return ""
    f, n = ln
    if f not in srcs:
        import glob
        g = glob.glob("/root/repo/**/" + f, recursive=True)
        srcs[f] = open(g[0]).read().split("\n") if g else []
    return srcs[f][n - 1].strip()[:100] if n - 1 < len(srcs[f]) else ""
print("total warp-instr %d, samples %d" % (tot, ts))
print("--- by instructions executed")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%9d %5.1f%% | smp %5.1f%% | sass %4d | %s | %s" % (a[0], 100 * a[0] / tot, 100 * a[1] / max(ts, 1), a[2], ln, srcline(ln)))
print("--- by stall samples")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%9d %5.1f%% | smp %5.1f%% | sass %4d | %s | %s" % (a[0], 100 * a[0] / tot, 100 * a[1] / max(ts, 1), a[2], ln, srcline(ln)))
