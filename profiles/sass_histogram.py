#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of the in-tree CUDA library: python profiles/sass_histogram.py > profiles/r02_sass_histogram.txt
(cuobjdump -sass; what proves which instructions the kernels are made of: VABSDIFF4 / REDUX in the search, packed fp32
FFMA2 / FADD2 / FMUL2 and 64/128-bit LDG / STG in the warp and pack kernels; UTMALDG + SYNCS (TMA on an mbarrier) in the
staged variant of the second-generation search only; no tensor ops anywhere)."""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
so = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "mpv-frame-interpolator_b200" / "csrc" / "libhopperrender_cuda.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print("arch lines:", sorted(set(re.findall(r"arch = (sm_\w+)", txt))))
for blk in re.split(r"\n\s*Function : ", txt)[1:]:
    name = subprocess.run(["c++filt", blk.split("\n")[0].strip()], capture_output=True, text=True).stdout.strip()
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", blk):
        op = m.group(1)
        parts = op.split(".")
        key = parts[0]
        if key in ("LDG", "STG", "LD", "ST", "LDS", "STS"):
            key = ".".join(p for p in parts if p in (key, "E", "64", "128", "U8", "U16", "CONSTANT")[:1] or p in ("64", "128", "U8", "U16"))
        elif key in ("FFMA2", "FADD2", "FMUL2", "VABSDIFF4", "REDUX", "ATOM", "RED"):
            key = ".".join(parts[:2]) if len(parts) > 1 and key != "VABSDIFF4" else key
        ops[key] += 1
    total = sum(ops.values())
    if total < 40:
        continue
    print("\n%s\n  %d instructions" % (name[:150], total))
    line = "  "
    for k, v in ops.most_common(28):
        item = "%s %d" % (k, v)
        if len(line) + len(item) > 118:
            print(line)
            line = "  "
        line += item + "  "
    print(line)
    note = ["%s %d" % (k, v) for k, v in sorted(ops.items()) if k.split(".")[0] in ("UTMALDG", "SYNCS", "VABSDIFF4", "REDUX", "CREDUX", "UBLKCP") or k in ("LDG.128", "STG.128")]
    if note:
        print("  of note: " + "  ".join(note))
