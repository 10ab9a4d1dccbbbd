#!/usr/bin/env python
"""Instruction mix of one kernel from an .ncu-rep captured with --import-source on: executed warp
instructions and stall samples per SASS opcode, fp64 share, shared-memory wavefronts.
usage: ncu_mix.py report.ncu-rep [--loop LO HI]   (LO/HI: row range to list with counts)"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if "Source" in r)
data = rows[rows.index(hdr) + 1:]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iw = hdr.index("L1 Wavefronts Shared")
tot = sum(int(r[ie]) for r in data); ts = sum(int(r[isamp]) for r in data)
c, s = Counter(), Counter()
for r in data:
    t = r[ia].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[ie]); s[op] += int(r[isamp])
f64 = sum(c[o] for o in ("DFMA", "DADD", "DMUL", "DSETP"))
print("warp instructions %d  samples %d  fp64 %.1f%%  smem wavefronts %d" % (
    tot, ts, 100.0 * f64 / tot, sum(int(r[iw]) for r in data)))
for op, n in c.most_common(28):
    print("  %-8s %6.2f%% inst  %6.2f%% samples" % (op, 100.0 * n / tot, 100.0 * s[op] / ts))
if "--loop" in sys.argv:
    k = sys.argv.index("--loop")
    lo, hi = int(sys.argv[k + 1]), int(sys.argv[k + 2])
    for i, r in enumerate(data):
        if lo <= i < hi:
            print(i, r[ie], r[isamp], r[ia])
