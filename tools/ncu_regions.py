#!/usr/bin/env python
"""Region-level stall analysis of a kernel from an .ncu-rep source page: regions are delimited by
VOTE instructions.  usage: ncu_regions.py report.ncu-rep"""
import csv, io, subprocess, sys
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, body = rows[1], rows[2:]
si, ie = h.index("# Samples"), h.index("Instructions Executed")
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
alls = sum(int(r[si]) for r in body)
marks = [0] + [i for i, r in enumerate(body) if "VOTE" in r[1]] + [len(body)]
print("total samples", alls)
for a, b in zip(marks[:-1], marks[1:]):
    if b <= a:
        continue
    t = sum(int(body[i][si]) for i in range(a, b))
    agg = {}
    f64 = 0
    for i in range(a, b):
        for k in stalls:
            agg[k] = agg.get(k, 0) + int(body[i][h.index(k)] or 0)
        op = body[i][1].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
        if op.startswith(("DFMA", "DMUL", "DADD", "DSETP")):
            f64 += 1
    ex = max(int(body[i][ie]) for i in range(a, b))
    top = ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(1, t)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:4])
    print("[%4d,%4d) n=%4d f64=%3d samples %6d %5.1f%% maxexec %9d | %s" % (a, b, b - a, f64, t, 100.0 * t / alls, ex, top))
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for i in range(a, b):
        r = body[i]
        st = sorted(((k[6:], int(r[h.index(k)] or 0)) for k in stalls), key=lambda x: -x[1])[:2]
        print(i, r[1].strip()[:70].ljust(70), r[si].rjust(6), st)
