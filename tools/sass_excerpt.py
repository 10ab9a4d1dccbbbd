#!/usr/bin/env python
"""SASS evidence for profiles/: for each hot kernel of libbart_b200.so the resource usage, the
mnemonic histogram and the listing lines around the instructions that identify the design (bulk-async
copy UBLKCP + mbarrier SYNCS, 256-bit loads LDG.E.ENL2.256, fp64 tensor-core DMMA, DFMA density,
warp shuffles / votes of the scan kernel).
usage: sass_excerpt.py > profiles/r02_sass_excerpts.txt"""
import os, re, subprocess, sys
from collections import Counter
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bart_b200", "libbart_b200.so")
KERNELS = [("eclipse_column_kernel<4,1,5,false,3,false> (W12 headline)", "_ZN4bart21eclipse_column_kernelILi4ELi1ELi5ELb0ELi3ELb0EEE"),
           ("eclipse_scan_kernel<4,1,5,3,false> (latency kernel: lanes <-> layers, warp scan)", "_ZN4bart19eclipse_scan_kernelILi4ELi1ELi5ELi3ELb0EEE"),
           ("eclipse_slot_kernel<4,1,5,3,false> (small batches, run-time-count fallback)", "_ZN4bart19eclipse_slot_kernelILi4ELi1ELi5ELi3ELb0EEE"),
           ("transit_mma_kernel<4,1,false,false> (transit geometry, DMMA)", "_ZN4bart18transit_mma_kernelILi4ELi1ELb0ELb0EEE"),
           ("extinction_kernel<4> (stand-alone opacity lookup)", "_ZN4bart17extinction_kernelILi4EEE"),
           ("accumulate_kernel<128> (opacity-grid builder)", "_ZN4bart17accumulate_kernelILi128EEE")]
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines()
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
for title, prefix in KERNELS:
    blk = next((b for b in blocks if b.startswith(prefix)), None)
    print("=" * 100)
    print(title)
    for i, ln in enumerate(res):
        if prefix in ln and i + 1 < len(res):
            print("  " + res[i + 1].strip())
    if blk is None:
        print("  (not found)")
        continue
    lines = [l for l in blk.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", re.sub(r"\s+/\*[0-9a-f]{4}\*/\s+", "", l)).split()[0].rstrip(";") for l in lines)
    print("  instructions: %d; " % len(lines) + ", ".join("%s %d" % kv for kv in ops.most_common(14)))
    marks = ("UBLKCP", "SYNCS", "LDG.E.ENL2.256", "DMMA", "MUFU.RCP64H", "VOTE", "REDUX", "SHFL")
    shown = 0
    for i, l in enumerate(lines):
        if any(m in l for m in marks) and shown < 6:
            shown += 1
            for k in lines[max(0, i - 2):i + 3]:
                print("    " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", k).rstrip())
            print("    ...")
    print("  counts: " + ", ".join("%s %d" % (m, sum(m in l for l in lines)) for m in marks))
