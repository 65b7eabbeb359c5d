#!/usr/bin/env python
"""Condense an `ncu --page source --csv` export: runs of SASS instructions with the same execution count."""
import csv
import sys

r = list(csv.reader(open(sys.argv[1])))
hdr = r[1]
rows = r[2:]
ia = hdr.index("Instructions Executed")
isrc = hdr.index("Source")
isamp = hdr.index("# Samples")
tot = sum(int(x[ia]) for x in rows)
tots = sum(int(x[isamp]) for x in rows)
print("total", tot / 1e6, "M warp instructions,", len(rows), "SASS lines")
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
prev = None
start = 0
acc = 0
sacc = 0
n = 0
for i, x in enumerate(rows + [None]):
    c = int(x[ia]) if x is not None else -1
    if prev is None or x is None or abs(c - prev) > 0.02 * max(prev, 1):
        if prev is not None and acc / tot > thresh:
            ops = " ".join(rows[k][isrc].split()[0] for k in range(start, min(start + 12, i)))
            print(f"{start:5d}-{i - 1:5d} n={n:4d} each={prev / 1e6:8.2f}M sum={acc / 1e6:8.1f}M {100 * acc / tot:5.1f}% stall {100 * sacc / tots:5.1f}%  {ops[:110]}")
        prev = c
        start = i
        acc = 0
        sacc = 0
        n = 0
    if x is not None:
        acc += c
        sacc += int(x[isamp])
        n += 1
