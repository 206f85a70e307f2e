#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) of an `ncu --page source --csv` export: python scripts/ncu_hot.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
body = rows[2:]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[si]) for r in body)
toti = sum(int(r[ii]) for r in body)
print(f"{len(body)} SASS instructions, {tot} samples, {toti} warp instructions executed")
order = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:n]
for i in sorted(order):
    r = body[i]
    print(f"{i:5d} {100 * int(r[si]) / tot:5.1f}%  exec {int(r[ii]):>10d}  {r[1].strip()}")
