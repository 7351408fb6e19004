#!/usr/bin/env python3
"""Aggregate an ncu source page (--page source --print-source cuda,sass --csv) by code region of kernel_tier.cuh.
usage: ncu_regions.py src.csv  [file:lo-hi=name ...]"""
import csv, sys
from collections import defaultdict
def fl(x):
    try: return float(x or 0)
    except ValueError: return 0.0
rows = list(csv.reader(open(sys.argv[1])))
spec = []
for a in sys.argv[2:]:
    loc, name = a.split("=")
    f, rng = loc.split(":")
    lo, hi = rng.split("-")
    spec.append((f, int(lo), int(hi), name))
cur, hdr, data = "", None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    data.append((cur, int(r[0]), fl(r[hdr.index("# Samples")]), fl(r[hdr.index("Instructions Executed")]), fl(r[hdr.index("Thread Instructions Executed")])))
def region(f, l):
    for sf, lo, hi, name in spec:
        if f == sf and lo <= l <= hi: return name
    return f
agg = defaultdict(lambda: [0, 0, 0])
for f, l, s, i, t in data:
    a = agg[region(f, l)]; a[0] += s; a[1] += i; a[2] += t
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1
for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
    print('%-34s samples %5.1f%%  inst %5.1f%% (%8.1fM)  thr/inst %4.1f' % (k, 100*a[0]/ts, 100*a[1]/ti, a[1]/1e6, a[2]/max(a[1], 1)))
print('total warp-inst %.1fM' % (ti/1e6))
