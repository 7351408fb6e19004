#!/usr/bin/env python3
"""Opcode histogram (weighted by executed warp instructions) of an ncu `--page source --print-source sass --csv` dump.
usage: ncu_sass_hist.py sass.csv [top]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = defaultdict(lambda: [0.0, 0.0, 0])
tot = 0.0
for r in rows[2:]:
    if len(r) <= iex: continue
    s = r[isrc].strip()
    toks = s.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "ATOM", "RED")) else op.split(".")[0]
    n = float(r[iex] or 0); sm = float(r[ismp] or 0)
    a = agg[op]; a[0] += n; a[1] += sm; a[2] += 1
    tot += n
ts = sum(a[1] for a in agg.values()) or 1
print("total warp-inst %.1fM, static %d" % (tot / 1e6, sum(a[2] for a in agg.values())))
for op, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%-14s exec %8.1fM %5.1f%%  samples %5.1f%%  static %d" % (op, a[0] / 1e6, 100 * a[0] / tot, 100 * a[1] / ts, a[2]))
