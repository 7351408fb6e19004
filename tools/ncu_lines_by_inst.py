#!/usr/bin/env python3
"""CUDA source lines ranked by executed warp instructions, from `ncu --page source --print-source cuda,sass --csv`.
usage: ncu_lines_by_inst.py src.csv [top]"""
import csv, sys
def fl(x):
    try: return float(x or 0)
    except ValueError: return 0.0
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, hdr, data = "", None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    data.append((fl(r[hdr.index("Instructions Executed")]), fl(r[hdr.index("# Samples")]), cur, r[0], r[1].strip()[:100]))
data.sort(reverse=True)
tot = sum(d[0] for d in data)
print("total %.1fM" % (tot / 1e6))
for d in data[:top]:
    print("%8.1fM %5.1f%% smp %6d %s:%s  %s" % (d[0] / 1e6, 100 * d[0] / tot, d[1], d[2], d[3], d[4]))
