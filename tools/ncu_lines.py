#!/usr/bin/env python3
"""Top CUDA source lines by executed warp instructions from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: ncu_lines.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples"); continue
    if hdr is None or not r[0].isdigit() or len(r) != len(hdr): continue
    try: ie = float(r[ii] or 0); sm = float(r[si] or 0)
    except ValueError: continue
    k = (cur, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(k, [0, 0]); a[0] += ie; a[1] += sm
tot = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print("total warp-inst %.1fM samples %d" % (tot / 1e6, ts))
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / ts, k[0], k[1], k[2]))
