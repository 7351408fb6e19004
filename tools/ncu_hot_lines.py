#!/usr/bin/env python3
"""Top CUDA source lines by warp-stall samples from
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur_file, hdr, data = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit():
        continue
    try:
        s = float(r[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    data.append((s, cur_file, r))
tot = sum(d[0] for d in data) or 1
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ie, te = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
data.sort(key=lambda x: -x[0])
print("total samples", int(tot))
for s, f, r in data[:top]:
    stalls = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:3]
    inst = float(r[ie] or 0)
    thr = float(r[te] or 0) / inst if inst else 0
    print("%5.1f%% %-18s:%-4s inst=%-10d thr/inst=%4.1f %-60s | %s" % (
        100 * s / tot, f, r[0], inst, thr, r[1].strip()[:60], ", ".join("%s=%d" % (n[6:], v) for v, n in stalls if v)))
