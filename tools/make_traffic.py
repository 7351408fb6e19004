#!/usr/bin/env python3
"""profiles/traffic.json from `ncu --set full` raw CSV dumps of k_tier_mask and k_tier_emit (one 1e9-char haystack of
configs[4]): DRAM bytes per launch of the headline kernels, stamped with a hash of the kernel sources so that bench.py
refuses to quote it for other kernels.
usage: python tools/make_traffic.py <mask_raw.csv> <emit_raw.csv> [chars] [keywords]"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["kernel_tier.cuh", "kernel_mask.cuh", "kernel_pair.cuh", "kernel_emit.cuh"]


def kernel_sources_sha():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, "ahocorasick_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def dram(path):
    rows = list(csv.reader(open(path)))
    hdr, units, r = rows[0], rows[1], rows[2]

    def get(name):
        i = hdr.index(name)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(r[i].replace(",", "")) * scale
    return {"dram_read": get("dram__bytes_read.sum"), "dram_write": get("dram__bytes_write.sum"),
            "ms": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}[units[hdr.index("gpu__time_duration.sum")]]}


if __name__ == "__main__":
    mask, emit = dram(sys.argv[1]), dram(sys.argv[2])
    chars = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000_000
    kws = int(sys.argv[4]) if len(sys.argv) > 4 else 1_000_000
    scan = {"dram_read": chars / 256 * 4, "dram_write": chars / 256 * 4, "note": "row counts in and out (size of its buffers; not captured)"}
    total = sum(k["dram_read"] + k["dram_write"] for k in (mask, emit, scan))
    out = {"workload": "configs[4], one %d-char haystack" % chars, "keywords": kws, "chars": chars, "k_tier_mask": mask, "k_tier_emit": emit,
           "k_row_scan": scan, "source": "ncu --set full: %s, %s" % (os.path.basename(sys.argv[1]), os.path.basename(sys.argv[2])),
           "kernel_sources_sha": kernel_sources_sha(), "kernel_sources": KERNEL_SOURCES, "dram_bytes_per_char": total / chars}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out))
