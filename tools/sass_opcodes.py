#!/usr/bin/env python3
"""Static SASS opcode summary of every kernel in libacgpu.so (cuobjdump -sass): instruction count and the memory / async
opcodes that say how a kernel moves data (LDG/STG vector widths, LDS/STS, TLD = texture gathers, RED/ATOM, and the TMA /
bulk-copy family UBLKCP, UTMALDG, LDGSTS - absent from the product kernels, see profiles/r02_summary.md for the A/B).
usage: python tools/sass_opcodes.py [lib] > profiles/rNN_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ahocorasick_b200", "libacgpu.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, data = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        data[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(2)
        data[kern]["_n"] += 1
        base = op.split(".")[0]
        if base in ("LDG", "STG", "LDS", "STS", "TLD", "TEX", "RED", "REDG", "ATOMG", "ATOMS", "ATOM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS",
                    "SYNCS", "LDL", "STL", "SHFL", "VOTE", "VOTEU", "REDUX", "BAR", "MEMBAR", "CCTL"):
            width = next((w for w in ("128", "64", "U16", "U8") if "." + w in op), "")
            data[kern][base + ("." + width if width and base in ("LDG", "STG", "LDS", "STS") else "")] += 1
print("| kernel | SASS instructions | memory / sync opcodes (static counts) |")
print("|---|---|---|")
for k, c in data.items():
    if not k.startswith(("void acgpu::", "acgpu::")):
        continue
    ops = ", ".join("%s %d" % (o, n) for o, n in sorted(c.items()) if o != "_n")
    print("| `%s` | %d | %s |" % (k.replace("void ", "").replace("acgpu::", ""), c["_n"], ops))
