import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import workloads as W
import ahocorasick_b200 as ac
c = W.config(4)
kws = c["keywords"]
hay = W.make_haystack(c["spec"], 4_000_000)
r2 = ac.AhoCorasickSet(kws, True).match_records(hay)
os.environ["ACGPU_FORCE_GEN1"] = "1"
r1 = ac.AhoCorasickSet(kws, True).match_records(hay)
print(len(r1), len(r2))
s1 = set(zip(r1.start.tolist(), r1.end.tolist()))
s2 = set(zip(r2.start.tolist(), r2.end.tolist()))
extra = sorted(s2 - s1); missing = sorted(s1 - s2)
print("extra", len(extra), "missing", len(missing))
kwset = set(kws)
txt = hay.astype('<u2').tobytes().decode('utf-16-le')
for s, e in extra[:25]:
    print(s, e, e - s, repr(txt[s:e]), txt[s:e] in kwset, "pos%8192", (e - 1) % 8192, "ctx", repr(txt[max(0, e - 14):e + 2]))
import collections
print(collections.Counter(e - s for s, e in extra))
print(collections.Counter(((e - 1) % 8192) % 8 for s, e in extra))
