#!/usr/bin/env python3
"""Experiment: independent match() calls of the headline workload issued round-robin on S streams, so k_tier_emit of one
haystack can run next to k_tier_mask of the next (the two kernels stall on different things).  Prints ms per haystack.
  [ACGPU_LIB=variant.so] python tools/exp_overlap.py --streams 2 --haystacks 6"""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ahocorasick_b200 as ac
import workloads as W
from ahocorasick_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=2)
ap.add_argument("--haystacks", type=int, default=6)
ap.add_argument("--chars", type=int, default=1_000_000_000)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.lib()
kws = W.config(4)["keywords"]
m = ac.AhoCorasickSet(kws, True)
n = args.chars
hays = [W.make_haystack_torch(W.HaystackSpec("lower", 2005 + i, kws), n, device=dev) for i in range(args.haystacks)]
torch.cuda.synchronize()
tot = C.c_int64(0)
counts = []
for h in hays:
    _lib.check(lib.acgpu_match_device(m.handle, h.data_ptr(), n, 0, n, None, None, 0, C.byref(tot), None))
    counts.append(tot.value)
cap = max(counts)
streams = [torch.cuda.Stream() for _ in range(args.streams)]
outs = [torch.empty((cap, 2), dtype=torch.int32, device=dev) for _ in range(args.streams)]
d_tot = torch.zeros(args.haystacks, dtype=torch.int64, device=dev)

def step():
    for j, h in enumerate(hays):
        s = streams[j % args.streams]
        _lib.check(lib.acgpu_match_device_async(m.handle, h.data_ptr(), n, 0, n, outs[j % args.streams].data_ptr(), None, cap,
                                                d_tot[j:].data_ptr(), C.c_void_p(s.cuda_stream)))

main = torch.cuda.current_stream()
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(main)
for s in streams:
    s.wait_event(e0)
for _ in range(args.steps):
    step()
for s in streams:
    ev = torch.cuda.Event()
    ev.record(s)
    main.wait_event(ev)
e1.record(main)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps / args.haystacks
assert d_tot.tolist() == counts
alg = 2 * n + 8 * sum(counts) / len(counts)
print(json.dumps({"lib": os.environ.get("ACGPU_LIB", "product"), "streams": args.streams, "ms_per_haystack": ms,
                  "roofline_frac_6457": alg / ms / 1e6 / 6457.4}))
