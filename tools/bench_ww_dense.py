#!/usr/bin/env python3
"""WholeWord on text where EVERY word is a keyword (the opposite of configs[3], where 1 word in 90 is): device-resident
throughput of WholeWordMatchSet / Map.  usage: [ACGPU_WW_GEN=2] python tools/bench_ww_dense.py [--chars N]"""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ahocorasick_b200 as ac
import workloads as W
from ahocorasick_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--chars", type=int, default=400_000_000)
a = ap.parse_args()
cfg = W.config(3)
kws = cfg["keywords"]
wc, tg = cfg["word_chars"]
rng = np.random.default_rng(5)
# one block of ~4 M chars of dictionary words separated by single spaces, tiled
words = [kws[i] for i in rng.integers(0, len(kws), 600_000)]
block = np.frombuffer((" ".join(words) + " ").encode("utf-16-le"), dtype=np.uint16)
reps = (a.chars + block.size - 1) // block.size
hay = torch.from_numpy(np.tile(block, reps)[:a.chars].view(np.int16).copy()).cuda()
lib = _lib.lib()
for name, m in (("WholeWordMatchSet", ac.WholeWordMatchSet(kws, True, wc, tg)), ("WholeWordMatchMap", ac.WholeWordMatchMap(kws, list(range(len(kws))), True, wc, tg))):
    tot = C.c_int64(0)
    _lib.check(lib.acgpu_match_device(m.handle, hay.data_ptr(), a.chars, 0, a.chars, None, None, 0, C.byref(tot), None))
    cap = tot.value
    d_pos = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
    d_val = torch.empty(cap, dtype=torch.int32, device="cuda") if m._is_map else None
    d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream()
    def run():
        _lib.check(lib.acgpu_match_device_async(m.handle, hay.data_ptr(), a.chars, 0, a.chars, d_pos.data_ptr(), d_val.data_ptr() if d_val is not None else None, cap, d_tot.data_ptr(), C.c_void_p(st.cuda_stream)))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"matcher": name, "gen": os.environ.get("ACGPU_WW_GEN", "3"), "chars": a.chars, "matches": cap, "matches_per_char": round(cap / a.chars, 3), "ms": round(ms, 3),
                      "ms_per_1e9_chars": round(ms * 1e9 / a.chars, 3), "haystack_GB_per_s": round(2 * a.chars / ms / 1e6, 1)}), flush=True)
