#!/usr/bin/env python3
"""Kernel-resident and end-to-end throughput of BASELINE.json configs[0..3] (and of workloads.config(5), the "real
dictionary" workload) on one B200 (the headline config[4] is bench.py).  One JSON line per (config, matcher).  Inputs are generated on the device; timing = CUDA events on the
launch stream, haystacks larger than L2.

  python tools/bench_configs.py [--scale S] [--configs 0,1,2,3]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import ahocorasick_b200 as ac  # noqa: E402
import workloads as W  # noqa: E402
from ahocorasick_b200 import _lib  # noqa: E402


def matchers_for(idx, cfg):
    kws = cfg["keywords"]
    vals = list(range(len(kws)))
    if idx == 0:
        return [("AhoCorasickSet", ac.AhoCorasickSet(kws, True))]
    if idx == 1:
        return [("AhoCorasickMap(ci)", ac.AhoCorasickMap(kws, vals, False))]
    if idx == 2:
        return [("LongestMatchMap", ac.LongestMatchMap(kws, vals, True)), ("ShortestMatchSet", ac.ShortestMatchSet(kws, True))]
    if idx == 3:
        wc, tg = cfg["word_chars"]
        # the same text with a phrase dictionary (every tenth entry is a two-word keyword)
        phrases = kws + [kws[i] + " " + kws[i + 1] for i in range(0, len(kws) - 1, 10)]
        return [("WholeWordMatchSet", ac.WholeWordMatchSet(kws, True, wc, tg)),
                ("WholeWordMatchMap", ac.WholeWordMatchMap(kws, vals, True, wc, tg)),
                ("WholeWordLongestMatchSet(+phrases)", ac.WholeWordLongestMatchSet(phrases, True, wc, tg))]
    if idx == 5:  # the "real dictionary" workload (workloads.config(5)): outside the narrow-alphabet envelope -> kernel_wide.cuh
        return [("AhoCorasickSet(english-like, 53 symbols, len 1-24)", ac.AhoCorasickSet(kws, True)),
                ("AhoCorasickMap(english-like)", ac.AhoCorasickMap(kws, vals, True))]
    raise ValueError(idx)


def measure(configs, scale=1.0, steps=5, warmup=3, e2e_chars=100_000_000):
    """Yields one dict per (config, matcher)."""
    args = argparse.Namespace(scale=scale, steps=steps, warmup=warmup, e2e_chars=e2e_chars)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = _lib.lib()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for idx in configs:
        cfg = W.config(idx)
        n = max(1 << 20, int(cfg["n"] * args.scale))
        n = min(n, 2_000_000_000)
        hay = W.make_haystack_torch(cfg["spec"], n, device=dev)
        torch.cuda.synchronize()
        for name, m in matchers_for(idx, cfg):
            is_map = m._is_map
            stream = torch.cuda.current_stream()
            sp = C.c_void_p(stream.cuda_stream)
            tot = C.c_int64(0)
            _lib.check(lib.acgpu_match_device(m.handle, hay.data_ptr(), n, 0, n, None, None, 0, C.byref(tot), sp))
            cap = max(tot.value, 1)
            d_pos = torch.empty((cap, 2), dtype=torch.int32, device=dev)
            d_val = torch.empty(cap, dtype=torch.int32, device=dev) if is_map else None
            d_tot = torch.zeros(1, dtype=torch.int64, device=dev)

            def step():
                _lib.check(lib.acgpu_match_device_async(m.handle, hay.data_ptr(), n, 0, n, d_pos.data_ptr(),
                                                        d_val.data_ptr() if is_map else None, cap, d_tot.data_ptr(), sp))
            for _ in range(args.warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            assert int(d_tot.item()) == tot.value
            rec = 12 if is_map else 8
            alg = 2 * n + rec * tot.value
            # end to end through the host-buffer call
            ne = min(n, args.e2e_chars)
            host = torch.empty(ne, dtype=torch.int16, pin_memory=True)
            host.copy_(hay[:ne])
            res = _lib.Result()
            ts = []
            res_n = 0
            for it in range(3):
                t0 = time.perf_counter()
                _lib.check(lib.acgpu_match_utf16(m.handle, host.data_ptr(), ne, C.byref(res)))
                ts.append(time.perf_counter() - t0)
                res_n = res.n
                lib.acgpu_free_result(C.byref(res))
            stream_gbps = None
            if is_map:
                # match(Readable): the same host chars fed through acgpu_stream_* in 4 Mi-char blocks (values only)
                # straight through the C ABI (the Python mirror would copy every record block into numpy arrays)
                best = None
                blk = 1 << 22
                for it in range(3):
                    t0 = time.perf_counter()
                    sh = C.c_uint64(0)
                    _lib.check(lib.acgpu_stream_begin(m.handle, C.byref(sh)))
                    if m._family != _lib.SHORTEST:   # what every host mirror does: ReadableMatchListener sees values only
                        _lib.check(lib.acgpu_stream_set_values_only(sh.value, 1))
                    n_rec = 0
                    for lo in range(0, ne, blk):
                        _lib.check(lib.acgpu_stream_feed(sh.value, host.data_ptr() + 2 * lo, min(blk, ne - lo), C.byref(res)))
                        n_rec += res.n
                        lib.acgpu_free_result(C.byref(res))
                    _lib.check(lib.acgpu_stream_end(sh.value, C.byref(res)))
                    n_rec += res.n
                    lib.acgpu_free_result(C.byref(res))
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                stream_gbps = 2 * ne / best / 1e9
            yield {
                "config": idx, "matcher": name, "chars": n, "keywords": len(cfg["keywords"]), "matches": tot.value,
                "ms": ms, "haystack_GB_per_s": 2 * n / ms / 1e6, "matches_per_s": tot.value / ms * 1e3,
                "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak, "frac": alg / ms / 1e6 / peak,
                             "algorithmic_bytes": alg},
                "e2e_GB_per_s": 2 * ne / min(ts[1:]) / 1e9, "e2e_chars": ne, "e2e_d2h_bytes": int(res_n) * rec,
                "readable_stream_GB_per_s": stream_gbps,
                "launches_per_match": lib.acgpu_launches_per_match(m.handle), "info": m.info()}
            del d_pos, d_val
            m.close()
        del hay
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="0,1,2,3")
    ap.add_argument("--scale", type=float, default=1.0, help="shrinks the haystack (dictionary stays full size)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--e2e-chars", type=int, default=100_000_000)
    args = ap.parse_args()
    for d in measure([int(x) for x in args.configs.split(",")], args.scale, args.steps, args.warmup, args.e2e_chars):
        print(json.dumps(d), flush=True)


if __name__ == "__main__":
    main()
