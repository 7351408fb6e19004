"""match(Readable) throughput against the device-block size: acgpu_stream_* fed from pinned host memory in fixed blocks of
2^16 .. 2^25 chars, and with the adaptive policy of the host mirrors (2^16 doubling to 2^24).  Values-only replay is not
included (a listener call per match is the caller's cost).  One JSON line per matcher.  GPU needed.
usage: python tools/bench_stream_sweep.py [--chars 200000000]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ahocorasick_b200 as ac  # noqa: E402
import workloads as W  # noqa: E402
from ahocorasick_b200 import _lib  # noqa: E402


def run(lib, m, host, ne, sizes):
    res = _lib.Result()
    sh = C.c_uint64(0)
    _lib.check(lib.acgpu_stream_begin(m.handle, C.byref(sh)))
    if m._is_map and m._family != _lib.SHORTEST:   # what every host mirror does: ReadableMatchListener sees values only
        _lib.check(lib.acgpu_stream_set_values_only(sh.value, 1))
    lo, k, n_rec = 0, 0, 0
    t0 = time.perf_counter()
    while lo < ne:
        blk = sizes(k)
        _lib.check(lib.acgpu_stream_feed(sh.value, host.data_ptr() + 2 * lo, min(blk, ne - lo), C.byref(res)))
        n_rec += res.n
        lib.acgpu_free_result(C.byref(res))
        lo += blk
        k += 1
    _lib.check(lib.acgpu_stream_end(sh.value, C.byref(res)))
    n_rec += res.n
    lib.acgpu_free_result(C.byref(res))
    return time.perf_counter() - t0, n_rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chars", type=int, default=200_000_000)
    ap.add_argument("--configs", default="3,1,2")
    a = ap.parse_args()
    only = {int(x) for x in a.configs.split(",")}
    torch.cuda.set_device(0)
    lib = _lib.lib()
    for idx, name, make in ((3, "WholeWordMatchMap", lambda c, k, v: ac.WholeWordMatchMap(k, v, True, *c["word_chars"])),
                            (1, "AhoCorasickMap(ci)", lambda c, k, v: ac.AhoCorasickMap(k, v, False)),
                            (2, "LongestMatchMap", lambda c, k, v: ac.LongestMatchMap(k, v, True))):
        if idx not in only:
            continue
        cfg = W.config(idx)
        kws = cfg["keywords"]
        m = make(cfg, kws, list(range(len(kws))))
        ne = a.chars
        hay = W.make_haystack_torch(cfg["spec"], ne, device="cuda")
        host = torch.empty(ne, dtype=torch.int16, pin_memory=True)
        host.copy_(hay)
        del hay
        torch.cuda.synchronize()
        row = {"config": idx, "matcher": name, "chars": ne, "GB_per_s": {}}
        run(lib, m, host, min(ne, 1 << 24), lambda k: 1 << 22)  # warm-up: contexts, pinned result blocks
        for sh in (16, 18, 20, 22, 23, 24, 25):
            best = min(run(lib, m, host, ne if sh >= 20 else ne // 8, lambda k, s=sh: 1 << s)[0] / (1 if sh >= 20 else 0.125)
                       for _ in range(2))
            row["GB_per_s"]["2^%d" % sh] = round(2 * ne / best / 1e9, 2)
        dt, n_rec = min(run(lib, m, host, ne, lambda k: min(1 << 24, 1 << (16 + k))) for _ in range(2))
        row["GB_per_s"]["adaptive 2^16..2^24"] = round(2 * ne / dt / 1e9, 2)
        row["records"] = n_rec
        print(json.dumps(row), flush=True)
        m.close()
        del host


if __name__ == "__main__":
    main()
