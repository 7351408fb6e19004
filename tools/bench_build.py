"""Dictionary-construction throughput - host only, no GPU needed.

Times the two host stages of a constructor for the BASELINE dictionaries: packing the Iterable into the C ABI's arrays
(ahocorasick_b200.matchers._pack_keywords) and the flattening inside libacgpu.so (acgpu_build_fingerprint runs exactly
the build of acgpu_create_from_keywords, minus the upload), serial vs sharded trie insert.  One JSON line per case.
usage: python tools/bench_build.py [--reps 3]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import workloads as W  # noqa: E402
from ahocorasick_b200 import _lib  # noqa: E402
from ahocorasick_b200.matchers import _pack_keywords  # noqa: E402

CASES = [(0, 0, True, False), (1, 0, False, True), (2, 1, True, True), (2, 2, True, False), (3, 3, True, False), (4, 0, True, False)]
NAMES = ["AhoCorasick", "LongestMatch", "ShortestMatch", "WholeWordMatch", "WholeWordLongestMatch"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    L = _lib.lib()
    for cfg, fam, cs, is_map in CASES:
        c = W.config(cfg)
        kws = c["keywords"]
        wc_table = None
        if c.get("word_chars") is not None:   # (chars, toggles) of the WholeWord constructor
            import numpy as np
            from ahocorasick_b200 import WordCharacters
            wc_table = np.ascontiguousarray(WordCharacters.generateWordCharsFlags(*c["word_chars"]).astype(np.uint8))
        t0 = time.perf_counter()
        chars, offsets, is_null, n = _pack_keywords(kws)
        pack_s = time.perf_counter() - t0
        row = {"config": cfg, "matcher": NAMES[fam] + ("Map" if is_map else "Set"), "keywords": n, "chars": int(offsets[-1]),
               "pack_ms": round(pack_s * 1e3, 1), "cores": os.cpu_count()}
        fps = {}
        for mode in ("serial", "sharded"):
            os.environ["ACGPU_BUILDER"] = mode
            best = None
            for _ in range(a.reps):
                fp, secs = C.c_uint64(0), C.c_double(0)
                wcp = wc_table.ctypes.data if wc_table is not None else None
                rc = L.acgpu_build_fingerprint(fam, chars.ctypes.data, offsets.ctypes.data, is_null.ctypes.data, n,
                                               n if is_map else -1, 1 if cs else 0, wcp, C.byref(fp), C.byref(secs))
                dt = secs.value
                assert rc == 0, L.acgpu_last_error()
                best = dt if best is None else min(best, dt)
            fps[mode] = fp.value
            row["build_ms_" + mode] = round(best * 1e3, 1)   # the flattening alone (the fingerprint is not timed)
        row["identical_tables"] = fps["serial"] == fps["sharded"]
        row["keywords_per_s_sharded"] = round(n / (row["build_ms_sharded"] / 1e3))
        print(json.dumps(row), flush=True)
    os.environ.pop("ACGPU_BUILDER", None)


if __name__ == "__main__":
    main()
