"""Golden fingerprints of the flattened dictionary tables (acgpu_build_fingerprint) for a few seeded dictionaries.

A changed fingerprint means the device tables' layout or content changed: the CUDA kernels read these tables, so the GPU
parity suite must be re-run before the new fingerprints are committed (`python tests/golden/make_builder_fingerprints.py`).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

OUT = os.path.join(HERE, "builder_fingerprints.json")
FAMILIES = ["ahocorasick", "longest", "shortest", "wholeword", "wholewordlongest"]


def dictionaries():
    rng = np.random.default_rng(20261017)

    def words(n, alpha, lo, hi):
        return ["".join(rng.choice(list(alpha), size=int(rng.integers(lo, hi + 1)))) for _ in range(n)]

    return {
        "az_3_12": words(4000, "abcdefghijklmnopqrstuvwxyz", 3, 12),
        "ab_dense": words(1500, "ab", 1, 16),
        "mixed_case_digits": words(1200, "ABCdef0123", 1, 9) + [None, ""],
        "wide_alphabet": [chr(c) for c in range(0x30, 0x800)] + words(100, "αβγЖж", 2, 5),
        "empty": [],
    }


def compute():
    from test_host_cpu import _fingerprint
    out = {}
    for name, kws in dictionaries().items():
        for fam, fname in enumerate(FAMILIES):
            for is_map in (False, True):
                for cs in (True, False):
                    key = "%s/%s/%s/%s" % (name, fname, "map" if is_map else "set", "cs" if cs else "ci")
                    try:
                        out[key] = "%016x" % _fingerprint(fam, kws, len(kws) if is_map else -1, cs)
                    except Exception as e:  # WholeWord refuses keywords with inner non-word chars
                        out[key] = type(e).__name__
    return out


if __name__ == "__main__":
    json.dump(compute(), open(OUT, "w"), indent=0, sort_keys=True)
    print("wrote", OUT)
