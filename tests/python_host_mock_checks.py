"""Run by tests/test_python_host_mock.py in a subprocess with ACGPU_LIB pointing at libacgpu_mock_oracle.so (the C ABI
answered by the oracle - test infrastructure).  Exercises the PYTHON host mirror (ahocorasick_b200/matchers.py,
streaming.py) on the CPU: constructors, Iterable zipping, listener replay with the early-stop quirks Q1/Q2, Readable
fills / device blocks / quirk Q4 - against the literal oracle's listener-call sequences."""
import io
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ahocorasick_b200 as ac  # noqa: E402
from ahocorasick_b200 import _lib  # noqa: E402
from ahocorasick_b200.streaming import match_readable  # noqa: E402
from golden_cases import FAMILIES  # noqa: E402
from oracle import oracle as ora  # noqa: E402

assert b"mock-oracle" in _lib.lib().acgpu_version(), "these checks must run against the mocked C ABI"

SETS = {"ahocorasick": ac.AhoCorasickSet, "longest": ac.LongestMatchSet, "shortest": ac.ShortestMatchSet,
        "wholeword": ac.WholeWordMatchSet, "wholewordlongest": ac.WholeWordLongestMatchSet}
MAPS = {"ahocorasick": ac.AhoCorasickMap, "longest": ac.LongestMatchMap, "shortest": ac.ShortestMatchMap,
        "wholeword": ac.WholeWordMatchMap, "wholewordlongest": ac.WholeWordLongestMatchMap}


class Collect:
    def __init__(self, stop_after=0):
        self.calls, self.stop_after = [], stop_after

    def match(self, *args):
        self.calls.append(args[1:] if len(args) >= 3 else args[0])
        return not (self.stop_after and len(self.calls) >= self.stop_after)


def oracle_calls(m, hay, **kw):
    return [(int(r["start"]), int(r["end"]), int(r["value"])) for r in m.match(hay, **kw)]


n_checks = 0
rng = random.Random(2026)
for family in FAMILIES:
    for it in range(60):
        alpha = rng.choice(["ab", "abc", "abAB"])
        kws = [("".join(rng.choice(alpha) for _ in range(rng.randint(1, 5)))) for _ in range(rng.randint(1, 10))]
        if rng.random() < 0.3:
            kws.insert(rng.randint(0, len(kws)), rng.choice([None, ""]))
        cs = rng.random() < 0.5
        hay = "".join(rng.choice(alpha + "  ,") for _ in range(rng.randint(0, 200)))
        n_values = rng.choice([len(kws), max(0, len(kws) - 1), len(kws) + 2])      # zip stops at the shorter Iterable
        values = ["v%d" % i for i in range(n_values)]
        om_set = ora.Matcher(family, kws, case_sensitive=cs)
        om_map = ora.Matcher(family, kws, n_values=n_values, case_sensitive=cs)
        gs, gm = SETS[family](iter(kws), cs), MAPS[family](iter(kws), iter(values), cs)
        full = oracle_calls(om_set, hay)
        for stop in [0] + list(range(1, min(len(full), 6) + 2)):
            c = Collect(stop)
            gs.match(hay, c)
            assert [(a, b) for a, b in c.calls] == [(s, e) for s, e, _ in oracle_calls(om_set, hay, stop_after=stop)], (family, kws, hay, stop)
            c = Collect(stop)
            gm.match(hay, c)
            assert [(a, b, v) for a, b, v in c.calls] == [(s, e, values[v]) for s, e, v in oracle_calls(om_map, hay, stop_after=stop)], (family, stop)
            n_checks += 2
        # Readable: adaptive blocks and fixed ones; values only; Shortest delivers fill-boundary matches twice (Q4)
        for stop in (0, 1, 3):
            want = [values[v] for _, _, v in oracle_calls(om_map, hay, readable=True, stop_after=stop)]
            for block in (0, 4096, 1 << 16):
                c = Collect(stop)
                match_readable(gm, io.StringIO(hay), c.match, block_chars=block)
                assert c.calls == want, (family, kws, hay, stop, block)
                n_checks += 1

# long streams: many charBufferSize fills, several device blocks, Q4 duplicates in the Shortest family
for family in FAMILIES:
    kws = sorted({"".join(rng.choice("abc") for _ in range(rng.randint(1, 6))) for _ in range(40)})
    hay = "".join(rng.choice("abc  ") for _ in range(300_000))
    values = list(range(len(kws)))
    om = ora.Matcher(family, kws, n_values=len(kws))
    want = [v for _, _, v in oracle_calls(om, hay, readable=True)]
    gm = MAPS[family](kws, values, True)
    for block in (0, 3 * 4096):
        got = []
        match_readable(gm, io.StringIO(hay), lambda v: got.append(v) or True, block_chars=block)
        assert got == want, (family, block)
        n_checks += 1
    if family == "shortest":
        plain = [v for _, _, v in oracle_calls(om, hay)]
        assert len(want) > len(plain), "the stream should hold fill-boundary duplicates (Q4)"

# constructor behaviour
try:
    ac.WholeWordMatchSet(["fine", " as if "], True)
    raise SystemExit("IllegalArgumentException expected")
except ac.IllegalArgumentException as e:
    assert "as if contains non-word characters." in str(e)
m = ac.WholeWordMatchSet(["key", "a=b"], True, ["_", "="], [False, True])
c = Collect()
m.match("key_a=b key=a=b _key_", c)
assert c.calls == [(0, 3), (4, 7), (17, 20)] and not m.getWordChars()[ord("_")] and m.getWordChars()[ord("=")]
for bad in (None,):
    try:
        ac.AhoCorasickSet(["a"], True).match(bad, lambda *a: True)
        raise SystemExit("TypeError expected")
    except TypeError:
        pass
# constructors from a flattened goto trie (acgpu_create, trie_desc.flatten_trie = the Java-side builder's loops)
from ahocorasick_b200 import trie_desc
_FAM = {"ahocorasick": 0, "longest": 1, "shortest": 2, "wholeword": 3, "wholewordlongest": 4}
for family in MAPS:
    kws = ["he", "she", None, "hers", "his", "he", "", "She"]
    hay = "ushers and She said his is hers; he!"
    for cs in (True, False):
        nv = len(kws) - 1  # a shorter values Iterable: the zip drops the last keyword
        want = oracle_calls(ora.Matcher(family, kws, n_values=nv, case_sensitive=cs), hay)
        gm = MAPS[family].from_trie(trie_desc.flatten_trie(_FAM[family], kws, list(range(nv)), cs), list(range(nv)))
        c = Collect()
        gm.match(hay, c)
        assert c.calls == want, (family, cs, c.calls, want)
        n_checks += 1
print("python host mirror ok: %d checks" % n_checks)
