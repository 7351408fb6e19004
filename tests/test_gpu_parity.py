"""GPU parity: the CUDA path (through the C ABI, via the host-side API mirror) must deliver the same
ordered listener calls as the CPU oracle (oracle/ac_oracle.c) on the same inputs.  Bit-exact: these are
integer positions and value indices."""
import io
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ac_spec as spec  # noqa: E402
from golden_cases import FAMILIES, ILLEGAL, LITERAL_CASES  # noqa: E402
from oracle import oracle as ora  # noqa: E402

import ahocorasick_b200 as ac  # noqa: E402
import workloads as W  # noqa: E402

SETS = {"ahocorasick": ac.AhoCorasickSet, "longest": ac.LongestMatchSet, "shortest": ac.ShortestMatchSet,
        "wholeword": ac.WholeWordMatchSet, "wholewordlongest": ac.WholeWordLongestMatchSet}
MAPS = {"ahocorasick": ac.AhoCorasickMap, "longest": ac.LongestMatchMap, "shortest": ac.ShortestMatchMap,
        "wholeword": ac.WholeWordMatchMap, "wholewordlongest": ac.WholeWordLongestMatchMap}


class Collect:
    """Listener that records every call and answers False on its stop_after-th call."""

    def __init__(self, stop_after=0):
        self.calls = []
        self.stop_after = stop_after

    def match(self, *args):
        self.calls.append(args[1:] if len(args) >= 3 else args)
        return not (self.stop_after and len(self.calls) >= self.stop_after)


def oracle_stream(m, hay, **kw):
    return [(int(r["start"]), int(r["end"]), int(r["value"])) for r in m.match(hay, **kw)]


def gpu_set_stream(s, hay, stop_after=0):
    c = Collect(stop_after)
    s.match(hay, c)
    return [(a, b) for a, b in c.calls]


def gpu_map_stream(m, hay, stop_after=0):
    c = Collect(stop_after)
    m.match(hay, c)
    return [(a, b, v) for a, b, v in c.calls]


@pytest.mark.parametrize("name", sorted(LITERAL_CASES))
@pytest.mark.parametrize("family", FAMILIES)
def test_literal_cases(name, family):
    hay, kws, expect = LITERAL_CASES[name]
    if expect.get(family) == ILLEGAL:
        with pytest.raises(ac.IllegalArgumentException, match="contains non-word characters"):
            SETS[family](kws, True)
        with pytest.raises(ac.IllegalArgumentException):
            MAPS[family](kws, kws, True)
        return
    if family == "wholeword" and family not in expect:
        # keywords with inner non-word chars would throw; the reference test classes expect that too
        try:
            om = ora.Matcher(family, kws)
        except ora.OracleError:
            with pytest.raises(ac.IllegalArgumentException):
                SETS[family](kws, True)
            return
    om = ora.Matcher(family, kws, n_values=len(kws))
    want = oracle_stream(om, hay)
    if isinstance(expect.get(family), list):
        assert [(s, e) for s, e, _ in want] == expect[family]
    assert gpu_set_stream(SETS[family](kws, True), hay) == [(s, e) for s, e, _ in want]
    values = ["v%d" % i for i in range(len(kws))]
    got = gpu_map_stream(MAPS[family](kws, values, True), hay)
    assert got == [(s, e, values[v]) for s, e, v in want]
    # Readable overload: values only, same order (MapTest.java:178-188 asserts equal counts)
    c = Collect()
    MAPS[family](kws, values, True).match(io.StringIO(hay), c)
    assert [v[0] for v in c.calls] == [values[int(r["value"])] for r in om.match(hay, readable=True)]


def _rand_word(rng, alphabet, lo, hi):
    return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("cs", [True, False])
def test_fuzz_small(family, cs):
    rng = random.Random(hash((family, cs, "gpu")) & 0xFFFF)
    for it in range(60):
        alphabet = rng.choice(["ab", "abc", "abAB", "abcdeXYZ", "aβΒbİi"])
        nk = rng.randint(0, 8)
        kws = [_rand_word(rng, alphabet, 1, rng.choice([2, 3, 5, 9])) for _ in range(nk)]
        if rng.random() < 0.2:
            kws.insert(rng.randint(0, len(kws)), rng.choice([None, ""]))
        sep = " " if family == "wholeword" or rng.random() < 0.3 else ""
        hay = "".join(rng.choice(alphabet + sep * 2) for _ in range(rng.randint(0, 200)))
        nv = rng.choice([len(kws), max(0, len(kws) - 1)])
        want = oracle_stream(ora.Matcher(family, kws, n_values=nv, case_sensitive=cs), hay)
        got = gpu_map_stream(MAPS[family](kws, list(range(nv)), cs), hay)
        assert got == want, (kws, hay, cs, nv)
        want_set = oracle_stream(ora.Matcher(family, kws, case_sensitive=cs), hay)
        assert gpu_set_stream(SETS[family](kws, cs), hay) == [(s, e) for s, e, _ in want_set]


@pytest.mark.parametrize("family", FAMILIES)
def test_fuzz_medium_dense(family):
    """Longer haystacks that span many tiles, dense matches, repetitive text (chains that never resynchronise)."""
    rng = random.Random(99)
    for alphabet, nk, n in (("ab", 12, 30000), ("abc", 40, 50000), ("ab ", 6, 20000)):
        kws = list({_rand_word(rng, alphabet.strip() or "ab", 1, 7) for _ in range(nk)})
        hay = "".join(rng.choice(alphabet) for _ in range(n))
        if family == "wholeword":
            hay = hay.replace("b", " ", n // 7) if " " not in alphabet else hay
        want = oracle_stream(ora.Matcher(family, kws, n_values=len(kws)), hay)
        got = gpu_map_stream(MAPS[family](kws, list(range(len(kws))), True), hay)
        assert got == want
    # periodic text: two interleaved greedy chains that never merge
    hay = "ab" * 20000
    kws = ["ab", "ba", "aba", "bab"]
    for cls_ in (SETS[family],):
        want = [(s, e) for s, e, _ in oracle_stream(ora.Matcher(family, kws), hay)]
        assert gpu_set_stream(cls_(kws, True), hay) == want


def test_early_stop_quirks():
    kws = ["ab", "b", "abc", "c"]
    hay = "abcabcabc"
    for family in FAMILIES:
        om = ora.Matcher(family, kws)
        full = oracle_stream(om, hay)
        for k in range(1, len(full) + 2):
            want = [(s, e) for s, e, _ in oracle_stream(om, hay, stop_after=k)]
            assert gpu_set_stream(SETS[family](kws, True), hay, stop_after=k) == want, (family, k)


def test_wide_alphabet_reference_style_dictionaries():
    """Generator.randomStrings(n, 2, 3) (Generator.java:61-76) + testFullNode: whole-BMP alphabets."""
    rng = random.Random(7)
    kws = list({"".join(chr(rng.randrange(256) if rng.random() < 0.5 else rng.randrange(65536))
                        for _ in range(2)) for _ in range(20000)})
    kws = [k for k in kws if not any(0xD800 <= ord(c) <= 0xDFFF for c in k)]
    hay = "The quick red fox, jumps over the lazy brown dog." + "".join(rng.choice(kws) for _ in range(500))
    for family in ("ahocorasick", "longest", "shortest"):
        want = [(s, e) for s, e, _ in oracle_stream(ora.Matcher(family, kws), hay)]
        assert gpu_set_stream(SETS[family](kws, True), hay) == want
        assert len(want) >= 250


def test_wholeword_custom_word_chars_config3_style():
    kws = ["a=b", "x-y", "_q_", "Hello"]
    wc = ora.word_chars(2, ["_", "="], [False, True])
    hay = "a=b_x-y q _q_ a=b= hello,Hello;HELLO"
    for cs in (True, False):
        want = [(s, e) for s, e, _ in oracle_stream(ora.Matcher("wholeword", kws, case_sensitive=cs, word_chars_table=wc), hay)]
        got = gpu_set_stream(ac.WholeWordMatchSet(kws, cs, ["_", "="], [False, True]), hay)
        assert got == want
    with pytest.raises(ac.IllegalArgumentException):
        ac.WholeWordMatchSet(["abc"], True, ["_", "="])


@pytest.mark.parametrize("seed", range(8))
def test_wholeword_case_insensitive_tables_not_closed_under_lowercase(seed):
    """Quirk Q7 (WholeWordMatchSet.java:96-101 vs :113,118; Readable: WholeWordMatchMap.java:325-339): with a custom
    word-character table that is NOT closed under toLowerCase the case-insensitive loop tests two different views of the
    table and its trie can hold non-word chars.  The GPU path follows the loop literally (kernel_wwlit.cuh): String and
    Readable overloads, Set and Map, equal the oracle's - also on text without any synchronisation point."""
    rng = random.Random(7700 + seed)
    letters = "abcdeABCDE"
    if seed % 2 == 0:   # generateWordCharsFlags(char[]): only the listed chars are word chars
        chars = [c for c in letters + "-_1" if rng.random() < 0.6] or ["A"]
        args, table = (chars,), ora.word_chars(1, chars, [])
    else:               # generateWordCharsFlags(char[], boolean[]): the default table with some chars toggled off
        chars = [c for c in letters if rng.random() < 0.4] or ["a"]
        toggles = [False] * len(chars)
        args, table = (chars, toggles), ora.word_chars(2, chars, toggles)
    closed = all(bool(table[ord(c)]) == bool(table[ord(c.lower())]) for c in letters)
    word = [c for c in letters + "-_1" if table[ord(c)]]
    kws = sorted({"".join(rng.choice(word) for _ in range(rng.randint(1, 5))) for _ in range(40)})
    values = list(range(len(kws)))
    om = ora.Matcher("wholeword", kws, n_values=len(kws), case_sensitive=False, word_chars_table=table)
    gs = ac.WholeWordMatchSet(kws, False, *args)
    gm = ac.WholeWordMatchMap(kws, values, False, *args)
    hays = ["", "a", "A b", "".join(rng.choice(letters) for _ in range(3000))]   # the last one: no separator at all
    for n, sep in ((50, " "), (700, " .,;"), (5000, " "), (70_001, " -_,")):
        hays.append("".join(rng.choice(sep) if rng.random() < 0.25 else rng.choice(letters + "1") for _ in range(n)))
    for hay in hays:
        want = oracle_stream(om, hay)
        assert gpu_set_stream(gs, hay) == [(s, e) for s, e, _ in want], (closed, len(hay))
        assert gpu_map_stream(gm, hay) == want, (closed, len(hay))
        c = Collect()
        gm.match(io.StringIO(hay), c)
        assert [v[0] for v in c.calls] == [int(r["value"]) for r in om.match(hay, readable=True)], (closed, len(hay))


@pytest.mark.parametrize("seed", range(6))
def test_wholewordlongest_case_insensitive_tables_not_closed_under_lowercase(seed):
    """Quirk Q7 for the fifth family (WholeWordLongestMatchSet.java:47-182, Map :183-310, Readable :54-181 with the scroll
    of :401-414): "followed by a non-word char" is decided on the lower-cased char, the scrolls on the raw char (String)
    or the lower-cased one (Readable).  k_segments<4> follows the loop literally."""
    rng = random.Random(8800 + seed)
    letters = "abcdeABCDE"
    if seed % 2 == 0:
        chars = [c for c in letters + "-_1" if rng.random() < 0.6] or ["A"]
        args, table = (chars,), ora.word_chars(1, chars, [])
    else:
        chars = [c for c in letters if rng.random() < 0.4] or ["a"]
        toggles = [False] * len(chars)
        args, table = (chars, toggles), ora.word_chars(2, chars, toggles)
    alphabet = letters + "-_1 "
    kws = sorted({"".join(rng.choice(alphabet) for _ in range(rng.randint(1, 7))).strip() for _ in range(60)} - {""})
    values = list(range(len(kws)))
    om = ora.Matcher("wholewordlongest", kws, n_values=len(kws), case_sensitive=False, word_chars_table=table)
    gs = ac.WholeWordLongestMatchSet(kws, False, *args)
    gm = ac.WholeWordLongestMatchMap(kws, values, False, *args)
    hays = ["", "a", "A b", "".join(rng.choice(letters) for _ in range(3000))]
    for n, sep in ((50, " "), (700, " .,;"), (5000, " "), (70_001, " -_,")):
        hays.append("".join(rng.choice(sep) if rng.random() < 0.25 else rng.choice(letters + "1") for _ in range(n)))
    for hay in hays:
        want = oracle_stream(om, hay)
        assert gpu_set_stream(gs, hay) == [(s, e) for s, e, _ in want], len(hay)
        assert gpu_map_stream(gm, hay) == want, len(hay)
        c = Collect()
        gm.match(io.StringIO(hay), c)
        assert [v[0] for v in c.calls] == [int(r["value"]) for r in om.match(hay, readable=True)], len(hay)


@pytest.mark.parametrize("family,long_len", [("longest", 2048), ("longest", 5000), ("shortest", 2048), ("shortest", 3001),
                                             ("wholewordlongest", 255), ("wholewordlongest", 4000)])
def test_keywords_longer_than_the_selection_kernels_hold(family, long_len):
    """The reference takes keywords of any length (LongestMatchSet.java:20-190 has no limit).  Dictionaries whose longest
    keyword exceeds what the selection kernels keep in shared memory (2 046 chars; 254 for WholeWordLongest) run the
    reference's loop one thread per synchronisation point (k_segments): same ordered stream as the oracle, String and
    Readable overloads, early stop."""
    rng = random.Random(long_len)
    sigma = "abc"
    longs = ["".join(rng.choice(sigma) for _ in range(long_len)), "ab" * (long_len // 2 - 3)]
    if family == "wholewordlongest":
        longs.append(("abc " * long_len)[:long_len - 1].strip())
    kws = sorted({_rand_word(rng, sigma, 1, 6) for _ in range(40)} | set(longs))
    values = list(range(len(kws)))
    om = ora.Matcher(family, kws, n_values=len(kws))
    gs, gm = SETS[family](kws, True), MAPS[family](kws, values, True)
    from ahocorasick_b200 import _lib
    assert _lib.lib().acgpu_launches_per_match(gm.handle) == 3   # count, scan, write of k_segments
    body = "".join(rng.choice(sigma + "  .") for _ in range(40_000))
    hays = ["", longs[0], longs[0][:-1], longs[1] + "ab", body + " " + longs[0] + " " + body[:500] + longs[1] + "." + longs[-1] + " x",
            "".join(rng.choice(sigma) for _ in range(20_000)) + longs[0] * 2]   # the last one: no synchronisation point
    for hay in hays:
        want = oracle_stream(om, hay)
        assert gpu_set_stream(gs, hay) == [(s, e) for s, e, _ in want], len(hay)
        assert gpu_map_stream(gm, hay) == want, len(hay)
        c = Collect()
        gm.match(io.StringIO(hay), c)
        assert [v[0] for v in c.calls] == [int(r["value"]) for r in om.match(hay, readable=True)], len(hay)
    want = oracle_stream(om, hays[4])
    assert len(want) > 100
    assert gpu_map_stream(gm, hays[4], stop_after=7) == oracle_stream(om, hays[4], stop_after=7)


@pytest.mark.parametrize("gen1", [False, True])
@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])
def test_baseline_configs_scaled(cfg, gen1, monkeypatch):
    """The five BASELINE.json configs at a size the oracle finishes in seconds: identical ordered streams.
    gen1=True forces the general anchored-trie kernel where the tiered kernel would be chosen."""
    if gen1:
        monkeypatch.setenv("ACGPU_FORCE_GEN1", "1")
    c = W.config(cfg, scale=0.02 if cfg != 0 else 0.25)
    n = min(c["n"], 2_000_000)
    hay = W.make_haystack(c["spec"], n)
    kws = c["keywords"]
    wc_args = ()
    wc_table = None
    if "word_chars" in c:
        wc_args = tuple(c["word_chars"])
        wc_table = ora.word_chars(2, *c["word_chars"])
    families = [c["family"]] if cfg != 2 else ["longest", "shortest"]
    for family in families:
        om = ora.Matcher(family, kws, n_values=len(kws), case_sensitive=c["cs"], word_chars_table=wc_table)
        want = om.match(hay)
        gm = MAPS[family](kws, list(range(len(kws))), c["cs"], *wc_args)
        rec = gm.match_records(hay)
        assert len(rec) == len(want) and len(want) > 1000
        assert np.array_equal(rec.start, want["start"])
        assert np.array_equal(rec.end, want["end"])
        assert np.array_equal(rec.value.astype(np.int64), want["value"].astype(np.int64))


class _NumpyReadable:
    """A Readable over a uint16 array (like java.io.CharArrayReader): read(n) returns up to n chars."""

    def __init__(self, arr):
        self.arr, self.at = arr, 0

    def read(self, n):
        out = self.arr[self.at:self.at + n]
        self.at += out.size
        return out


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("block", [4096, 3 * 4096, 1 << 16])
def test_readable_streaming_blocks(family, block):
    """match(Readable): device blocks of several sizes (context / chain state carried across blocks) must give
    the oracle's Readable value stream, including ShortestMatchMap's fill-boundary duplicates (Q4)."""
    from ahocorasick_b200.streaming import match_readable
    rng = random.Random(block)
    kws = list({_rand_word(rng, "abc", 1, 9) for _ in range(60)}) + ["a" * 40, "cab" * 20]
    n = 150_000
    hay = "".join(rng.choice("abc  ") for _ in range(n))
    if family == "wholeword":
        kws = [k for k in kws]
    values = list(range(len(kws)))
    om = ora.Matcher(family, kws, n_values=len(kws))
    want = [int(r["value"]) for r in om.match(hay, readable=True)]
    gm = MAPS[family](kws, values, True)
    got = []
    match_readable(gm, io.StringIO(hay), lambda v: got.append(v) or True, block_chars=block)
    assert got == want
    # early stop through the Readable path
    for stop in (1, 7, 500):
        if stop > len(want):
            continue
        c = Collect(stop)
        match_readable(gm, io.StringIO(hay), c.match, block_chars=block)
        w = [int(r["value"]) for r in om.match(hay, readable=True, stop_after=stop)]
        assert [x[0] for x in c.calls] == w


@pytest.mark.parametrize("family", ["longest", "shortest"])
@pytest.mark.parametrize("block", [1000, 8192, 50_000, 1 << 18])
def test_readable_streaming_chain_blocks(family, block, monkeypatch):
    """Longest / Shortest match(Readable) on the start-mask path: every feed is a chain shard (whole 8 192-position tiles with
    look-ahead, entry offset carried by the composed map).  The value stream - with ShortestMatchMap's fill-boundary duplicates
    (Q4) - and the (start, end, value) records equal the oracle's for feeds smaller than, equal to and larger than a tile, on
    text with and without separators, and equal what the generation-1 kernels deliver."""
    from ahocorasick_b200.streaming import DeviceStream, match_readable
    rng = random.Random(block + len(family))
    kws = sorted({_rand_word(rng, "abcde", 1, 12) for _ in range(3000)})
    values = list(range(len(kws)))
    om = ora.Matcher(family, kws, n_values=len(kws))
    gm, gs = MAPS[family](kws, values, True), SETS[family](kws, True)
    for n, alphabet in ((0, "abcde "), (5, "abcde"), (70_000, "abcde  "), (300_001, "abcde")):
        hay = "".join(rng.choice(alphabet) for _ in range(n))
        want = [int(r["value"]) for r in om.match(hay, readable=True)]
        got = []
        match_readable(gm, io.StringIO(hay), lambda v: got.append(v) or True, block_chars=block)
        assert got == want, (n, len(got), len(want))
        # the records themselves, block by block
        arr = np.frombuffer(hay.encode("utf-16-le"), dtype=np.uint16)
        st = DeviceStream(gs)
        recs = []
        for at in range(0, arr.size, block):
            r = st.feed(arr[at:at + block])
            recs += list(zip(r.start.tolist(), r.end.tolist()))
        r = st.end()
        recs += list(zip(r.start.tolist(), r.end.tolist()))
        assert recs == [(a, b) for a, b, _ in oracle_stream(om, hay, cap=max(1 << 16, n + 1))], n
    if len(want) > 600:
        for stop in (1, 7, 500):
            got = []
            match_readable(gm, io.StringIO(hay), lambda v: got.append(v) or len(got) < stop, block_chars=block)
            assert got == [int(r["value"]) for r in om.match(hay, readable=True, stop_after=stop)]


def test_readable_config3_wholeword_map():
    """configs[3]: WholeWordMatchMap with the toggle word-char constructor, streamed via Readable."""
    c = W.config(3, scale=0.02)
    hay = W.make_haystack(c["spec"], 1_000_000)
    kws = c["keywords"]
    om = ora.Matcher("wholeword", kws, n_values=len(kws), word_chars_table=ora.word_chars(2, *c["word_chars"]))
    want = [int(r["value"]) for r in om.match(hay, readable=True)]
    gm = ac.WholeWordMatchMap(kws, list(range(len(kws))), True, *c["word_chars"])
    got = []
    gm.match(_NumpyReadable(hay), lambda v: got.append(v) or True)
    assert got == want and len(want) > 100


def test_full_1m_dictionary_parity(monkeypatch):
    """configs[4] dictionary at FULL size (10^6 keywords, ~4.4M trie nodes) on a 4M-char slice: both kernel
    generations against the oracle (dictionary-size dependent bugs do not show at the scaled configs)."""
    c = W.config(4)
    kws = c["keywords"]
    hay = W.make_haystack(c["spec"], 4_000_000)
    want = ora.Matcher("ahocorasick", kws).match(hay)
    for gen1 in (False, True):
        if gen1:
            monkeypatch.setenv("ACGPU_FORCE_GEN1", "1")
        rec = ac.AhoCorasickSet(kws, True).match_records(hay)
        assert len(rec) == len(want), (gen1, len(rec), len(want))
        assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"])


# ---------------------------------------------------------------- committed fixtures (tests/golden/oracle_streams.json)

def _golden_cases():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_streams.json")))


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_committed_fixture_streams(case):
    """The CUDA path against the committed ordered streams — no oracle call in this test."""
    kws, hay, cs = case["keywords"], case["haystack"], case["case_sensitive"]
    values = list(range(len(kws)))
    for fam in FAMILIES:
        want = case["streams"][fam]
        extra = ()
        if case.get("word_chars") and fam.startswith("wholeword"):
            w = case["word_chars"]
            extra = (w["chars"], w["toggles"])
        if "error" in want:
            with pytest.raises(ac.IllegalArgumentException):
                SETS[fam](kws, cs, *extra)
            continue
        assert gpu_set_stream(SETS[fam](kws, cs, *extra), hay) == [(s, e) for s, e, _ in want["string"]]
        assert gpu_map_stream(MAPS[fam](kws, values, cs, *extra), hay) == [tuple(t) for t in want["string"]]
        c = Collect()
        MAPS[fam](kws, values, cs, *extra).match(io.StringIO(hay), c)
        assert [v[0] for v in c.calls] == want["readable_values"]


# ---------------------------------------------------------------- mask / scan / emit path: shapes the fuzz tests miss

def _records(m, hay):
    rec = m.match_records(hay)
    return list(zip(rec.start.tolist(), rec.end.tolist())), (rec.value.tolist() if rec.value is not None else None)


@pytest.mark.parametrize("alphabet,max_kw", [("ab", 16), ("acgt", 16), ("0123456789", 12), ("abcdefghijklmnopqrstuvwxyz", 12),
                                              ("abcdefghijklmnopqrstuvwxyzABCDE", 9)])
def test_tier_path_alphabets_and_boundary_lengths(alphabet, max_kw):
    """Class counts 3..32 pick different K (direct-indexed levels 8..3) and LOW variants of k_tier_mask; haystack
    lengths straddle the 8-char lane, 256-char row and 8 192-char chunk boundaries; Set and Map streams must equal
    the oracle's, including the empty haystack."""
    rng = random.Random(len(alphabet) * 100 + max_kw)
    kws = sorted({_rand_word(rng, alphabet, 1 if len(alphabet) < 8 else 2, max_kw) for _ in range(400)})
    values = list(range(len(kws)))
    om = ora.Matcher("ahocorasick", kws, n_values=len(kws))
    gs, gm = ac.AhoCorasickSet(kws, True), ac.AhoCorasickMap(kws, values, True)
    sep = " " if len(alphabet) > 4 else ""
    base = "".join(rng.choice(alphabet + sep) for _ in range(70_001))
    for n in (0, 1, 7, 8, 9, 255, 256, 257, 511, 8191, 8192, 8193, 70_001):
        hay = base[:n]
        want = oracle_stream(om, hay)
        pos, _ = _records(gs, hay)
        assert pos == [(s, e) for s, e, _ in want], (alphabet, n)
        pos, val = _records(gm, hay)
        assert pos == [(s, e) for s, e, _ in want] and [int(v) for v in val] == [v for _, _, v in want], (alphabet, n)


def test_tier_path_dense_rows_and_long_chains():
    """'aaaa…' with nested keywords a^1..a^16 emits 16 records per position: rows overflow the staging window of
    k_tier_emit (windowed path) and every context continues past level K (deep-probe queue runs full)."""
    kws = ["a" * i for i in range(1, 17)] + ["ab", "ba" * 4, "b" * 11]
    hay = "a" * 3000 + "b" * 40 + ("ab" * 700) + "a" * 513
    want = oracle_stream(ora.Matcher("ahocorasick", kws, n_values=len(kws)), hay)
    pos, _ = _records(ac.AhoCorasickSet(kws, True), hay)
    assert pos == [(s, e) for s, e, _ in want] and len(pos) > 50_000
    pos, val = _records(ac.AhoCorasickMap(kws, list(range(len(kws))), True), hay)
    assert pos == [(s, e) for s, e, _ in want] and [int(v) for v in val] == [v for _, _, v in want]


def test_device_range_shards_and_cap():
    """acgpu_match_device on end-position ranges of a resident haystack (the multi-GPU shard entry): shards at odd
    boundaries concatenate to the single stream; a capacity below the total truncates the stream, not the count."""
    import ctypes as C
    import torch
    from ahocorasick_b200 import _lib
    from ahocorasick_b200.sharding import plan_range_shards
    c = W.config(0, scale=0.25)
    kws = c["keywords"]
    hay = W.make_haystack(c["spec"], 300_000)
    want = ora.Matcher("ahocorasick", kws).match(hay)
    want_pos = np.stack([want["start"], want["end"]], axis=1).astype(np.int32)
    m = ac.AhoCorasickSet(kws, True)
    lib = _lib.lib()
    d_hay = torch.from_numpy(hay.astype(np.int16)).cuda()
    d_pos = torch.empty((len(want) + 16, 2), dtype=torch.int32, device="cuda")

    def run(lo, hi, cap):
        tot = C.c_int64(0)
        _lib.check(lib.acgpu_match_device(m.handle, d_hay.data_ptr(), hay.size, lo, hi, d_pos.data_ptr(), None, cap, C.byref(tot), None))
        torch.cuda.synchronize()
        return tot.value, d_pos[:min(tot.value, cap)].cpu().numpy()

    for world in (2, 3, 8):
        parts = []
        for sh in plan_range_shards(hay.size, world, max_len=12, align=1 if world == 3 else 8):
            n_sh, rec = run(sh.emit_from, sh.emit_to, len(want) + 16)
            parts.append(rec.copy())
        got = np.concatenate(parts, axis=0)
        assert got.shape == want_pos.shape and np.array_equal(got, want_pos), world
    total, rec = run(0, hay.size, 1000)
    assert total == len(want) and np.array_equal(rec, want_pos[:1000])
    total, rec = run(0, hay.size, 0)
    assert total == len(want)


@pytest.mark.parametrize("is_map", [False, True])
def test_wholeword_word_start_range_shards(is_map):
    """SURVEY 8e, WholeWord: a haystack cut by word-START range at arbitrary boundaries (also inside words) - every shard
    sees only [read_from, read_to) semantics-wise (n = read_to) and the rank-ordered concatenation is the single stream."""
    import ctypes as C
    import torch
    from ahocorasick_b200 import _lib
    from ahocorasick_b200.sharding import plan_word_shards
    c = W.config(3, scale=0.05)
    kws = c["keywords"]
    hay = W.make_haystack(c["spec"], 400_003)
    want = ora.Matcher("wholeword", kws, n_values=len(kws), word_chars_table=ora.word_chars(2, *c["word_chars"])).match(hay)
    want_pos = np.stack([want["start"], want["end"]], axis=1).astype(np.int32)
    assert len(want) > 500
    m = (ac.WholeWordMatchMap(kws, list(range(len(kws))), True, *c["word_chars"]) if is_map
         else ac.WholeWordMatchSet(kws, True, *c["word_chars"]))
    max_len = m.info()["max_len"]
    lib = _lib.lib()
    d_hay = torch.from_numpy(hay.astype(np.int16)).cuda()
    d_pos = torch.empty((len(want) + 16, 2), dtype=torch.int32, device="cuda")
    d_val = torch.empty(len(want) + 16, dtype=torch.int32, device="cuda")
    for world in (1, 2, 3, 7, 64):
        pos_parts, val_parts = [], []
        for sh in plan_word_shards(hay.size, world, max_len):
            tot = C.c_int64(0)
            _lib.check(lib.acgpu_match_device(m.handle, d_hay.data_ptr(), sh.read_to, sh.emit_from, sh.emit_to, d_pos.data_ptr(),
                                              d_val.data_ptr() if is_map else None, len(want) + 16, C.byref(tot), None))
            torch.cuda.synchronize()
            pos_parts.append(d_pos[:tot.value].cpu().numpy().copy())
            val_parts.append(d_val[:tot.value].cpu().numpy().copy())
        got = np.concatenate(pos_parts, axis=0)
        assert got.shape == want_pos.shape and np.array_equal(got, want_pos), world
        if is_map:
            assert np.array_equal(np.concatenate(val_parts).astype(np.int64), want["value"].astype(np.int64)), world
    tot = C.c_int64(0)
    rc = lib.acgpu_match_device(m.handle, d_hay.data_ptr(), hay.size, 10, hay.size + 1, d_pos.data_ptr(), None, 16, C.byref(tot), None)
    assert rc == _lib.EINVAL


@pytest.mark.parametrize("family,is_map", [("longest", True), ("shortest", False), ("ahocorasick", True)])
def test_sync_point_shards_on_device(family, is_map):
    """SURVEY 8e for the chain families: one resident haystack cut at synchronisation points (chars that occur in no
    keyword, matcher.char_classes()) into independent pieces; every piece is scanned as a haystack of its own
    (unaligned device pointers included) and the shifted, rank-ordered concatenation is the oracle's single stream."""
    import torch
    from ahocorasick_b200.sharding import match_sync_shard, plan_sync_shards
    c = W.config(2, scale=0.02)
    kws = c["keywords"]
    hay = W.make_haystack(c["spec"], 500_000)
    want = ora.Matcher(family, kws, n_values=len(kws) if is_map else -1).match(hay)
    want_pos = np.stack([want["start"], want["end"]], axis=1).astype(np.int64)
    m = (MAPS if is_map else SETS)[family](*((kws, list(range(len(kws))), True) if is_map else (kws, True)))
    classes, has_other = m.char_classes()
    assert has_other and classes[ord(" ")] == 0 and classes[ord("a")] != 0 and classes[ord("A")] == 0
    d_hay = torch.from_numpy(hay.astype(np.int16)).cuda()
    cap = len(want) + 16
    d_pos = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
    d_val = torch.empty(cap, dtype=torch.int32, device="cuda")
    for world in (2, 3, 8):
        shards = plan_sync_shards(d_hay, world, classes, has_other)
        assert shards is not None and len(shards) == world
        assert any(s.lo % 8 for s in shards[1:])          # pieces start at arbitrary (unaligned) chars
        pos_parts, val_parts = [], []
        for sh in shards:
            k = match_sync_shard(m, d_hay.data_ptr(), sh, d_pos.data_ptr(), d_val.data_ptr() if is_map else None, cap)
            torch.cuda.synchronize()
            pos_parts.append(d_pos[:k].cpu().numpy().astype(np.int64) + sh.lo)
            val_parts.append(d_val[:k].cpu().numpy().astype(np.int64))
        got = np.concatenate(pos_parts, axis=0)
        assert got.shape == want_pos.shape and np.array_equal(got, want_pos), (family, world)
        if is_map:
            assert np.array_equal(np.concatenate(val_parts), want["value"].astype(np.int64)), (family, world)
    assert len(want) > 1000


# ---------------------------------------------------------------- Longest / Shortest on the start-mask path (kernel_sel2.cuh)

@pytest.mark.parametrize("family", ["longest", "shortest"])
@pytest.mark.parametrize("alphabet,max_kw", [("ab", 16), ("acgt", 13), ("abcdefghijklmnopqrstuvwxyz", 12)])
def test_sel2_alphabets_and_boundary_lengths(family, alphabet, max_kw):
    """Mirrored k_tier_mask + exit maps: haystack lengths straddle the 8-char lane group, the 32-position lane
    sub-tile, the 256-position row, the 1 024-position warp and the 8 192-position tile; Set and Map streams must
    equal the oracle's.  gen-1 (k_fwd_v + k_sel_*) is checked on the same inputs."""
    rng = random.Random(len(alphabet) * 1000 + max_kw + len(family))
    kws = sorted({_rand_word(rng, alphabet, 1 if len(alphabet) < 8 else 2, max_kw) for _ in range(300)})
    values = list(range(len(kws)))
    om = ora.Matcher(family, kws, n_values=len(kws))
    gs, gm = SETS[family](kws, True), MAPS[family](kws, values, True)
    sep = " " if len(alphabet) > 4 else ""
    base = "".join(rng.choice(alphabet + sep) for _ in range(70_001))
    for n in (0, 1, 7, 8, 9, 31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 8191, 8192, 8193, 8207, 16385, 70_001):
        hay = base[:n]
        want = oracle_stream(om, hay)
        pos, _ = _records(gs, hay)
        assert pos == [(s, e) for s, e, _ in want], (alphabet, n)
        pos, val = _records(gm, hay)
        assert pos == [(s, e) for s, e, _ in want] and [int(v) for v in val] == [v for _, _, v in want], (alphabet, n)


@pytest.mark.parametrize("family", ["longest", "shortest"])
def test_sel2_dense_periodic_and_sparse(family):
    """Chains that never resynchronise (periodic text), one match per position (nested a^k), sparse text where whole
    lanes / warps / tiles are skipped, and keywords of the maximum length 16 ending at the very end."""
    cases = [
        (["a" * i for i in range(1, 17)] + ["ab", "ba" * 4, "b" * 11], "a" * 3000 + "b" * 40 + ("ab" * 700) + "a" * 513),
        (["ab", "ba", "aba", "bab", "abab" * 4], "ab" * 40000),
        (["needle", "nee", "dle", "edl"], ("x" * 20011 + "needle") * 9 + "x" * 7 + "nee"),
        (["abcdefghijklmnop", "p", "op", "ponm"], ("abcdefghijklmnop" * 2100)),
        (["aab", "ab", "b", "aaaaaaaaab"], ("a" * 9 + "b") * 5000 + "aaaa"),
    ]
    for kws, hay in cases:
        want = oracle_stream(ora.Matcher(family, kws, n_values=len(kws)), hay)
        pos, val = _records(MAPS[family](kws, list(range(len(kws))), True), hay)
        assert pos == [(s, e) for s, e, _ in want], (family, kws)
        assert [int(v) for v in val] == [v for _, _, v in want], (family, kws)
        assert len(pos) >= 10


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("family", ["longest", "shortest"])
def test_sel2_many_tiles_scan_slices(family, fused, monkeypatch):
    """10^7 chars = 1 221 tiles = 5 composition groups (k_sel2_group / _top / _tiles); fused=True runs the opt-in
    single-pass kernel instead (k_sel2_fused: decoupled look-back over the tile maps)."""
    if fused:
        monkeypatch.setenv("ACGPU_SEL2_FUSED", "1")
    c = W.config(2, scale=0.02)
    hay = W.make_haystack(c["spec"], 10_000_000)
    kws = c["keywords"]
    want = ora.Matcher(family, kws).match(hay)
    rec = SETS[family](kws, True).match_records(hay)
    assert len(rec) == len(want) and len(want) > 100_000
    assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"])


@pytest.mark.parametrize("family", ["longest", "shortest"])
def test_sel2_device_unaligned_and_cap(family):
    """acgpu_match_device on a resident haystack whose first char sits at every 16-byte misalignment (the mirrored
    kernel derives its row origin from the pointer and the length), and with a record capacity below the total."""
    import ctypes as C
    import torch
    from ahocorasick_b200 import _lib
    c = W.config(2, scale=0.01)
    kws = c["keywords"]
    full = W.make_haystack(c["spec"], 100_000 + 16)
    m = SETS[family](kws, True)
    om = ora.Matcher(family, kws)
    lib = _lib.lib()
    d_full = torch.from_numpy(full.astype(np.int16)).cuda()
    d_pos = torch.empty((60_000, 2), dtype=torch.int32, device="cuda")
    for off in range(8):
        for n in (100_000, 99_997):
            hay = full[off:off + n]
            want = om.match(hay)
            want_pos = np.stack([want["start"], want["end"]], axis=1).astype(np.int32)
            tot = C.c_int64(0)
            _lib.check(lib.acgpu_match_device(m.handle, d_full.data_ptr() + 2 * off, n, 0, n, d_pos.data_ptr(), None, 60_000,
                                              C.byref(tot), None))
            torch.cuda.synchronize()
            assert tot.value == len(want), (off, n)
            assert np.array_equal(d_pos[:tot.value].cpu().numpy(), want_pos), (off, n)
    tot = C.c_int64(0)
    _lib.check(lib.acgpu_match_device(m.handle, d_full.data_ptr(), 100_000, 0, 100_000, d_pos.data_ptr(), None, 777, C.byref(tot), None))
    torch.cuda.synchronize()
    want = om.match(full[:100_000])
    assert tot.value == len(want)
    assert np.array_equal(d_pos[:777].cpu().numpy(), np.stack([want["start"], want["end"]], axis=1).astype(np.int32)[:777])


# ---------------------------------------------------------------- WholeWord on the hash path (kernel_ww.cuh)

@pytest.mark.parametrize("cs", [True, False])
def test_ww_hash_path_shapes(cs):
    """Words that straddle the 4 096-position tiles, words longer than every keyword, keywords up to 200 chars (the
    right context staged per tile grows with max_len), words at both ends of the haystack, non-ASCII letters, words
    that share a hash bucket's probe path, and a dictionary where every word of the text is a keyword (dense output)."""
    rng = random.Random(4242 + cs)
    alphabet = "abcdefgXYZ0123456789-_" + "βΒжЖé"
    kws = sorted({_rand_word(rng, alphabet, 1, 9) for _ in range(3000)})
    kws += ["q" * 200, "Ab" * 60, "z" * 17]
    words = kws + [_rand_word(rng, alphabet, 1, 12) for _ in range(2000)] + ["q" * 201, "q" * 199, "Z" * 17, "z" * 18]
    seps = [" ", ", ", ".", "\n", " ; ", "(", ")", "  "]
    parts, total = [], 0
    while total < 120_000:
        w = rng.choice(words)
        sp = rng.choice(seps)
        parts.append(w + sp)
        total += len(w) + len(sp)
    body = "".join(parts)
    for hay in (body, body.rstrip(" ,.;()\n") , "q" * 200, "", " ", "ab", body[:4096], body[:4097], body[1:8193]):
        om = ora.Matcher("wholeword", kws, n_values=len(kws), case_sensitive=cs)
        want = oracle_stream(om, hay)
        pos, val = _records(ac.WholeWordMatchMap(kws, list(range(len(kws))), cs), hay)
        assert pos == [(s, e) for s, e, _ in want], (cs, len(hay))
        assert [int(v) for v in val] == [v for _, _, v in want], (cs, len(hay))
        pos, _ = _records(ac.WholeWordMatchSet(kws, cs), hay)
        assert pos == [(s, e) for s, e, _ in want], (cs, len(hay))
    # every word a keyword
    dense = " ".join(rng.choice(kws[:500]) for _ in range(30_000))
    want = oracle_stream(ora.Matcher("wholeword", kws[:500]), dense)
    pos, _ = _records(ac.WholeWordMatchSet(kws[:500], True), dense)
    assert pos == [(s, e) for s, e, _ in want] and len(pos) == 30_000


# ---------------------------------------------------------------- WholeWordLongest (the fifth public family)

@pytest.mark.parametrize("cs", [True, False])
def test_wholewordlongest_multiword_fuzz(cs):
    """Keywords with inner / outer non-word chars, walks that consume several words, carried fail matches, walk start
    at position 0 on a non-word char: ordered Set and Map streams equal the oracle's."""
    from test_oracle_golden import _wwl_case
    rng = random.Random(9001 + cs)
    for it in range(120):
        kws, hay = _wwl_case(rng)
        hay = hay * rng.choice([1, 1, 7])
        nv = rng.choice([len(kws), max(0, len(kws) - 1)])
        want = oracle_stream(ora.Matcher("wholewordlongest", kws, n_values=nv, case_sensitive=cs), hay)
        got = gpu_map_stream(ac.WholeWordLongestMatchMap(kws, list(range(nv)), cs), hay)
        assert got == want, (kws, hay, cs, nv)
        want_set = oracle_stream(ora.Matcher("wholewordlongest", kws, case_sensitive=cs), hay)
        assert gpu_set_stream(ac.WholeWordLongestMatchSet(kws, cs), hay) == [(s, e) for s, e, _ in want_set]


def test_wholewordlongest_phrases_large():
    """A phrase dictionary over a 300 000-char text (many selection tiles; walks cross tile boundaries), String and
    Readable overloads, early stop."""
    rng = random.Random(31337)
    vocab = ["new", "york", "city", "hall", "of", "fame", "san", "jose", "los", "angeles", "x-ray", "st", "louis", "a", "b"]
    kws = sorted({" ".join(rng.choice(vocab) for _ in range(rng.randint(1, 4))) for _ in range(400)}) + ["  padded  ", "st. louis"]
    seps = [" ", " ", " ", ", ", ". ", "  ", "; ", "-", "_"]
    parts, total = [], 0
    while total < 300_000:
        w = rng.choice(vocab + ["yorker", "halls", "zz"])
        sp = rng.choice(seps)
        parts.append(w + sp)
        total += len(w) + len(sp)
    hay = "".join(parts)
    values = list(range(len(kws)))
    om = ora.Matcher("wholewordlongest", kws, n_values=len(kws), case_sensitive=False)
    want = oracle_stream(om, hay)
    gm = ac.WholeWordLongestMatchMap(kws, values, False)
    assert gpu_map_stream(gm, hay) == want and len(want) > 20_000
    for block in (4096, 1 << 16):
        from ahocorasick_b200.streaming import match_readable
        got = []
        match_readable(gm, io.StringIO(hay), lambda v: got.append(v) or True, block_chars=block)
        assert got == [int(r["value"]) for r in om.match(hay, readable=True)]
    for stop in (1, 5, 1000):
        want_s = [(s, e) for s, e, _ in oracle_stream(om, hay, stop_after=stop)]
        assert gpu_set_stream(ac.WholeWordLongestMatchSet(kws, False), hay, stop_after=stop) == want_s


# ---------------------------------------------------------------- full-size dictionaries, large haystacks: properties

@pytest.mark.parametrize("cfg,family", [(2, "longest"), (2, "shortest"), (3, "wholeword")])
def test_large_haystack_two_generations_agree(cfg, family, monkeypatch):
    """BASELINE configs[2] / configs[3] with their FULL dictionaries over 3*10^8 device-generated chars (too long for the
    CPU oracle): the start-mask / hash kernels and the generation-1 anchored-trie kernels are independent
    implementations, so identical record streams pin both; plus the stream properties that hold at any size
    (ascending, non-overlapping, lengths within the dictionary's range)."""
    import ctypes as C
    import torch
    from ahocorasick_b200 import _lib
    c = W.config(cfg)
    kws = c["keywords"]
    extra = tuple(c["word_chars"]) if "word_chars" in c else ()
    n = 300_000_000
    hay = W.make_haystack_torch(c["spec"], n, device=torch.device("cuda", 0))
    lib = _lib.lib()

    def run(m):
        tot = C.c_int64(0)
        _lib.check(lib.acgpu_match_device(m.handle, hay.data_ptr(), n, 0, n, None, None, 0, C.byref(tot), None))
        d_pos = torch.empty((tot.value, 2), dtype=torch.int32, device="cuda")
        _lib.check(lib.acgpu_match_device(m.handle, hay.data_ptr(), n, 0, n, d_pos.data_ptr(), None, tot.value, C.byref(tot), None))
        torch.cuda.synchronize()
        return d_pos

    fast = run(SETS[family](kws, c["cs"], *extra))
    monkeypatch.setenv("ACGPU_FORCE_GEN1", "1")
    slow = run(SETS[family](kws, c["cs"], *extra))
    assert fast.shape == slow.shape and fast.shape[0] > 100_000
    assert torch.equal(fast, slow)
    start, end = fast[:, 0].long(), fast[:, 1].long()
    assert bool((end > start).all()) and bool((end - start <= 12).all())
    assert bool((start[1:] >= end[:-1]).all())          # ascending and non-overlapping
    assert int(start[0]) >= 0 and int(end[-1]) <= n


# ---------------------------------------------------------------- randomised medium-size stress over every family

@pytest.mark.parametrize("seed", range(12))
def test_random_dictionaries_medium_haystacks(seed):
    """Random alphabets (2..30 symbols, sometimes beyond ASCII), keyword lengths up to 16, separator densities from none to
    one in three, case folding on and off, 40 000..120 000 chars: all five families, Set and Map, against the oracle."""
    rng = random.Random(1000 + seed)
    pool = "abcdefghijklmnopqrstuvwxyzABCD0123456789" + "βγδжзиé"
    alphabet = "".join(rng.sample(pool, rng.randint(2, 30)))
    max_kw = rng.choice([3, 6, 12, 16])
    nk = rng.choice([5, 40, 400, 3000])
    kws = sorted({_rand_word(rng, alphabet, 1, max_kw) for _ in range(nk)})
    if rng.random() < 0.5:  # phrases for WholeWordLongest (the other families treat the space as a keyword char)
        kws += [rng.choice(kws) + " " + rng.choice(kws) for _ in range(max(1, nk // 10))]
        kws = [k for k in kws if len(k) <= 16]
    sep_rate = rng.choice([0.0, 0.05, 0.15, 0.33])
    n = rng.randint(40_000, 120_000)
    seps = " ,.-"
    hay = "".join(rng.choice(seps) if rng.random() < sep_rate else rng.choice(alphabet) for _ in range(n))
    # plant keywords so that long ones occur too
    hay_l = list(hay)
    for _ in range(n // 50):
        k = rng.choice(kws)
        at = rng.randrange(0, n - len(k))
        hay_l[at:at + len(k)] = k
    hay = "".join(hay_l)
    cs = rng.random() < 0.5
    if not cs:
        hay = "".join(c.upper() if rng.random() < 0.3 else c for c in hay)
    values = list(range(len(kws)))
    for family in FAMILIES:
        try:
            om = ora.Matcher(family, kws, n_values=len(kws), case_sensitive=cs)
        except ora.OracleError:
            with pytest.raises(ac.IllegalArgumentException):
                MAPS[family](kws, values, cs)
            continue
        want = om.match(hay)
        rec = MAPS[family](kws, values, cs).match_records(hay)
        assert len(rec) == len(want), (seed, family, len(rec), len(want))
        assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"]), (seed, family)
        assert np.array_equal(rec.value.astype(np.int64), want["value"].astype(np.int64)), (seed, family)
        rec = SETS[family](kws, cs).match_records(hay)
        assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"]), (seed, family, "set")


@pytest.mark.parametrize("family", FAMILIES)
def test_concurrent_match_on_one_instance(family):
    """Threading contract of the reference: the automaton is immutable after the constructor and every
    match() keeps its state in locals, so concurrent match() calls on ONE instance are safe.  Here: 6 threads share one
    Map, each runs String matches over its own haystacks plus one Readable match, all against oracle streams computed
    up front (ctypes releases the GIL during the native call, so the calls really overlap)."""
    import threading
    word = family.startswith("wholeword")
    c = W.config(3 if word else 2, scale=0.02)
    kws = list(c["keywords"])
    if family == "wholewordlongest":
        kws = kws + [a + " " + b for a, b in zip(kws[::9], kws[4::9])]
    if word:
        m = MAPS[family](kws, list(range(len(kws))), True, *c["word_chars"])
        o = ora.Matcher(family, kws, n_values=len(kws), word_chars_table=ora.word_chars(2, *c["word_chars"]))
    else:
        m = MAPS[family](kws, list(range(len(kws))), True)
        o = ora.Matcher(family, kws, n_values=len(kws))
    n_threads, per_thread = 6, 3
    hays = [W.make_haystack(c["spec"], 150_000 + 37_000 * i, start=i * W.BLOCK * 8192) for i in range(n_threads * per_thread)]
    want = [o.match(h) for h in hays]
    want_readable = [o.match(hays[t * per_thread], readable=True)["value"].astype(np.int64) for t in range(n_threads)]
    errors = []
    barrier = threading.Barrier(n_threads)

    def work(t):
        try:
            barrier.wait()
            for rep in range(2):
                for i in range(t * per_thread, (t + 1) * per_thread):
                    rec = m.match_records(hays[i])
                    w = want[i]
                    assert len(rec) == len(w), (t, i, len(rec), len(w))
                    assert np.array_equal(rec.start, w["start"]) and np.array_equal(rec.end, w["end"]), (t, i)
                    assert np.array_equal(rec.value.astype(np.int64), w["value"].astype(np.int64)), (t, i)
            got = []
            m.match(_NumpyReadable(hays[t * per_thread]), lambda v: got.append(v) or True)
            assert np.array_equal(np.array(got, dtype=np.int64), want_readable[t]), (t, "readable")
        except BaseException as e:  # noqa: BLE001 - collected and re-raised on the main thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        raise errors[0]
    assert sum(len(w) for w in want) > 1000


# ---------------------------------------------------------------- compact wire format (acgpu_match_utf16_compact)

def _compact_raw(m, hay):
    """(kind, n, records) of one acgpu_match_utf16_compact call."""
    import ctypes as C
    from ahocorasick_b200 import _lib
    from ahocorasick_b200.matchers import _Records
    arr = hay if isinstance(hay, np.ndarray) else np.frombuffer(hay.encode("utf-16-le"), dtype=np.uint16)
    res = _lib.Matches()
    _lib.check(_lib.lib().acgpu_match_utf16_compact(m.handle, arr.ctypes.data if arr.size else None, arr.size, C.byref(res)))
    try:
        kind, n = int(res.kind), int(res.n)
        if kind == _lib.MATCHES_MASKS:
            assert int(res.n_chars) == arr.size
        rec = _Records(res, False)
        return kind, n, rec
    finally:
        _lib.lib().acgpu_free_matches(C.byref(res))


def test_compact_masks_equal_records_and_oracle():
    """Dense AhoCorasickSet streams come back as per-char hit masks (2 B/char on the wire); expanded in the mirror they
    must be the very stream acgpu_match_utf16 returns as records and the oracle's - across 8 Mi-char pipeline chunks,
    odd lengths and unaligned host pointers.  Sparse streams and Maps stay records."""
    from ahocorasick_b200 import _lib
    rng = random.Random(5150)
    kws = ["a", "ab", "ba", "bab", "abba", "b" * 7, "ab" * 8, "a" * 16, "bbbab"]
    gs = ac.AhoCorasickSet(kws, True)
    om = ora.Matcher("ahocorasick", kws)
    big = np.frombuffer("".join(rng.choice("ab ") for _ in range(1 << 16)).encode("utf-16-le"), dtype=np.uint16)
    big = np.tile(big, 300)[: (2 << 23) + 12_347]            # two full pipeline chunks and a ragged third
    big = big.copy()
    big[rng.randrange(big.size)] = ord("b")
    for n, off in ((1, 0), (7, 0), (255, 0), (70_001, 0), (70_001, 1), (70_001, 3), (big.size, 0), (big.size - 5, 5)):
        hay = big[off:off + n]
        kind, cnt, rec = _compact_raw(gs, hay)
        plain = gs.match_records(hay, compact=False)
        first = min(n, 1 << 23)  # the first pipeline chunk decides the wire format
        dense = int((plain.end <= first).sum()) >= 0.25 * first
        assert (kind == _lib.MATCHES_MASKS) == dense, (n, off)
        assert n < 255 or dense
        assert cnt == len(plain) == len(rec), (n, off, cnt, len(plain))
        assert np.array_equal(rec.start, plain.start) and np.array_equal(rec.end, plain.end), (n, off)
        if n <= 70_001:
            want = om.match(hay, cap=4 * n + 16)
            assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"]), (n, off)
    want = om.match(big, cap=2 * big.size)
    rec = gs.match_records(big)
    assert len(rec) == len(want) and np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"])
    # listener replay through the mirror: early stop inside a position's records (longest first)
    hay = "abbab" * 3
    full = [(s, e) for s, e, _ in oracle_stream(om, hay)]
    for k in range(1, len(full) + 2):
        assert gpu_set_stream(gs, hay, stop_after=k) == full[:k]
    # sparse stream: records on the wire
    sparse = ac.AhoCorasickSet(["abbab", "bbbbbbb"], True)
    kind, cnt, rec = _compact_raw(sparse, big[:500_000])
    assert kind == _lib.MATCHES_RECORDS and cnt == len(sparse.match_records(big[:500_000], compact=False))
    # empty haystack, empty dictionary
    kind, cnt, _ = _compact_raw(gs, big[:0])
    assert cnt == 0
    kind, cnt, _ = _compact_raw(ac.AhoCorasickSet([], True), big[:1000])
    assert cnt == 0 and kind == _lib.MATCHES_RECORDS


def test_compact_call_for_maps_and_other_families_is_records():
    from ahocorasick_b200 import _lib
    kws = ["a", "ab", "ba", "bab"]
    hay = "abbab ab" * 5000
    for fam in FAMILIES:
        m = SETS[fam](kws, True)
        kind, cnt, rec = _compact_raw(m, hay)
        want = oracle_stream(ora.Matcher(fam, kws), hay)
        assert (kind == _lib.MATCHES_MASKS) == (fam == "ahocorasick")
        assert list(zip(rec.start.tolist(), rec.end.tolist())) == [(s, e) for s, e, _ in want], fam
    gm = ac.AhoCorasickMap(kws, list(range(4)), True)
    assert gpu_map_stream(gm, hay) == oracle_stream(ora.Matcher("ahocorasick", kws, n_values=4), hay)


# ---------------------------------------------------------------- AhoCorasick outside the narrow-alphabet envelope (kernel_wide.cuh)

def _launches(m):
    from ahocorasick_b200 import _lib
    return _lib.lib().acgpu_launches_per_match(m.handle)


@pytest.mark.parametrize("alphabet,max_kw", [
    ("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ'", 24),          # 53 symbols: class-pair table in shared memory
    ("abcdefghijklmnopqrstuvwxyz", 32),                                      # 26 symbols, keywords longer than 12: 32-bit masks
    ("".join(chr(c) for c in range(0x30, 0x30 + 70)) + "αβγδεжзий", 9),      # 79 symbols incl. Greek / Cyrillic: no pair table
    ("".join(chr(c) for c in range(0x400, 0x400 + 200)), 5),                 # 200 symbols beyond Latin-1
])
def test_wide_path_alphabets_and_boundary_lengths(alphabet, max_kw):
    """Dictionaries the tier path cannot take (more than 31 symbols, keywords of 13..32 chars) run k_wide_mask / k_wide_emit:
    Set and Map streams equal the oracle's for haystack lengths that straddle lane, row and chunk boundaries, for
    case-insensitive matching, end-range shards and the Readable overload."""
    rng = random.Random(len(alphabet) * 1000 + max_kw)
    kws = sorted({_rand_word(rng, alphabet, 1, max_kw) for _ in range(600)})
    kws += [alphabet[0] * k for k in (1, 2, max_kw)] + [alphabet[:max_kw]]
    values = list(range(len(kws)))
    om = ora.Matcher("ahocorasick", kws, n_values=len(kws))
    gs, gm = ac.AhoCorasickSet(kws, True), ac.AhoCorasickMap(kws, values, True)
    assert _launches(gs) in (3, 4) and _launches(gm) in (3, 4)  # mask (tile + tail), scan, emit - not the generation-1 single kernel
    base = [rng.choice(alphabet + "  ") for _ in range(70_001)]
    for _ in range(1500):
        k = rng.choice(kws)
        at = rng.randrange(0, len(base) - len(k))
        base[at:at + len(k)] = k
    base = "".join(base)
    for n in (0, 1, 7, 8, 9, 31, 32, 33, 255, 256, 257, 8191, 8192, 8193, 70_001):
        hay = base[:n]
        want = oracle_stream(om, hay)
        pos, _ = _records(gs, hay)
        assert pos == [(s, e) for s, e, _ in want], (n,)
        pos, val = _records(gm, hay)
        assert pos == [(s, e) for s, e, _ in want] and [int(v) for v in val] == [v for _, _, v in want], (n,)
    # unaligned host pointer, case-insensitive twin, Readable
    arr = np.frombuffer(base.encode("utf-16-le"), dtype=np.uint16)
    for off in (1, 3, 5):
        want = om.match(arr[off:], cap=1 << 20)
        rec = gm.match_records(arr[off:])
        assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"]) and \
            np.array_equal(rec.value.astype(np.int64), want["value"].astype(np.int64))
    oi = ora.Matcher("ahocorasick", kws, n_values=len(kws), case_sensitive=False)
    gi = ac.AhoCorasickMap(kws, values, False)
    mixed = "".join(c.upper() if rng.random() < 0.3 else c for c in base)
    assert gpu_map_stream(gi, mixed) == oracle_stream(oi, mixed)
    c = Collect()
    gm.match(io.StringIO(base), c)
    assert [v[0] for v in c.calls] == [int(r["value"]) for r in om.match(base, readable=True)]


def test_wide_path_dense_rows_and_range_shards():
    """a^1..a^32 over 'aaaa...' emits 32 records per position (rows overflow the staging window of k_wide_emit); end-range
    shards of a resident haystack concatenate to the single stream; a capacity below the total truncates the stream only."""
    import ctypes as C
    import torch
    from ahocorasick_b200 import _lib
    from ahocorasick_b200.sharding import plan_range_shards
    kws = ["a" * i for i in range(1, 33)] + ["ab", "ba" * 9, "b" * 21, "abAB", "Zz"]
    hay = "a" * 2100 + "b" * 40 + ("ab" * 700) + "abABZz" * 50 + "a" * 513
    want = ora.Matcher("ahocorasick", kws, n_values=len(kws)).match(hay, cap=1 << 18)
    gs, gm = ac.AhoCorasickSet(kws, True), ac.AhoCorasickMap(kws, list(range(len(kws))), True)
    assert _launches(gs) in (3, 4)
    rec = gs.match_records(hay)
    assert len(rec) == len(want) > 60_000 and np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"])
    rec = gm.match_records(hay)
    assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.value.astype(np.int64), want["value"].astype(np.int64))
    arr = np.frombuffer(hay.encode("utf-16-le"), dtype=np.uint16)
    d_hay = torch.from_numpy(arr.astype(np.int16)).cuda()
    d_pos = torch.empty((len(want) + 16, 2), dtype=torch.int32, device="cuda")
    want_pos = np.stack([want["start"], want["end"]], axis=1).astype(np.int32)
    lib = _lib.lib()

    def run(lo, hi, cap):
        tot = C.c_int64(0)
        _lib.check(lib.acgpu_match_device(gs.handle, d_hay.data_ptr(), arr.size, lo, hi, d_pos.data_ptr(), None, cap, C.byref(tot), None))
        torch.cuda.synchronize()
        return tot.value, d_pos[:min(tot.value, cap)].cpu().numpy()

    for world in (2, 3, 8):
        parts = [run(sh.emit_from, sh.emit_to, len(want) + 16)[1].copy()
                 for sh in plan_range_shards(arr.size, world, max_len=32, align=1 if world == 3 else 8)]
        got = np.concatenate(parts, axis=0)
        assert got.shape == want_pos.shape and np.array_equal(got, want_pos), world
    total, part = run(0, arr.size, 1000)
    assert total == len(want) and np.array_equal(part, want_pos[:1000])


@pytest.mark.parametrize("scale,n", [(0.1, 2_000_000), (1.0, 4_000_000)])
def test_wide_path_real_dictionary(scale, n):
    """workloads.config(5): 236 000 English-like words (1-24 chars, mixed case, apostrophes - the reference's README quotes a
    235 886-word English dictionary) over English-like text, case-sensitive: Set and Map streams equal the oracle's."""
    c = W.config(5, scale=scale)
    kws = c["keywords"]
    hay = W.make_haystack(c["spec"], n)
    want = ora.Matcher("ahocorasick", kws, n_values=len(kws)).match(hay, cap=2 * n)
    gs = ac.AhoCorasickSet(kws, True)
    assert _launches(gs) in (3, 4) and gs.info()["n_classes"] == 54 and gs.info()["max_len"] == 24
    rec = gs.match_records(hay)
    assert len(rec) == len(want) > n // 8
    assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"])
    rec = ac.AhoCorasickMap(kws, list(range(len(kws))), True).match_records(hay)
    assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"])
    assert np.array_equal(rec.value.astype(np.int64), want["value"].astype(np.int64))


@pytest.mark.parametrize("cap,min_rounds,hand_over", [(1, 0, 4096), (100, 1, 4096), (10_000_000, 0, 4096), (10_000_000, 9, 64)])
def test_wide_path_tail_hand_over_corner_cases(monkeypatch, cap, min_rounds, hand_over):
    """k_wide_tile hands the walks of its thin rounds to k_wide_tail: a full list (the tile finishes its walks itself), a
    hand-over before the first round (every walk goes through k_wide_tail) and a late one give the oracle's stream."""
    monkeypatch.setenv("ACGPU_WIDE_TAIL_CAP", str(cap))
    monkeypatch.setenv("ACGPU_WT_MIN_ROUNDS", str(min_rounds))
    monkeypatch.setenv("ACGPU_WT_HANDOVER", str(hand_over))
    c = W.config(5, scale=0.05)
    kws = c["keywords"]
    n = 300_001
    hay = W.make_haystack(c["spec"], n)
    want = ora.Matcher("ahocorasick", kws, n_values=len(kws)).match(hay, cap=2 * n)
    rec = ac.AhoCorasickSet(kws, True).match_records(hay)
    assert len(rec) == len(want) > n // 8
    assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"])


# ---------------------------------------------------------------- Longest / Shortest: range shards with composed chain maps

@pytest.mark.parametrize("family,is_map", [("longest", False), ("longest", True), ("shortest", False), ("shortest", True)])
@pytest.mark.parametrize("text", ["periodic", "random", "sparse"])
def test_chain_shards_compose_to_the_single_stream(family, is_map, text):
    """SURVEY 8e, general cut: ONE haystack with NO synchronisation point (every char occurs in a keyword; the periodic
    text has interleaved chains that never merge) is cut into runs of whole tiles; every shard scans its own window
    (a separate device buffer holding only [lo, hi + look-ahead)), the 16-entry maps are composed, every shard emits from
    its true entry - the rank-ordered concatenation is the oracle's stream, values included."""
    import ctypes as C
    import torch
    from ahocorasick_b200 import _lib
    from ahocorasick_b200.sharding import (CHAIN_ENTRIES, chain_shard_begin, chain_shard_finish, compose_chain_maps,
                                           plan_chain_shards)
    rng = random.Random(hash((family, is_map, text)) & 0xFFFF)
    n = 8192 * 9 + 1234
    if text == "periodic":
        kws = ["ab", "ba", "aba", "bab", "abab", "b" * 5]
        hay = ("ab" * (n // 2 + 1))[:n]
    elif text == "random":
        kws = sorted({_rand_word(rng, "abc", 1, 9) for _ in range(60)}) + ["a" * 16, "cab" * 5]
        hay = "".join(rng.choice("abc") for _ in range(n))
    else:
        kws = sorted({_rand_word(rng, "abcdefgh", 3, 12) for _ in range(40)})
        base = [rng.choice("abcdefgh") for _ in range(n)]
        for _ in range(n // 40):
            k = rng.choice(kws)
            at = rng.randrange(0, n - len(k))
            base[at:at + len(k)] = k
        hay = "".join(base)
    arr = np.frombuffer(hay.encode("utf-16-le"), dtype=np.uint16)
    want = ora.Matcher(family, kws, n_values=len(kws) if is_map else -1).match(hay, cap=n)
    want_pos = np.stack([want["start"], want["end"]], axis=1).astype(np.int64)
    m = (MAPS if is_map else SETS)[family](*((kws, list(range(len(kws))), True) if is_map else (kws, True)))
    classes, has_other = m.char_classes()
    assert all(classes[ord(c)] != 0 for c in set(hay))            # no synchronisation point anywhere
    tile, look, ent = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    _lib.check(_lib.lib().acgpu_chain_shard_layout(m.handle, C.byref(tile), C.byref(look), C.byref(ent)))
    assert (tile.value, look.value, ent.value) == (8192, 256, CHAIN_ENTRIES)
    cap = len(want) + 16
    for world in (1, 2, 3, 5, 9, 12):
        shards = plan_chain_shards(n, world)
        assert len(shards) == world and shards[0].lo == 0 and shards[-1].hi == n
        handles, maps, bufs = [], [], []
        for sh in shards:
            if sh.empty:
                handles.append(None)
                maps.append(list(range(CHAIN_ENTRIES)))
                continue
            d_win = torch.from_numpy(arr[sh.lo:sh.read_to].astype(np.int16)).cuda()     # the rank's own buffer
            d_map = torch.zeros(CHAIN_ENTRIES, dtype=torch.int64, device="cuda")
            n_dom = sh.hi - sh.lo if sh.hi < n else sh.read_to - sh.lo
            handles.append(chain_shard_begin(m, d_win.data_ptr(), sh.read_to - sh.lo, n_dom, d_map.data_ptr()))
            bufs.append((d_win, d_map))
            torch.cuda.synchronize()
            maps.append(d_map.cpu().tolist())
        entries, firsts = compose_chain_maps(maps)
        pos_parts, val_parts = [], []
        for sh, h, entry, first in zip(shards, handles, entries, firsts):
            if h is None:
                continue
            d_pos = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
            d_val = torch.empty(cap, dtype=torch.int32, device="cuda")
            d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
            chain_shard_finish(h, entry, sh.lo, d_pos.data_ptr(), d_val.data_ptr() if is_map else None, cap, d_tot.data_ptr())
            torch.cuda.synchronize()
            k = int(d_tot.item())
            assert first == sum(p.shape[0] for p in pos_parts)
            pos_parts.append(d_pos[:k].cpu().numpy().astype(np.int64))
            val_parts.append(d_val[:k].cpu().numpy().astype(np.int64))
        got = np.concatenate(pos_parts, axis=0)
        assert got.shape == want_pos.shape and np.array_equal(got, want_pos), (world, family, text)
        if is_map:
            assert np.array_equal(np.concatenate(val_parts), want["value"].astype(np.int64)), world
    assert len(want) > 1000


def test_chain_shards_whole_buffer_pointers_and_refusals():
    """The shards of a haystack that is resident as ONE buffer (pointer = base + 2 * lo: boundaries are multiples of the tile,
    so every window stays aligned), an unaligned base (the first tile is resolved again behind a hidden prefix), and the
    argument checks."""
    import ctypes as C
    import torch
    from ahocorasick_b200 import _lib
    from ahocorasick_b200.sharding import chain_shard_begin, chain_shard_finish, compose_chain_maps, plan_chain_shards
    kws = ["ab", "ba", "aba", "bab", "bbbbabab"]
    n = 8192 * 5 + 77
    hay = ("abb" * (n // 3 + 1))[:n]
    arr = np.frombuffer(hay.encode("utf-16-le"), dtype=np.uint16)
    want = ora.Matcher("longest", kws).match(hay, cap=n)
    want_pos = np.stack([want["start"], want["end"]], axis=1).astype(np.int64)
    m = ac.LongestMatchSet(kws, True)
    cap = len(want) + 16
    for shift in (0, 3):                                   # shift 3: the base pointer is not 16-byte aligned
        d_all = torch.zeros(n + 8, dtype=torch.int16, device="cuda")
        d_all[shift:shift + n] = torch.from_numpy(arr.astype(np.int16)).cuda()
        base = d_all.data_ptr() + 2 * shift
        shards = plan_chain_shards(n, 1 if shift else 4)
        maps, handles = [], []
        for sh in shards:
            d_map = torch.zeros(16, dtype=torch.int64, device="cuda")
            n_dom = sh.hi - sh.lo if sh.hi < n else sh.read_to - sh.lo
            handles.append(chain_shard_begin(m, base + 2 * sh.lo, sh.read_to - sh.lo, n_dom, d_map.data_ptr()))
            torch.cuda.synchronize()
            maps.append(d_map.cpu().tolist())
        entries, _ = compose_chain_maps(maps)
        parts = []
        for sh, h, entry in zip(shards, handles, entries):
            d_pos = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
            d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
            chain_shard_finish(h, entry, sh.lo, d_pos.data_ptr(), None, cap, d_tot.data_ptr())
            torch.cuda.synchronize()
            parts.append(d_pos[:int(d_tot.item())].cpu().numpy().astype(np.int64))
        assert np.array_equal(np.concatenate(parts, axis=0), want_pos), shift
    # an unaligned LAST shard entered at a non-zero offset: the first tile is resolved again behind the hidden prefix
    tail = arr[8192 * 2:].astype(np.int16)
    d_tail = torch.zeros(tail.size + 8, dtype=torch.int16, device="cuda")
    d_tail[5:5 + tail.size] = torch.from_numpy(tail).cuda()
    om = ora.Matcher("longest", kws)
    for entry in (0, 1, 2, 7):
        d_map = torch.zeros(16, dtype=torch.int64, device="cuda")
        h = chain_shard_begin(m, d_tail.data_ptr() + 10, tail.size, tail.size, d_map.data_ptr())
        d_pos = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
        d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
        chain_shard_finish(h, entry, 0, d_pos.data_ptr(), None, cap, d_tot.data_ptr())
        torch.cuda.synchronize()
        w = om.match(arr[8192 * 2 + entry:], cap=n)       # entering at `entry` = the haystack that starts there
        got = d_pos[:int(d_tot.item())].cpu().numpy().astype(np.int64)
        assert np.array_equal(got[:, 0], w["start"].astype(np.int64) + entry) and np.array_equal(got[:, 1], w["end"].astype(np.int64) + entry), entry
    # refusals
    d_map = torch.zeros(16, dtype=torch.int64, device="cuda")
    h = C.c_uint64(0)
    lib = _lib.lib()
    d_all = torch.from_numpy(arr.astype(np.int16)).cuda()
    assert lib.acgpu_chain_shard_begin(m.handle, d_all.data_ptr(), n, 1000, d_map.data_ptr(), C.byref(h), None) == _lib.EINVAL   # not a tile boundary
    assert lib.acgpu_chain_shard_begin(ac.AhoCorasickSet(kws, True).handle, d_all.data_ptr(), n, n, d_map.data_ptr(), C.byref(h), None) == _lib.EINVAL
    wide = ac.LongestMatchSet(["".join(chr(0x41 + i) for i in range(40)), "ab"], True)
    assert lib.acgpu_chain_shard_begin(wide.handle, d_all.data_ptr(), n, n, d_map.data_ptr(), C.byref(h), None) == _lib.EUNSUPPORTED


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("cs", [True, False])
def test_create_from_trie_descriptor(family, cs):
    """acgpu_create(const acgpu_automaton_desc*): a matcher built from the flattened goto trie a Java-side builder hands over
    (trie_desc.flatten_trie restates the reference's constructor loops: null skip, zip, last duplicate wins) gives the
    oracle's ordered stream - nulls, empty keywords, duplicates, a shorter values Iterable, failure links checked."""
    from ahocorasick_b200 import _lib, trie_desc
    fam_id = {"ahocorasick": 0, "longest": 1, "shortest": 2, "wholeword": 3, "wholewordlongest": 4}[family]
    rng = random.Random(hash((family, cs, "desc")) & 0xFFFF)
    for it in range(40):
        alphabet = rng.choice(["ab", "abc", "abAB", "abcdeXYZ"])
        nk = rng.randint(1, 12)
        kws = [_rand_word(rng, alphabet, 1, rng.choice([2, 3, 5, 9])) for _ in range(nk)]
        if rng.random() < 0.4:
            kws.insert(rng.randint(0, len(kws)), rng.choice([None, "", kws[0]]))
        sep = " " if family.startswith("wholeword") or rng.random() < 0.3 else ""
        hay = "".join(rng.choice(alphabet + sep * 2) for _ in range(rng.randint(0, 300)))
        nv = rng.choice([len(kws), max(0, len(kws) - 1)])
        want = oracle_stream(ora.Matcher(family, kws, n_values=nv, case_sensitive=cs), hay)
        trie = trie_desc.flatten_trie(fam_id, kws, list(range(nv)), cs)
        m = MAPS[family].from_trie(trie, list(range(nv)))
        assert gpu_map_stream(m, hay) == want, (kws, hay, cs, nv)
        want_set = [(s, e) for s, e, _ in oracle_stream(ora.Matcher(family, kws, case_sensitive=cs), hay)]
        s = SETS[family].from_trie(trie_desc.flatten_trie(fam_id, kws, None, cs))
        assert gpu_set_stream(s, hay) == want_set, (kws, hay, cs)
    # a larger dictionary through the tier path, and a descriptor the library must refuse
    kws = W.make_keywords(5000, 77)
    hay = W.make_haystack(W.HaystackSpec("lower", 5, kws), 300_000)
    trie = trie_desc.flatten_trie(fam_id, kws, list(range(len(kws))), cs)
    got = MAPS[family].from_trie(trie, list(range(len(kws)))).match_records(hay)
    want = ora.Matcher(family, kws, n_values=len(kws), case_sensitive=cs).match(hay)
    assert np.array_equal(got.start, want["start"]) and np.array_equal(got.end, want["end"])
    assert np.array_equal(got.value.astype(np.int64), want["value"].astype(np.int64))
    trie.fail[len(trie.fail) // 2] ^= 1
    with pytest.raises(_lib.AcgpuError):
        MAPS[family].from_trie(trie, list(range(len(kws))))
