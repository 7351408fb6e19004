"""Pin the CPU oracle (oracle/ac_oracle.c) against every known answer the reference's own tests
hold for the matching path, then against spec-level brute force under random fuzz.
CPU-only (no GPU, no product code)."""
import random

import numpy as np
import pytest

from oracle import oracle as ora
import ac_spec as spec
from golden_cases import FAMILIES, ILLEGAL, LITERAL_CASES, MATCH_QUEUE_KATS


def stream(rec):
    return [(int(r["start"]), int(r["end"])) for r in rec]


def stream_v(rec):
    return [(int(r["start"]), int(r["end"]), int(r["value"])) for r in rec]


@pytest.mark.parametrize("name", sorted(MATCH_QUEUE_KATS))
def test_match_queue_kats(name):
    """MatchQueueTest.java:8-57 — exact ordered (end, length) sequences."""
    kat = MATCH_QUEUE_KATS[name]
    q = ora.MatchQueue()
    got = []
    for step in kat["steps"]:
        if step[0] == "push":
            q.push(step[1], step[2])
        else:
            got += q.match_and_clear(step[1])
    assert got == kat["expect"]


@pytest.mark.parametrize("name", sorted(LITERAL_CASES))
@pytest.mark.parametrize("family", FAMILIES)
def test_literal_cases(name, family):
    """SetTest.java:67-130 literals + README examples, as ordered streams; Map and Readable
    overloads must agree with the Set/String stream (MapTest.java:178-188 asserts equal counts)."""
    hay, kws, expect = LITERAL_CASES[name]
    if family not in expect:
        pytest.skip("no expectation for this family")
    exp = expect[family]
    if exp == ILLEGAL:
        with pytest.raises(ora.OracleError, match="contains non-word characters"):
            ora.Matcher(family, kws)
        return
    m = ora.Matcher(family, kws)
    got = stream(m.match(hay))
    if isinstance(exp, int):
        assert len(got) == exp
    else:
        assert got == exp
    # brute-force model agrees (the reference's own check is count-equality with a brute force)
    assert got == [(s, e) for s, e, _ in spec.MODELS[family](kws, hay)]
    # Map, keywords as values (MapTest.java:133-137): value index points at an equal keyword
    mm = ora.Matcher(family, kws, n_values=len(kws))
    gv = stream_v(mm.match(hay))
    assert [(s, e) for s, e, _ in gv] == got
    for s, e, v in gv:
        assert kws[v] == hay[s:e]
    # Readable overload: same values in the same order
    rv = mm.match(hay, readable=True)
    assert [int(r["value"]) for r in rv] == [v for _, _, v in gv]
    assert all(int(r["start"]) == -1 for r in rv)


def test_empty_inputs():
    for fam in FAMILIES:
        m = ora.Matcher(fam, [])
        assert len(m.match("anything at all")) == 0
        m = ora.Matcher(fam, ["abc", "", None])
        assert len(m.match("")) == 0
        assert stream(m.match("xabcx abc")) in ([(1, 4), (6, 9)], [(6, 9)])


def test_map_zip_and_duplicates():
    """Q6: zip to the shorter iterable; skipped keywords still consume a value; last duplicate wins
    (first wins for Shortest)."""
    kws = ["ab", None, "", "ab", "cd", "zz"]
    for fam in ("ahocorasick", "longest", "wholeword"):
        m = ora.Matcher(fam, kws, n_values=5)
        assert stream_v(m.match("ab cd zz")) == [(0, 2, 3), (3, 5, 4)]
    m = ora.Matcher("shortest", kws, n_values=5)
    assert stream_v(m.match("ab cd zz")) == [(0, 2, 0), (3, 5, 4)]


def test_case_insensitive_and_tables():
    assert ora.lib().ora_to_lower(0x0130) == 0x0069
    assert ora.lib().ora_to_lower(ord("Σ")) == ord("σ")
    assert ora.lib().ora_to_lower(0xD801) == 0xD801
    assert ora.lib().ora_is_letter_or_digit(ord("é")) == 1
    assert ora.lib().ora_is_letter_or_digit(ord("_")) == 0
    m = ora.Matcher("ahocorasick", ["Straße", "ΣΟΦΟΣ", "İi"], case_sensitive=False)
    assert stream(m.match("xSTRAßE σοφοσ iI")) == [(1, 7), (8, 13), (14, 16)]
    # default word chars are closed under toLowerCase, so the maximal-run model is exact
    wc = ora.word_chars(0)
    lower = np.array([ora.lib().ora_to_lower(c) for c in range(65536)])
    assert np.array_equal(wc[lower], wc)


def test_early_stop_quirks():
    """Q1: Shortest re-delivers the match on which the listener said false; Q3: clean truncation."""
    kws = ["ab", "b", "abc", "c"]
    hay = "abcabcabc"
    for fam in ("ahocorasick", "longest"):
        m = ora.Matcher(fam, kws)
        full = stream(m.match(hay))
        for k in range(1, len(full) + 1):
            assert stream(m.match(hay, stop_after=k)) == full[:k]
    m = ora.Matcher("shortest", kws)
    full = stream(m.match(hay))
    assert full == [(0, 2), (2, 3), (3, 5), (5, 6), (6, 8), (8, 9)]
    for k in range(1, len(full)):
        assert stream(m.match(hay, stop_after=k)) == full[:k] + [full[k - 1]]
    assert stream(m.match(hay, stop_after=len(full))) == full  # final match is emitted once, post-loop
    m = ora.Matcher("wholeword", ["ab", "cd"])
    assert stream(m.match("ab cd ab", stop_after=2)) == [(0, 2), (3, 5)]


def test_shortest_readable_fill_boundary_q4():
    """Q4: ShortestMatchMap.match(Readable) re-emits a match that ends exactly on a fill boundary."""
    m = ora.Matcher("shortest", ["ab"], n_values=1)
    hay = "ab" + "x" * 10
    assert [int(r["value"]) for r in m.match_readable_schedule(hay, [])] == [0]
    # boundary right after "ab": emitted at end of fill 1 and again by the first char of fill 2
    assert [int(r["value"]) for r in m.match_readable_schedule(hay, [2, 100])] == [0, 0]
    # a zero-length read re-emits once more
    assert [int(r["value"]) for r in m.match_readable_schedule(hay, [2, 0, 100])] == [0, 0, 0]


def test_wholeword_custom_word_chars():
    """README.md:114-118 toggle constructor; configs[3] of BASELINE.json."""
    wc = ora.word_chars(2, ["_", "="], [False, True])
    m = ora.Matcher("wholeword", ["a=b", "x-y", "_q_"], word_chars_table=wc)
    assert stream(m.match("a=b_x-y q _q_ a=b=")) == [(0, 3), (4, 7), (8, 9), (11, 12)]
    # custom-only constructor: alphanumeric keywords are rejected
    with pytest.raises(ora.OracleError):
        ora.Matcher("wholeword", ["abc"], word_chars_table=ora.word_chars(1, ["_", "="]))


def _rand_word(rng, alphabet, lo, hi):
    return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("cs", [True, False])
def test_fuzz_against_spec(family, cs):
    """Random small dictionaries/haystacks: literal oracle == spec-level brute force (ordered, with values)."""
    rng = random.Random(hash((family, cs)) & 0xFFFF)
    for it in range(400):
        alphabet = rng.choice(["ab", "abc", "abAB", "abcdeXYZ", "aβΒbİi"])
        nk = rng.randint(0, 8)
        kws = [_rand_word(rng, alphabet, 1, rng.choice([2, 3, 5, 9])) for _ in range(nk)]
        if rng.random() < 0.2:
            kws.insert(rng.randint(0, len(kws)), rng.choice([None, ""]))
        sep = " " if family == "wholeword" or rng.random() < 0.3 else ""
        hay = "".join(rng.choice(alphabet + sep * 2) for _ in range(rng.randint(0, 60)))
        nv = rng.choice([-1, len(kws), max(0, len(kws) - 1)])
        m = ora.Matcher(family, kws, n_values=nv, case_sensitive=cs)
        got = stream_v(m.match(hay))
        want = spec.MODELS[family](kws, hay, cs, nv)
        assert got == want, (kws, hay, cs, nv)
        if nv >= 0 and family != "shortest":
            rv = m.match(hay, readable=True)
            assert [int(r["value"]) for r in rv] == [v for _, _, v in want]


def _wwl_case(rng):
    """Dictionaries whose keywords hold non-word chars inside / around them, haystacks of words and punctuation."""
    letters = rng.choice(["ab", "abc", "abAB", "aβΒb"])
    seps = rng.choice([" ", " ,", " -_"])
    kws = []
    for _ in range(rng.randint(0, 8)):
        words = [_rand_word(rng, letters, 1, 3) for _ in range(rng.randint(1, 3))]
        kw = rng.choice(seps).join(words)
        if rng.random() < 0.25:
            kw = rng.choice([" ", ",", ""]) + kw + rng.choice([" ", ", ", ""])
        kws.append(kw)
    if rng.random() < 0.2:
        kws.insert(rng.randint(0, len(kws)), rng.choice([None, "", " ", ","]))
    hay = "".join(rng.choice(letters * 2 + seps) for _ in range(rng.randint(0, 80)))
    return kws, hay


@pytest.mark.parametrize("cs", [True, False])
def test_wholewordlongest_fuzz_against_spec(cs):
    """WholeWordLongest with multi-word keywords: literal oracle == the definition-level model (ordered, with
    values, String and Readable overloads), and the count equals the reference's own brute force
    (WholeWordLongestMatchTest.java:48-66) whenever that brute force applies (case-sensitive, no '-'/'_')."""
    rng = random.Random(7117 + cs)
    for it in range(1500):
        kws, hay = _wwl_case(rng)
        nv = rng.choice([-1, len(kws), max(0, len(kws) - 1)])
        m = ora.Matcher("wholewordlongest", kws, n_values=nv, case_sensitive=cs)
        got = stream_v(m.match(hay))
        want = spec.wholewordlongest(kws, hay, cs, nv)
        assert got == want, (kws, hay, cs, nv)
        if nv >= 0:
            rv = m.match(hay, readable=True)
            assert [int(r["value"]) for r in rv] == [v for _, _, v in want]


def test_wholewordlongest_reference_bruteforce_counts():
    """The reference's test asserts count == its brute force on its literal inputs (SetTest.java:124-130)."""
    for name, (hay, kws, expect) in LITERAL_CASES.items():
        if "wholewordlongest" not in expect or len(kws) > 1000:
            continue
        got = ora.Matcher("wholewordlongest", kws).match(hay)
        assert len(got) == spec.reference_count_wholewordlongest(kws, hay), name


def test_reference_random_unicode_dictionary():
    """Generator.randomStrings(n, 2, 3) style (Generator.java:61-76): length-2 keywords, 50% Latin-1,
    50% anywhere in the BMP — exercises wide alphabets (testFullRandom, SetTest.java:81-89)."""
    rng = random.Random(7)
    kws = list({"".join(chr(rng.randrange(256) if rng.random() < 0.5 else rng.randrange(65536))
                        for _ in range(2)) for _ in range(20000)})
    kws = [k for k in kws if not any(0xD800 <= ord(c) <= 0xDFFF for c in k)]
    hay = "The quick red fox, jumps over the lazy brown dog." + "".join(rng.choice(kws) for _ in range(50))
    for fam in ("ahocorasick", "longest", "shortest"):
        m = ora.Matcher(fam, kws)
        got = stream(m.match(hay))
        want = [(s, e) for s, e, _ in spec.MODELS[fam](kws, hay)]
        assert got == want
        assert len(got) >= 25


# ---------------------------------------------------------------- committed fixtures (tests/golden/)

def _golden(name):
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)))


def test_reference_kats_fixture_in_sync():
    """tests/golden/reference_kats.json is the committed form of golden_cases.py (MatchQueueTest.java:8-57,
    SetTest.java:67-130, README examples); tests/golden/make_golden.py regenerates it."""
    kats = _golden("reference_kats.json")
    assert set(kats["match_queue"]) == set(MATCH_QUEUE_KATS)
    for name, kat in MATCH_QUEUE_KATS.items():
        assert [tuple(x) for x in kats["match_queue"][name]["expect"]] == kat["expect"]
        assert [tuple(x) for x in kats["match_queue"][name]["steps"]] == kat["steps"]
    assert set(kats["literal"]) == set(LITERAL_CASES)
    for name, (hay, kws, expect) in LITERAL_CASES.items():
        got = kats["literal"][name]["expect"]
        assert set(got) == set(expect)
        for fam, exp in expect.items():
            assert got[fam] == ([list(x) for x in exp] if isinstance(exp, list) else exp)


@pytest.mark.parametrize("case", _golden("oracle_streams.json"), ids=lambda c: c["name"])
def test_oracle_reproduces_committed_streams(case):
    """The seeded fixtures freeze the oracle's ordered (start, end, value) streams (String and Readable overloads)."""
    for fam in FAMILIES:
        want = case["streams"][fam]
        wc = None
        if case.get("word_chars") and fam.startswith("wholeword"):
            w = case["word_chars"]
            wc = ora.word_chars(w["mode"], w["chars"], w["toggles"])
        if "error" in want:
            with pytest.raises(ora.OracleError):
                ora.Matcher(fam, case["keywords"], n_values=len(case["keywords"]), case_sensitive=case["case_sensitive"],
                            word_chars_table=wc)
            continue
        m = ora.Matcher(fam, case["keywords"], n_values=len(case["keywords"]), case_sensitive=case["case_sensitive"],
                        word_chars_table=wc)
        assert [list(t) for t in stream_v(m.match(case["haystack"]))] == want["string"]
        assert [int(r["value"]) for r in m.match(case["haystack"], readable=True)] == want["readable_values"]


def test_java_char_tables_cover_all_65536_code_units_like_unicodedata():
    """Appendix B of SURVEY.md: Character.toLowerCase(char) / Character.isLetterOrDigit(char) per UTF-16 code unit, Unicode 15.
    All 65 536 entries of the generated table (oracle/java_char_tables.h; the product's csrc/java_char_tables.h must be the
    same file, and the product's word-character table is checked through the C ABI) against Python's unicodedata."""
    import ctypes as C
    import hashlib
    import os
    import unicodedata
    import numpy as np
    from oracle import oracle as ora
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = open(os.path.join(root, "oracle", "java_char_tables.h"), "rb").read()
    b = open(os.path.join(root, "ahocorasick_b200", "csrc", "java_char_tables.h"), "rb").read()
    assert hashlib.sha256(a).digest() == hashlib.sha256(b).digest()
    assert unicodedata.unidata_version.startswith("15."), "regenerate the expectation for another Unicode version"
    lib = ora.lib()
    changed = 0
    for c in range(65536):
        ch = chr(c)
        if 0xD800 <= c <= 0xDFFF:
            want_lower, want_lod = c, False          # surrogates: unchanged, never letters (quirk Q8)
        else:
            low = ch.lower()
            want_lower = ord(low) if len(low) == 1 else (0x69 if c == 0x130 else c)
            want_lod = unicodedata.category(ch) in ("Lu", "Ll", "Lt", "Lm", "Lo", "Nd")
        assert lib.ora_to_lower(c) == want_lower, hex(c)
        assert bool(lib.ora_is_letter_or_digit(c)) == want_lod, hex(c)
        changed += want_lower != c
    assert changed == 1173  # 1 172 simple mappings + U+0130 (full mapping "i" + U+0307, Java returns U+0069)
    # the product's copy, through the C ABI (host-only call): default word characters = isLetterOrDigit + '-' + '_'
    from ahocorasick_b200 import _lib
    out = np.zeros(65536, np.uint8)
    _lib.check(_lib.lib().acgpu_word_chars(0, None, None, 0, out.ctypes.data))
    want = np.array([1 if (not 0xD800 <= c <= 0xDFFF and unicodedata.category(chr(c)) in ("Lu", "Ll", "Lt", "Lm", "Lo", "Nd")) or c in (0x2D, 0x5F)
                     else 0 for c in range(65536)], dtype=np.uint8)
    assert np.array_equal(out, want) and int(want.sum()) == 49335 + 2
