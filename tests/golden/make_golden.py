#!/usr/bin/env python3
"""Regenerate the committed fixtures under tests/golden/.

Two kinds of vectors, kept apart because they do not have the same authority:

* reference_kats.json — transcribed from the reference's own tests and README (the generating "script" for these is
  the transcription in tests/golden_cases.py, each case citing its source line): MatchQueueTest.java:8-57 (the only
  ORDERED known answers the reference holds), the literal inputs of SetTest.java:67-130 / MapTest.java:68-131 (the
  reference asserts their match COUNT against a brute force; the ordered streams stored here are derived from the
  literal semantics and agree with those counts) and README.md:90,96,102,109,124.
* oracle_streams.json — seeded inputs run through the CPU oracle (oracle/ac_oracle.c).  The reference is Java and no
  JVM exists in this image, so these are NOT outputs of the reference itself: they freeze the oracle's behaviour
  (a regression pin for the restatement and a fixture the GPU parity tests can compare against without
  re-deriving anything).  Every case lists dictionary, haystack and the full ordered (start, end, value) stream.

    python tests/golden/make_golden.py          # rewrites both files
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from golden_cases import FAMILIES, ILLEGAL, LITERAL_CASES, MATCH_QUEUE_KATS  # noqa: E402


def reference_kats():
    lit = {}
    for name, (hay, kws, expect) in LITERAL_CASES.items():
        if len(kws) > 1000:  # testFullNode: all 65 536 single-char keywords — described, not listed
            lit[name] = {"haystack": [ord(c) for c in hay], "keywords": "all 65536 single code units", "expect": expect}
        else:
            lit[name] = {"haystack": hay, "keywords": kws, "expect": expect}
    return {"match_queue": MATCH_QUEUE_KATS, "literal": lit, "illegal_marker": ILLEGAL}


def rand_word(rng, alphabet, lo, hi):
    return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))


def seeded_cases():
    """Small dictionaries/haystacks that exercise overlaps, nested keywords, case folding beyond ASCII, duplicates
    (value rules Q6), null/empty keywords, custom word characters and Readable fills."""
    cases = []
    rng = random.Random(20261017)
    # 1: dense overlaps over a tiny alphabet
    kws = sorted({rand_word(rng, "ab", 1, 6) for _ in range(24)})
    hay = "".join(rng.choice("ab ") for _ in range(160))
    cases.append(dict(name="dense_ab", keywords=kws, haystack=hay, case_sensitive=True))
    # 2: lower-case words with planted keywords and separators
    kws = sorted({rand_word(rng, "abcdefghij", 2, 7) for _ in range(60)})
    parts = []
    for _ in range(60):
        parts.append(rng.choice(kws) if rng.random() < 0.5 else rand_word(rng, "abcdefghij", 1, 8))
        parts.append(rng.choice([" ", ", ", "-", "_", ".", ""]))
    cases.append(dict(name="words", keywords=kws, haystack="".join(parts), case_sensitive=True))
    # 3: case-insensitive with Latin-1 / Greek / Cyrillic letters (Character.toLowerCase per UTF-16 unit)
    kws = ["straße", "ÉCOLE", "Ωmega", "привет", "İstanbul", "abc", "ABCd", "σίγμα"]
    hay = "Ecole école ÉCOLE ωMEGA ПРИВЕТ привет i̇stanbul İSTANBUL abcD STRASSE straße ΣΊΓΜΑ σίγμα"
    cases.append(dict(name="case_fold", keywords=kws, haystack=hay, case_sensitive=False))
    # 4: duplicates, null and empty keywords: value rules (last wins; first wins for Shortest; skipped keywords
    #    still consume a value)
    kws = ["ab", None, "", "abc", "ab", "bc", "abc", "c"]
    cases.append(dict(name="duplicates", keywords=kws, haystack="xabcabc ab c", case_sensitive=True))
    # 5: custom word characters (toggle constructor: '_' is no word char, '=' is) — whole-word only matters
    kws = ["a=b", "key", "k_y", "x"]
    cases.append(dict(name="toggle_wordchars", keywords=["a=b", "key", "x"], haystack="a=b key_x x=y key a=b=", case_sensitive=True,
                      word_chars=dict(mode=2, chars=["_", "="], toggles=[False, True])))
    # 6: long nested keywords up to 16 chars
    kws = ["a" * i for i in range(1, 17)] + ["ab" * i for i in range(1, 9)]
    cases.append(dict(name="nested16", keywords=kws, haystack="a" * 20 + "b" + "ab" * 9 + "aab", case_sensitive=True))
    # 7: multi-word keywords (WholeWordLongest: carried fail matches, consumed words are not rescanned; the plain
    #    WholeWord constructor rejects them)
    kws = ["new york", "new", "york city", "new york city hall", "city", "hall of fame", "of", "  padded  ", "x-ray", "x"]
    hay = ("new york city hall, new york city halls new yorker; york city new-york hall of fame of fames "
           "padded x-ray x ray new york city hall")
    cases.append(dict(name="multi_word", keywords=kws, haystack=hay, case_sensitive=False))
    return cases


def main():
    from oracle import oracle as ora
    ora.build()
    json.dump(reference_kats(), open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1, ensure_ascii=True)
    out = []
    for c in seeded_cases():
        entry = dict(c)
        entry["streams"] = {}
        for fam in FAMILIES:
            wc = None
            if c.get("word_chars") and fam.startswith("wholeword"):
                w = c["word_chars"]
                wc = ora.word_chars(w["mode"], w["chars"], w["toggles"])
            try:
                m = ora.Matcher(fam, c["keywords"], n_values=len(c["keywords"]), case_sensitive=c["case_sensitive"], word_chars_table=wc)
            except ora.OracleError as e:
                entry["streams"][fam] = {"error": str(e)}
                continue
            rec = m.match(c["haystack"])
            stream = [[int(r["start"]), int(r["end"]), int(r["value"])] for r in rec]
            rd = [int(r["value"]) for r in m.match(c["haystack"], readable=True)]
            entry["streams"][fam] = {"string": stream, "readable_values": rd}
        out.append(entry)
    json.dump(out, open(os.path.join(HERE, "oracle_streams.json"), "w"), indent=1, ensure_ascii=True)
    print("wrote", len(out), "seeded cases")


if __name__ == "__main__":
    main()
