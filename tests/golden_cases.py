"""Known-answer vectors for the matching path.

Sources (reference paths relative to /root/reference):
* ``MATCH_QUEUE_KATS`` — src/test/java/com/roklenarcic/util/strings/MatchQueueTest.java:8-57,
  the only *ordered* known answers in the reference's test-suite.
* ``LITERAL_CASES`` — the literal inputs of SetTest.java:67-130 / MapTest.java:68-131 and the
  README worked examples (README.md:90,96,102,109,124).  The reference asserts only the match
  COUNT for these (equal to a brute-force count, SetTest.java:191); the ordered streams below
  are derived from the literal semantics and agree with those counts.
"""

# (pushes [(length, idx)...] / flush steps) -> expected [(end, length)...] over all flushes
MATCH_QUEUE_KATS = {
    "testMatchQueue": dict(
        steps=[("push", 3, 3), ("push", 3, 6), ("push", 3, 9), ("push", 9, 10), ("flush", 10),
               ("push", 7, 10), ("flush", 10)],
        expect=[(3, 3), (6, 3), (9, 3), (10, 7)]),
    "testMatchQueueExtendingOverlap": dict(
        steps=[("push", 3, 3), ("push", 4, 4), ("push", 2, 5), ("flush", 4)],
        expect=[(4, 4)]),
    "testMatchQueueSimple": dict(
        steps=[("push", 3, 3), ("push", 2, 3), ("push", 2, 4), ("push", 2, 5), ("flush", 5)],
        expect=[(3, 3), (5, 2)]),
    "testPartialClear": dict(
        steps=[("push", 3, 3), ("push", 3, 6), ("push", 3, 9), ("push", 9, 10), ("flush", 4),
               ("push", 7, 10), ("flush", 10)],
        expect=[(3, 3), (10, 7)]),
}

FOX = "The quick red fox, jumps over the lazy brown dog."
FOX_WORDS = ["The", "quick", "red", "fox", "jumps", "over", "the", "lazy", "brown", "dog"]
FOX_STREAM = [(0, 3), (4, 9), (10, 13), (14, 17), (19, 24), (25, 29), (30, 33), (34, 38), (39, 44), (45, 48)]

A4 = ["a", "aa", "aaa", "aaaa"]
A100 = ["a" * i for i in range(1, 101)]

ILLEGAL = "IllegalArgumentException"

# name -> (haystack, keywords, {family: expected ordered [(start,end)...] | count | ILLEGAL})
LITERAL_CASES = {
    "testFailureTransitions": ("abbccddeef", ["bc", "cc", "bcc", "ccddee", "ccddeee", "d"], {
        "ahocorasick": [(2, 4), (2, 5), (3, 5), (5, 6), (6, 7), (3, 9)],
        "longest": [(2, 5), (5, 6), (6, 7)],
        "shortest": [(2, 4), (5, 6), (6, 7)],
        "wholeword": []}),
    "testLiteral": (FOX, FOX_WORDS, {
        "ahocorasick": FOX_STREAM, "longest": FOX_STREAM, "shortest": FOX_STREAM, "wholeword": FOX_STREAM,
        "wholewordlongest": FOX_STREAM}),
    "testLongestMatch": ("XXXYYZZ", ["XXX", "YY", "XXXYYZZZ"], {
        "ahocorasick": [(0, 3), (3, 5)], "longest": [(0, 3), (3, 5)], "shortest": [(0, 3), (3, 5)],
        "wholeword": []}),
    "testOverlap1": ("aaaa", A4, {
        "ahocorasick": [(0, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 3), (0, 4), (1, 4), (2, 4), (3, 4)],
        "longest": [(0, 4)],
        "shortest": [(0, 1), (1, 2), (2, 3), (3, 4)],
        "wholeword": [(0, 4)], "wholewordlongest": [(0, 4)]}),
    "testOverlap2": (" aaaaaaa aaababababaabaa ", A4, {
        "ahocorasick": 37,
        "longest": [(1, 5), (5, 8), (9, 12), (13, 14), (15, 16), (17, 18), (19, 21), (22, 24)],
        "shortest": 17,
        "wholeword": []}),
    "testShortestMatch2": ("abcyyyy", ["abcd", "bcxxxx", "cyyyy"], {
        "ahocorasick": [(2, 7)], "longest": [(2, 7)], "shortest": [(2, 7)]}),
    "testLongKeywords": ("a" * 100, A100, {
        "ahocorasick": 5050, "longest": [(0, 100)],
        "shortest": [(i, i + 1) for i in range(100)], "wholeword": [(0, 100)], "wholewordlongest": [(0, 100)]}),
    "readmeLongest": ("a1b2c3d4", ["b", "b2", "2c3d4"], {"longest": [(2, 4)]}),
    "readmeShortest1": ("a1b2c3d4", ["2", "b2", "2c3d4"], {"shortest": [(2, 4)]}),
    "readmeShortest2": ("a1b2c3d4", ["b", "2", "b2"], {"shortest": [(2, 3), (3, 4)]}),
    "readmeWholeWord": ("late evening", ["la", "late", "eve", "evening"], {"wholeword": [(0, 4), (5, 12)]}),
    "testWholeWordLongest1": ("as if", ["as", "if", "as if"], {
        "ahocorasick": [(0, 2), (0, 5), (3, 5)], "longest": [(0, 5)], "shortest": [(0, 2), (3, 5)],
        "wholeword": ILLEGAL, "wholewordlongest": [(0, 5)]}),
    # SetTest.java:124-130 (testWholeWordLongest) and README.md:124; the reference asserts the counts
    "testWholeWordLongest2": ("ax if", ["as", "if", "as if"], {"wholewordlongest": [(3, 5)]}),
    "testWholeWordLongest3": ("as in", ["as", "if", "as if"], {"wholewordlongest": [(0, 2)]}),
    "testWholeWordLongest4": ("123 4x 1234 5x 1234 56 123 45 1x 345 12 34x 12 345x 123xb 1234 56s",
                              ["123", "123 45", "1234 56", "12 345"], {
        "wholewordlongest": [(0, 3), (15, 22), (23, 29)]}),
    "testWholeWordLongest5": ("abc 12", ["abc", "abc 123"], {"wholewordlongest": [(0, 3)]}),
    "readmeWholeWordLongest": ("as of", ["as", "if", "as if"], {"wholewordlongest": [(0, 2)]}),
    "divergenceProbe": ("abcd", ["abcd", "bc", "d"], {"shortest": [(1, 3), (3, 4)], "longest": [(0, 4)]}),
    "testFullNode": ("\u0000\uffff\ufffe", [chr(i) for i in range(65536)], {
        "ahocorasick": [(0, 1), (1, 2), (2, 3)], "longest": [(0, 1), (1, 2), (2, 3)],
        "shortest": [(0, 1), (1, 2), (2, 3)], "wholeword": ILLEGAL}),
}

FAMILIES = ("ahocorasick", "longest", "shortest", "wholeword", "wholewordlongest")
