// C++ tests of the host-side mirror (include/acgpu.hpp) written the way the reference tests its own classes:
// every family runs the same literal inputs (reference: src/test/java/com/roklenarcic/util/strings/SetTest.java:67-130,
// MapTest.java:68-131), a counting listener checks every reported span against the dictionary, and the number of
// matches is compared with a per-family brute-force count (reference: AhoCorasickTest.java:28-38,
// LongestMatchTest.java:29-42, ShortestMatchTest.java:30-42, WholeWordMatchTest.java:71-86,
// WholeWordLongestMatchTest.java:46-65).  On top of the reference's count-only checks this file asserts the ORDERED
// streams of SURVEY.md section 8c, the Map extras (values, Readable count == String count, MapTest.java:178-188),
// the IllegalArgumentException cases (WholeWordMatchTest.java:31-57), early stop and its quirks.
//
//   reference_style_test              needs a CUDA device, runs everything
//   reference_style_test --host-only  host-side checks only (argument validation before device work, tables)
//   reference_style_test --no-device  the same plus: without a CUDA device constructors throw acgpu::Error(ACGPU_ENODEVICE)
//   reference_style_test --dump <family 0-4> <set|map|readable> <cs 0|1> <keywords.u16> <haystack.u16>
//                                     prints "start end [value]" lines (pytest cross-checks them against the oracle);
//                                     keywords.u16 = UTF-16LE keywords separated by U+000A
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "acgpu.hpp"

using namespace acgpu;
using Span = std::pair<int, int>;
using Spans = std::vector<Span>;
using Keywords = std::vector<String>;

static int g_checks = 0, g_failures = 0;
static std::string g_test;

static std::string narrow(const String &s) {
    std::string o;
    for (char16_t c : s) o += (c >= 32 && c < 127) ? (char)c : '?';
    return o.size() > 40 ? o.substr(0, 40) + "..." : o;
}
static std::string show(const Spans &v) {
    std::ostringstream o;
    for (size_t i = 0; i < v.size() && i < 24; i++) o << "(" << v[i].first << "," << v[i].second << ")";
    if (v.size() > 24) o << "... [" << v.size() << "]";
    return o.str();
}
#define EXPECT(cond, msg)                                                                      \
    do {                                                                                       \
        g_checks++;                                                                            \
        if (!(cond)) {                                                                         \
            g_failures++;                                                                      \
            std::cerr << "FAIL " << g_test << " (" << __LINE__ << "): " << msg << std::endl;   \
        }                                                                                      \
    } while (0)

template <class Ex, class F>
static bool throws(F f) {
    try {
        f();
    } catch (const Ex &) {
        return true;
    } catch (...) {
    }
    return false;
}

// ------------------------------------------------------------------------------------------------ families

enum Fam { AC = 0, LONGEST = 1, SHORTEST = 2, WHOLEWORD = 3, WWLONGEST = 4 };
static const char *fam_name[] = {"AhoCorasick", "LongestMatch", "ShortestMatch", "WholeWordMatch", "WholeWordLongestMatch"};

static std::unique_ptr<StringSet> instantiateSet(Fam f, const Keywords &kw, bool cs) {
    switch (f) {
        case AC: return std::make_unique<AhoCorasickSet>(kw, cs);
        case LONGEST: return std::make_unique<LongestMatchSet>(kw, cs);
        case SHORTEST: return std::make_unique<ShortestMatchSet>(kw, cs);
        case WHOLEWORD: return std::make_unique<WholeWordMatchSet>(kw, cs);
        default: return std::make_unique<WholeWordLongestMatchSet>(kw, cs);
    }
}
static std::unique_ptr<StringMap<String>> instantiateMap(Fam f, const Keywords &kw, bool cs) {
    switch (f) {
        case AC: return std::make_unique<AhoCorasickMap<String>>(kw, kw, cs);
        case LONGEST: return std::make_unique<LongestMatchMap<String>>(kw, kw, cs);
        case SHORTEST: return std::make_unique<ShortestMatchMap<String>>(kw, kw, cs);
        case WHOLEWORD: return std::make_unique<WholeWordMatchMap<String>>(kw, kw, cs);
        default: return std::make_unique<WholeWordLongestMatchMap<String>>(kw, kw, cs);
    }
}

static bool at(const String &hay, size_t i, const String &needle) {
    return !needle.empty() && i + needle.size() <= hay.size() && hay.compare(i, needle.size(), needle) == 0;
}

// The reference's "normal count" per family.  For Longest / Shortest the dictionary is ordered longest / shortest
// first and the scan restarts after every hit (valid for the reference's inputs; Shortest is really "earliest end",
// see divergenceProbe below).
static int correctCount(Fam f, Keywords kw, const String &hay, const WordCharacters::Flags &wc) {
    auto word = [&](size_t i) { return wc[(uint16_t)hay[i]] != 0; };
    auto bounded = [&](size_t i, const String &k) {
        return at(hay, i, k) && (i + k.size() == hay.size() || !word(i + k.size())) && (i == 0 || !word(i - 1));
    };
    if (f == LONGEST || f == WWLONGEST)
        std::stable_sort(kw.begin(), kw.end(), [](const String &a, const String &b) { return a.size() > b.size(); });
    if (f == SHORTEST)
        std::stable_sort(kw.begin(), kw.end(), [](const String &a, const String &b) { return a.size() < b.size(); });
    int count = 0;
    if (f == AC) {
        for (const String &k : kw)
            for (size_t i = 0; i + k.size() <= hay.size(); i++) count += at(hay, i, k);
        return count;
    }
    for (size_t i = 0; i < hay.size(); i++) {
        for (const String &k : kw) {
            bool hit = (f == LONGEST || f == SHORTEST) ? at(hay, i, k) : bounded(i, k);
            if (!hit) continue;
            count++;
            if (f != WHOLEWORD) i += k.size() - 1;
            if (f == WWLONGEST) {  // the scan resumes at the next word start
                while (i + 1 < hay.size() && !word(i + 1)) i++;
            }
            break;
        }
    }
    return count;
}

struct Expect {
    bool illegal = false;   // constructor must throw IllegalArgumentException
    bool has_stream = false;
    Spans stream;
};
static Expect S(Spans s) {
    Expect e;
    e.has_stream = true;
    e.stream = std::move(s);
    return e;
}
static Expect ILLEGAL() {
    Expect e;
    e.illegal = true;
    return e;
}

// One reference-style test: every family over (haystack, needles); `ordered` holds the known ordered streams.
static void test(const std::string &name, const String &hay, const Keywords &needles,
                 const std::map<Fam, Expect> &ordered = {}, std::vector<Fam> fams = {AC, LONGEST, SHORTEST, WHOLEWORD, WWLONGEST}) {
    for (Fam f : fams) {
        g_test = name + "/" + fam_name[f];
        auto ex = ordered.find(f);
        if (ex != ordered.end() && ex->second.illegal) {
            EXPECT(throws<IllegalArgumentException>([&] { instantiateSet(f, needles, true); }), "Set constructor must throw");
            EXPECT(throws<IllegalArgumentException>([&] { instantiateMap(f, needles, true); }), "Map constructor must throw");
            continue;
        }
        std::unique_ptr<StringSet> set;
        try {
            set = instantiateSet(f, needles, true);
        } catch (const IllegalArgumentException &e) {
            EXPECT(false, std::string("unexpected IllegalArgumentException: ") + e.what());
            continue;
        }
        WordCharacters::Flags wc = WordCharacters::generateWordCharsFlags();
        Keywords kw = needles;
        if (f == WHOLEWORD || f == WWLONGEST)
            for (String &k : kw) k = WordCharacters::trim(k, wc);

        // counting listener: every reported span is a keyword (and word-bounded for the WholeWord families)
        struct Counting : SetMatchListener {
            Spans got;
            const Keywords *kw;
            const WordCharacters::Flags *wc;
            bool bounded;
            bool match(const String &h, int s, int e) override {
                got.emplace_back(s, e);
                String sub = h.substr(s, e - s);
                EXPECT(std::find(kw->begin(), kw->end(), sub) != kw->end(), "reported span is not a keyword: " << narrow(sub));
                if (bounded) {
                    EXPECT((size_t)e == h.size() || !(*wc)[(uint16_t)h[e]], "span does not end at a word boundary");
                    EXPECT(s == 0 || !(*wc)[(uint16_t)h[s - 1]], "span does not start at a word boundary");
                }
                return true;
            }
        } listener;
        listener.kw = &kw;
        listener.wc = &wc;
        listener.bounded = (f == WHOLEWORD || f == WWLONGEST);
        set->match(hay, listener);

        int normal = correctCount(f, kw, hay, wc);
        EXPECT((int)listener.got.size() == normal, "set found " << listener.got.size() << ", normal matching found " << normal);
        if (ex != ordered.end() && ex->second.has_stream)
            EXPECT(listener.got == ex->second.stream, "ordered stream " << show(listener.got) << " != " << show(ex->second.stream));

        // Map twin: same spans, value == the matched text, Readable count == String count (MapTest.java:178-188)
        auto map = instantiateMap(f, needles, true);
        Spans mgot;
        map->match(hay, [&](const String &h, int s, int e, const String &v) {
            mgot.emplace_back(s, e);
            String want = h.substr(s, e - s);
            String have = (f == WHOLEWORD || f == WWLONGEST) ? WordCharacters::trim(v, wc) : v;
            EXPECT(have == want, "value " << narrow(v) << " is not the matched text " << narrow(want));
            return true;
        });
        EXPECT(mgot == listener.got, "Map stream differs from Set stream: " << show(mgot));
        StringReader reader(hay);
        int rcount = 0;
        map->match(reader, [&](const String &) {
            rcount++;
            return true;
        });
        EXPECT(rcount == (int)mgot.size(), "Readable count " << rcount << " != String count " << mgot.size());
    }
}

// ------------------------------------------------------------------------------------------------ generators

// Deterministic stand-ins for the reference's Generator (random a-z strings, random numbers, combined strings).
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {  // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    int below(int n) { return (int)(next() % (uint64_t)n); }
};
static Keywords distinct(Keywords v) {
    Keywords out;
    std::unordered_set<String> seen;
    for (String &k : v)
        if (seen.insert(k).second) out.push_back(k);
    return out;
}
static Keywords randomStrings(int n, int minLen, int maxLen, uint64_t seed) {
    Rng r(seed);
    Keywords v;
    for (int i = 0; i < n; i++) {
        String k;
        int len = minLen + r.below(maxLen - minLen + 1);
        for (int j = 0; j < len; j++) k += (char16_t)(u'a' + r.below(26));
        v.push_back(k);
    }
    return distinct(v);
}
static Keywords randomNumbers(int n, uint64_t seed) {
    Rng r(seed);
    Keywords v;
    for (int i = 0; i < n; i++) {
        std::string d = std::to_string(r.below(100000));
        v.emplace_back(d.begin(), d.end());
    }
    return distinct(v);
}
static String combined(const Keywords &kw, int n, uint64_t seed) {
    Rng r(seed);
    String s;
    for (int i = 0; i < n; i++) {
        s += kw[r.below((int)kw.size())];
        s += u' ';
    }
    return s;
}
static String repeat(const String &unit, int n) {
    String s;
    for (int i = 0; i < n; i++) s += unit;
    return s;
}

// ------------------------------------------------------------------------------------------------ the tests

static const String FOX = u"The quick red fox, jumps over the lazy brown dog.";
static const Keywords FOX_WORDS = {u"The", u"quick", u"red", u"fox", u"jumps", u"over", u"the", u"lazy", u"brown", u"dog"};
static const Spans FOX_STREAM = {{0, 3}, {4, 9}, {10, 13}, {14, 17}, {19, 24}, {25, 29}, {30, 33}, {34, 38}, {39, 44}, {45, 48}};

static void deviceTests() {
    const std::vector<Fam> plain = {AC, LONGEST, SHORTEST};

    // testEmptyString (SetTest.java:60-64): the random 2-3 letter dictionary is legal for every family here
    test("testEmptyString", u"", randomStrings(10000, 2, 3, 1));
    // testFailureTransitions
    test("testFailureTransitions", u"abbccddeef", {u"bc", u"cc", u"bcc", u"ccddee", u"ccddeee", u"d"},
         {{AC, S({{2, 4}, {2, 5}, {3, 5}, {5, 6}, {6, 7}, {3, 9}})},
          {LONGEST, S({{2, 5}, {5, 6}, {6, 7}})},
          {SHORTEST, S({{2, 4}, {5, 6}, {6, 7}})},
          {WHOLEWORD, S({})}});
    // testFullNode: one keyword per UTF-16 code unit; the WholeWord family refuses the non-word ones
    {
        Keywords all;
        for (int i = 0; i < 65536; i++) all.push_back(String(1, (char16_t)i));
        String hay = {(char16_t)0, (char16_t)0xffff, (char16_t)0xfffe};
        test("testFullNode", hay, all,
             {{AC, S({{0, 1}, {1, 2}, {2, 3}})}, {LONGEST, S({{0, 1}, {1, 2}, {2, 3}})}, {SHORTEST, S({{0, 1}, {1, 2}, {2, 3}})},
              {WHOLEWORD, ILLEGAL()}},
             {AC, LONGEST, SHORTEST, WHOLEWORD});
    }
    // testFullRandom (the WholeWord test classes skip it)
    test("testFullRandom/small", FOX, randomStrings(10000, 2, 3, 2), {}, plain);
    test("testFullRandom/medium", FOX, randomStrings(100000, 2, 3, 3), {}, plain);
    // testLiteral
    test("testLiteral", FOX, FOX_WORDS,
         {{AC, S(FOX_STREAM)}, {LONGEST, S(FOX_STREAM)}, {SHORTEST, S(FOX_STREAM)}, {WHOLEWORD, S(FOX_STREAM)}, {WWLONGEST, S(FOX_STREAM)}});
    // testLongestMatch
    test("testLongestMatch", u"XXXYYZZ", {u"XXX", u"YY", u"XXXYYZZZ"},
         {{AC, S({{0, 3}, {3, 5}})}, {LONGEST, S({{0, 3}, {3, 5}})}, {SHORTEST, S({{0, 3}, {3, 5}})}, {WHOLEWORD, S({})}});
    // testLongKeywords: a^1 .. a^100 over a^100
    {
        Keywords ks;
        for (int i = 1; i <= 100; i++) ks.push_back(repeat(u"a", i));
        Spans shortest;
        for (int i = 0; i < 100; i++) shortest.emplace_back(i, i + 1);
        test("testLongKeywords", repeat(u"a", 100), ks,
             {{LONGEST, S({{0, 100}})}, {SHORTEST, S(shortest)}, {WHOLEWORD, S({{0, 100}})}, {WWLONGEST, S({{0, 100}})}});
    }
    // testOverlap
    test("testOverlap/1", u"aaaa", {u"a", u"aa", u"aaa", u"aaaa"},
         {{AC, S({{0, 1}, {0, 2}, {1, 2}, {0, 3}, {1, 3}, {2, 3}, {0, 4}, {1, 4}, {2, 4}, {3, 4}})},
          {LONGEST, S({{0, 4}})},
          {SHORTEST, S({{0, 1}, {1, 2}, {2, 3}, {3, 4}})},
          {WHOLEWORD, S({{0, 4}})},
          {WWLONGEST, S({{0, 4}})}});
    test("testOverlap/2", u" aaaaaaa aaababababaabaa ", {u"a", u"aa", u"aaa", u"aaaa"},
         {{LONGEST, S({{1, 5}, {5, 8}, {9, 12}, {13, 14}, {15, 16}, {17, 18}, {19, 21}, {22, 24}})}, {WHOLEWORD, S({})}});
    // testShortestMatch
    {
        Keywords nums = randomNumbers(1000, 4);
        test("testShortestMatch/numbers", combined(nums, 50, 5), nums);
        test("testShortestMatch/literal", u"abcyyyy", {u"abcd", u"bcxxxx", u"cyyyy"},
             {{AC, S({{2, 7}})}, {LONGEST, S({{2, 7}})}, {SHORTEST, S({{2, 7}})}});
    }
    // testWholeWordLongest: "as if" holds a non-word char -> WholeWordMatch refuses it (WholeWordMatchTest.java:51-55)
    test("testWholeWordLongest/1", u"as if", {u"as", u"if", u"as if"},
         {{AC, S({{0, 2}, {0, 5}, {3, 5}})}, {LONGEST, S({{0, 5}})}, {SHORTEST, S({{0, 2}, {3, 5}})}, {WHOLEWORD, ILLEGAL()},
          {WWLONGEST, S({{0, 5}})}});
    test("testWholeWordLongest/2", u"ax if", {u"as", u"if", u"as if"}, {{WHOLEWORD, ILLEGAL()}, {WWLONGEST, S({{3, 5}})}});
    test("testWholeWordLongest/3", u"as in", {u"as", u"if", u"as if"}, {{WHOLEWORD, ILLEGAL()}, {WWLONGEST, S({{0, 2}})}});
    test("testWholeWordLongest/4", u"123 4x 1234 5x 1234 56 123 45 1x 345 12 34x 12 345x 123xb 1234 56s",
         {u"123", u"123 45", u"1234 56", u"12 345"}, {{WWLONGEST, S({{0, 3}, {15, 22}, {23, 29}})}}, {WWLONGEST});
    test("testWholeWordLongest/5", u"abc 12", {u"abc", u"abc 123"}, {{WWLONGEST, S({{0, 3}})}}, {WWLONGEST});

    // README worked examples (README.md:90-124 of the reference)
    test("readme/longest", u"a1b2c3d4", {u"b", u"b2", u"2c3d4"}, {{LONGEST, S({{2, 4}})}}, {LONGEST});
    test("readme/shortest1", u"a1b2c3d4", {u"2", u"b2", u"2c3d4"}, {{SHORTEST, S({{2, 4}})}}, {SHORTEST});
    test("readme/shortest2", u"a1b2c3d4", {u"b", u"2", u"b2"}, {{SHORTEST, S({{2, 3}, {3, 4}})}}, {SHORTEST});
    test("readme/wholeword", u"late evening", {u"la", u"late", u"eve", u"evening"}, {{WHOLEWORD, S({{0, 4}, {5, 12}})}}, {WHOLEWORD});

    // Shortest is "earliest end", not "leftmost start" (SURVEY.md A.3): the brute force of the reference's test would say 1
    g_test = "divergenceProbe";
    {
        ShortestMatchSet s(Keywords{u"abcd", u"bc", u"d"}, true);
        Spans got;
        s.match(u"abcd", [&](const String &, int a, int b) {
            got.emplace_back(a, b);
            return true;
        });
        EXPECT(got == Spans({{1, 3}, {3, 4}}), show(got));
    }

    // Case-insensitive: Character.toLowerCase per UTF-16 unit on both sides (AhoCorasickSet.java:33,229)
    g_test = "caseInsensitive";
    {
        AhoCorasickMap<int> m(Keywords{u"he", u"SHE", u"hers", u"Über", u"Σι"}, std::vector<int>{1, 2, 3, 4, 5}, false);
        std::vector<std::tuple<int, int, int>> got;
        m.match(u"uSHErs üBER σΙ", [&](const String &, int s, int e, const int &v) {
            got.emplace_back(s, e, v);
            return true;
        });
        std::vector<std::tuple<int, int, int>> want = {{1, 4, 2}, {2, 4, 1}, {2, 6, 3}, {7, 11, 4}, {12, 14, 5}};
        EXPECT(got == want, "case-insensitive Map stream differs (" << got.size() << " records)");
        AhoCorasickSet cs(Keywords{u"he"}, true);
        int n = 0;
        cs.match(u"HE he He", [&](const String &, int, int) { return ++n, true; });
        EXPECT(n == 1, "case-sensitive matcher found " << n);
    }

    // Map construction rules: zip to the shorter Iterable, null/empty keywords consume a value, last duplicate wins
    // (AhoCorasickMap.java:32-36,49-50); ShortestMatchMap keeps the first duplicate (ShortestMatchMap.java:44-54)
    g_test = "mapConstruction";
    {
        std::vector<std::optional<String>> kws = {String(u"ab"), std::nullopt, String(u""), String(u"cd"), String(u"ab"), String(u"zz")};
        std::vector<int> vals = {10, 11, 12, 13, 14};  // "zz" has no value -> not in the dictionary
        AhoCorasickMap<int> m(kws, vals, true);
        std::vector<int> got;
        m.match(u"ab cd zz", [&](const String &, int, int, const int &v) { return got.push_back(v), true; });
        EXPECT(got == std::vector<int>({14, 13}), "AhoCorasickMap values");
        ShortestMatchMap<int> sm(kws, vals, true);
        got.clear();
        sm.match(u"ab cd zz", [&](const String &, int, int, const int &v) { return got.push_back(v), true; });
        EXPECT(got == std::vector<int>({10, 13}), "ShortestMatchMap values");
        std::vector<const char16_t *> raw = {u"ab", nullptr, u"cd"};
        AhoCorasickSet s(raw, true);
        int n = 0;
        s.match(u"abcd", [&](const String &, int, int) { return ++n, true; });
        EXPECT(n == 2, "null keyword skipped, found " << n);
        AhoCorasickSet empty(Keywords{}, true);  // an empty dictionary is legal and matches nothing
        n = 0;
        empty.match(u"anything", [&](const String &, int, int) { return ++n, true; });
        EXPECT(n == 0, "empty dictionary found " << n);
    }

    // Early stop: a false return ends the scan (README.md:70); Shortest delivers the refused match once more unless it
    // ends at the end of the haystack (ShortestMatchSet.java:204-206,223-226)
    g_test = "earlyStop";
    {
        const Keywords ks = {u"a", u"aa", u"aaa", u"aaaa"};
        auto calls = [&](const StringSet &s, const String &hay, int stop_after) {
            Spans got;
            s.match(hay, [&](const String &, int a, int b) {
                got.emplace_back(a, b);
                return (int)got.size() < stop_after;
            });
            return got;
        };
        EXPECT(calls(AhoCorasickSet(ks, true), u"aaaa", 3) == Spans({{0, 1}, {0, 2}, {1, 2}}), "AhoCorasick early stop");
        EXPECT(calls(LongestMatchSet(ks, true), u"aaaa aaaa", 1) == Spans({{0, 4}}), "Longest early stop");
        EXPECT(calls(ShortestMatchSet(ks, true), u"aaaa", 2) == Spans({{0, 1}, {1, 2}, {1, 2}}), "Shortest early stop re-delivers (Q1)");
        EXPECT(calls(ShortestMatchSet(ks, true), u"aaaa", 4) == Spans({{0, 1}, {1, 2}, {2, 3}, {3, 4}}), "Shortest stop on the last match (Q2)");
        EXPECT(calls(WholeWordMatchSet(ks, true), u"aa aaa a", 2) == Spans({{0, 2}, {3, 6}}), "WholeWord early stop");
        AhoCorasickMap<String> m(ks, ks, true);
        StringReader in(u"aaaa");
        int n = 0;
        m.match(in, [&](const String &) { return ++n < 4; });
        EXPECT(n == 4, "Readable early stop after " << n);
    }

    // Custom word characters (WholeWordMatchSet.java:21-45; BASELINE configs[3]): '_' becomes a separator, '=' a word char
    g_test = "customWordChars";
    {
        std::vector<char16_t> chars = {u'_', u'='};
        std::vector<bool> toggles = {false, true};
        WholeWordMatchSet s(Keywords{u"a=b", u"key"}, true, chars, toggles);
        EXPECT(!s.getWordChars()[u'_'] && s.getWordChars()[u'='] && s.getWordChars()[u'-'] && s.getWordChars()[u'k'], "toggled table");
        Spans got;
        s.match(u"key_a=b key=a=b _key_", [&](const String &, int a, int b) { return got.emplace_back(a, b), true; });
        EXPECT(got == Spans({{0, 3}, {4, 7}, {17, 20}}), show(got));
        WholeWordMatchSet only(Keywords{u"ab"}, true, std::vector<char16_t>{u'a', u'b'});
        got.clear();
        only.match(u"ab cab abb xabx", [&](const String &, int a, int b) { return got.emplace_back(a, b), true; });
        EXPECT(got == Spans({{0, 2}, {4, 6}, {12, 14}}), show(got));
        EXPECT(throws<IllegalArgumentException>([&] { WholeWordMatchSet(Keywords{u"a_b"}, true, chars, toggles); }), "a_b must be refused");
        WholeWordMatchSet thr(Keywords{u"key"}, true, RangeNodeThreshold());  // trailing Thresholder overloads compile and are ignored
        got.clear();
        thr.match(u"key", [&](const String &, int a, int b) { return got.emplace_back(a, b), true; });
        EXPECT(got == Spans({{0, 3}}), show(got));
    }

    // Matchers own their device automaton: movable (containers, factories), not copyable
    g_test = "moveSemantics";
    {
        static_assert(std::is_move_constructible<AhoCorasickSet>::value && !std::is_copy_constructible<AhoCorasickSet>::value, "Set");
        static_assert(std::is_move_constructible<WholeWordMatchMap<int>>::value && !std::is_copy_constructible<WholeWordMatchMap<int>>::value, "Map");
        std::vector<LongestMatchSet> sets;
        sets.push_back(LongestMatchSet(Keywords{u"ab", u"abc"}, true));
        sets.push_back(LongestMatchSet(Keywords{u"x"}, true));
        LongestMatchSet moved = std::move(sets[0]);
        Spans got;
        moved.match(u"abcab", [&](const String &, int a, int b) { return got.emplace_back(a, b), true; });
        EXPECT(got == Spans({{0, 3}, {3, 5}}), show(got));
        WholeWordMatchMap<int> a(Keywords{u"key"}, std::vector<int>{7}, true), b(Keywords{u"other"}, std::vector<int>{9}, true);
        b = std::move(a);
        int v = 0;
        b.match(u"a key", [&](const String &, int, int, const int &x) { return v = x, true; });
        EXPECT(v == 7 && b.getWordChars()[u'k'], "moved-into Map delivers " << v);
    }

    // Readable in many fills: a 300 000-char stream crosses charBufferSize fills and one device block boundary
    g_test = "readableLong";
    {
        Keywords ks = randomStrings(2000, 3, 6, 6);
        String hay = combined(ks, 60000, 7);
        for (Fam f : {AC, LONGEST, SHORTEST, WHOLEWORD, WWLONGEST}) {
            auto map = instantiateMap(f, ks, true);
            std::vector<String> a, b;
            map->match(hay, [&](const String &, int, int, const String &v) { return a.push_back(v), true; });
            struct Small : Readable {  // 1000-char reads: fills are ragged relative to charBufferSize
                StringReader in;
                explicit Small(const String &s) : in(s) {}
                int read(char16_t *d, int cap) override { return in.read(d, std::min(cap, 1000)); }
            } in(hay);
            struct L : ReadableMatchListener<String> {
                std::vector<String> *out;
                bool match(const String &v) override { return out->push_back(v), true; }
            } l;
            l.out = &b;
            map->match(in, l);
            if (f == SHORTEST) {
                // quirk Q4 may deliver a match ending on a fill boundary twice: drop immediate repeats on both sides
                a.erase(std::unique(a.begin(), a.end()), a.end());
                b.erase(std::unique(b.begin(), b.end()), b.end());
            }
            EXPECT(a == b, fam_name[f] << ": Readable values (" << b.size() << ") differ from String values (" << a.size() << ")");
            EXPECT(a.size() > 10000, fam_name[f] << ": only " << a.size() << " matches");
        }
    }
}

static void hostTests(bool have_device) {
    g_test = "wordCharacters";
    {
        auto d = WordCharacters::generateWordCharsFlags();
        EXPECT(d.size() == 65536 && d[u'a'] && d[u'Z'] && d[u'7'] && d[u'-'] && d[u'_'] && !d[u' '] && !d[u'='] && d[0x03b1] && !d[0x2028],
               "default table");
        auto c = WordCharacters::generateWordCharsFlags(std::vector<char16_t>{u'x', u'='});
        int n = 0;
        for (uint8_t f : c) n += f;
        EXPECT(n == 2 && c[u'x'] && c[u'='], "custom-only table");
        auto t = WordCharacters::generateWordCharsFlags(std::vector<char16_t>{u'_', u'='}, std::vector<bool>{false, true});
        EXPECT(!t[u'_'] && t[u'='] && t[u'a'], "toggled table");
        EXPECT(throws<std::out_of_range>([] { WordCharacters::generateWordCharsFlags(std::vector<char16_t>{u'a', u'b'}, std::vector<bool>{true}); }),
               "short toggle array");
        EXPECT(WordCharacters::trim(u"  as if. ", d) == u"as if", "trim");
        EXPECT(WordCharacters::trim(u" .. ", d) == u" .. ", "trim of a keyword without word chars returns it unchanged");
    }
    g_test = "rangeNodeThreshold";
    {
        RangeNodeThreshold t;
        EXPECT(t.isOverThreshold(1, 0, 8), "intervals of 8 or less are always range nodes");
        EXPECT(!t.isOverThreshold(2, 1, 1000) && t.isOverThreshold(900, 1, 1000), "density rule");
    }
    // WholeWordMatchTest.java:45-49 testKeywordsWithNWCRejection and friends: validation happens before any device work
    g_test = "illegalArguments";
    {
        EXPECT(throws<IllegalArgumentException>([] { WholeWordMatchSet(Keywords{u"A B"}, true); }), "A B");
        EXPECT(throws<IllegalArgumentException>([] { WholeWordMatchMap<int>(Keywords{u"A B"}, std::vector<int>{1}, true); }), "A B (Map)");
        try {
            WholeWordMatchSet(Keywords{u"fine", u" as if "}, true);
            EXPECT(false, "no exception");
        } catch (const IllegalArgumentException &e) {
            EXPECT(std::string(e.what()) == "as if contains non-word characters.", "message: " << e.what());
        }
    }
    if (!have_device) {
        g_test = "noDevice";
        try {
            AhoCorasickSet s(Keywords{u"a"}, true);
            EXPECT(false, "constructor succeeded without a device");
        } catch (const Error &e) {
            EXPECT(e.code == ACGPU_ENODEVICE, "code " << e.code << ": " << e.what());
        }
    }
}

// ------------------------------------------------------------------------------------------------ --dump

static String readU16(const char *path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot read ") + path);
    std::string b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    String s(b.size() / 2, u'\0');
    std::memcpy(&s[0], b.data(), s.size() * 2);
    return s;
}

static int dump(int fam, const std::string &kind, bool cs, const char *kwPath, const char *hayPath) {
    String all = readU16(kwPath), hay = readU16(hayPath);
    Keywords kw;
    size_t b = 0;
    for (size_t i = 0; i <= all.size(); i++)
        if (i == all.size() || all[i] == u'\n') {
            if (i > b || i < all.size()) kw.push_back(all.substr(b, i - b));
            b = i + 1;
        }
    std::vector<int> idx(kw.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
    if (kind == "set") {
        auto s = instantiateSet((Fam)fam, kw, cs);
        s->match(hay, [](const String &, int a, int e) { return std::printf("%d %d\n", a, e), true; });
        return 0;
    }
    std::unique_ptr<StringMap<int>> m;
    switch (fam) {
        case AC: m = std::make_unique<AhoCorasickMap<int>>(kw, idx, cs); break;
        case LONGEST: m = std::make_unique<LongestMatchMap<int>>(kw, idx, cs); break;
        case SHORTEST: m = std::make_unique<ShortestMatchMap<int>>(kw, idx, cs); break;
        case WHOLEWORD: m = std::make_unique<WholeWordMatchMap<int>>(kw, idx, cs); break;
        default: m = std::make_unique<WholeWordLongestMatchMap<int>>(kw, idx, cs); break;
    }
    if (kind == "map") {
        m->match(hay, [](const String &, int a, int e, const int &v) { return std::printf("%d %d %d\n", a, e, v), true; });
    } else {
        StringReader in(hay);
        m->match(in, [](const int &v) { return std::printf("%d\n", v), true; });
    }
    return 0;
}

int main(int argc, char **argv) {
    try {
        if (argc >= 2 && std::string(argv[1]) == "--dump") {
            if (argc != 7) return std::fprintf(stderr, "usage: --dump family set|map|readable cs keywords.u16 haystack.u16\n"), 2;
            return dump(std::atoi(argv[2]), argv[3], std::atoi(argv[4]) != 0, argv[5], argv[6]);
        }
        std::string mode = argc >= 2 ? argv[1] : "";
        hostTests(mode != "--no-device");
        if (mode != "--no-device" && mode != "--host-only") deviceTests();
    } catch (const std::exception &e) {
        std::cerr << "ERROR in " << g_test << ": " << e.what() << std::endl;
        return 3;
    }
    std::printf("%d checks, %d failures\n", g_checks, g_failures);
    return g_failures ? 1 : 0;
}
