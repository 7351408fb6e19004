"""Spec-level brute-force models of the four matcher families.

Independent of both the literal oracle (oracle/ac_oracle.c) and the CUDA path: they
work from the *definition* of the result (sets of occurrences), not from any automaton.
Used to cross-check the oracle under random fuzz, mirroring the reference's own test
strategy (brute-force getCorrectCount, reference src/test/.../AhoCorasickTest.java:28-38,
LongestMatchTest.java:30-58) but asserting the full ordered stream, not only the count.
"""
from __future__ import annotations

import unicodedata
from typing import Dict, List, Optional, Sequence, Tuple


def java_lower(ch: str) -> str:
    if "\ud800" <= ch <= "\udfff":
        return ch
    low = ch.lower()
    return low if len(low) == 1 else "i"  # only U+0130 is multi-char


def fold(s: str, cs: bool) -> str:
    return s if cs else "".join(java_lower(c) for c in s)


def is_letter_or_digit(ch: str) -> bool:
    if "\ud800" <= ch <= "\udfff":
        return False
    return unicodedata.category(ch) in ("Lu", "Ll", "Lt", "Lm", "Lo", "Nd")


def default_word_char(ch: str) -> bool:
    return ch in "-_" or is_letter_or_digit(ch)


def effective_dict(keywords: Sequence[Optional[str]], n_values: int, cs: bool, first_wins: bool) -> Dict[str, int]:
    """folded keyword -> value index (Q6: last wins, except Shortest where the first wins)."""
    n = len(keywords) if n_values < 0 else min(len(keywords), n_values)
    d: Dict[str, int] = {}
    for i in range(n):
        k = keywords[i]
        if k is None or len(k) == 0:
            continue
        f = fold(k, cs)
        if first_wins and f in d:
            continue
        d[f] = i if n_values >= 0 else -1
    return d


def occurrences(d: Dict[str, int], hay: str, cs: bool) -> List[Tuple[int, int, int]]:
    h = fold(hay, cs)
    lens = sorted({len(k) for k in d})
    occ = []
    for s in range(len(h)):
        for L in lens:
            if s + L <= len(h):
                v = d.get(h[s:s + L])
                if v is not None:
                    occ.append((s, s + L, v))
    return occ


def ahocorasick(keywords, hay, cs=True, n_values=-1):
    d = effective_dict(keywords, n_values, cs, False)
    return sorted(occurrences(d, hay, cs), key=lambda t: (t[1], t[0]))


def longest(keywords, hay, cs=True, n_values=-1):
    d = effective_dict(keywords, n_values, cs, False)
    occ = occurrences(d, hay, cs)
    best: Dict[int, Tuple[int, int]] = {}
    for s, e, v in occ:
        if s not in best or e > best[s][0]:
            best[s] = (e, v)
    out = []
    pos = 0
    for s in sorted(best):
        if s >= pos:
            e, v = best[s]
            out.append((s, e, v))
            pos = e
    return out


def shortest(keywords, hay, cs=True, n_values=-1):
    """Earliest end, then the longest occurrence with that end that starts at/after pos.

    Value rule (ShortestMatchSet.java:32-36, ShortestMatchMap.java:44-54): a keyword that has
    an *earlier-inserted* keyword as prefix (or is equal to it) is dropped at insertion."""
    n = len(keywords) if n_values < 0 else min(len(keywords), n_values)
    d: Dict[str, int] = {}
    for i in range(n):
        k = keywords[i]
        if k is None or len(k) == 0:
            continue
        f = fold(k, cs)
        if any(f[:j] in d for j in range(1, len(f) + 1)):
            continue
        d[f] = i if n_values >= 0 else -1
    occ = occurrences(d, hay, cs)
    out = []
    pos = 0
    while True:
        cand = [(e, s, v) for s, e, v in occ if s >= pos]
        if not cand:
            break
        e, s, v = min(cand)
        out.append((s, e, v))
        pos = e
    return out


def wholeword(keywords, hay, cs=True, n_values=-1, is_word=default_word_char):
    """Maximal word-char runs that are (folded) keywords. Valid when is_word(fold(c)) == is_word(c)
    for every haystack char (always true for the default table). Raises ValueError like the
    reference's IllegalArgumentException for keywords with inner non-word chars."""
    n = len(keywords) if n_values < 0 else min(len(keywords), n_values)
    d: Dict[str, int] = {}
    for i in range(n):
        k = keywords[i]
        if k is None:
            continue
        idx = [j for j, c in enumerate(k) if is_word(c)]
        if idx:
            k = k[idx[0]:idx[-1] + 1]
        if any(not is_word(c) for c in k):
            raise ValueError(k + " contains non-word characters.")
        if len(k) == 0:
            continue
        d[fold(k, cs)] = i if n_values >= 0 else -1
    h = fold(hay, cs)
    out = []
    i = 0
    nh = len(hay)
    while i < nh:
        if not is_word(hay[i]):
            i += 1
            continue
        j = i
        while j < nh and is_word(hay[j]):
            j += 1
        v = d.get(h[i:j])
        if v is not None:
            out.append((i, j, v))
        i = j
    return out


def wholewordlongest(keywords, hay, cs=True, n_values=-1, is_word=default_word_char):
    """WholeWordLongestMatchSet/Map (WholeWordLongestMatchSet.java:47-182), from the definition: a walk starts at
    position 0 and then at word starts; from a start s it reports the LONGEST keyword that is a prefix of hay[s:] and is
    followed by a non-word char or the end of the input; the walk consumes hay[s:idx] where idx is the first position at
    which hay[s:idx+1] is no longer a prefix of any keyword, and the next walk starts at the first word start after idx
    (words inside the consumed stretch are never rescanned).  Keywords are trimmed, not validated; last duplicate wins.
    Valid when is_word(fold(c)) == is_word(c) for every haystack char."""
    n = len(keywords) if n_values < 0 else min(len(keywords), n_values)
    d: Dict[str, int] = {}
    for i in range(n):
        k = keywords[i]
        if k is None:
            continue
        idx = [j for j, c in enumerate(k) if is_word(c)]
        if idx:
            k = k[idx[0]:idx[-1] + 1]
        if len(k) == 0:
            continue
        d[fold(k, cs)] = i if n_values >= 0 else -1
    prefixes = {k[:j] for k in d for j in range(1, len(k) + 1)}
    h = fold(hay, cs)
    nh = len(hay)
    out = []
    s = 0
    while s < nh:
        L = 0
        while s + L < nh and h[s:s + L + 1] in prefixes:
            L += 1
        for dlen in range(L, 0, -1):
            if h[s:s + dlen] in d and (s + dlen == nh or not is_word(h[s + dlen])):
                out.append((s, s + dlen, d[h[s:s + dlen]]))
                break
        p = s + L + 1
        while p < nh and not (is_word(hay[p]) and not is_word(hay[p - 1])):
            p += 1
        s = p
    return out


def reference_count_wholewordlongest(keywords, hay, is_word=default_word_char):
    """The brute-force count the reference's own test asserts against (WholeWordLongestMatchTest.java:48-66):
    keywords trimmed and sorted longest first, boundaries tested with Character.isLetterOrDigit."""
    kws = []
    for k in keywords:
        idx = [j for j, c in enumerate(k) if is_word(c)]
        kws.append(k[idx[0]:idx[-1] + 1] if idx else k)
    kws.sort(key=lambda k: -len(k))
    count = 0
    i = 0
    nh = len(hay)
    while i < nh:
        for needle in kws:
            e = i + len(needle)
            if (len(needle) > 0 and e <= nh and hay[i:e] == needle and (e == nh or not is_letter_or_digit(hay[e]))
                    and (i == 0 or not is_letter_or_digit(hay[i - 1]))):
                count += 1
                i += len(needle) - 1
                i += 1
                while i < nh and not is_word(hay[i]):
                    i += 1
                i -= 1
                break
        i += 1
    return count


MODELS = {"ahocorasick": ahocorasick, "longest": longest, "shortest": shortest, "wholeword": wholeword,
          "wholewordlongest": wholewordlongest}
