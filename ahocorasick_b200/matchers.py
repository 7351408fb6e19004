"""Host-side mirror of the reference's public API (Java package com.roklenarcic.util.strings).

Class names, constructor argument order, listener contract (``True`` continues, ``False`` stops)
and error behaviour follow the reference; citations are relative to
/root/reference/src/main/java/com/roklenarcic/util/strings/.  Construction flattens the dictionary
inside libacgpu.so and uploads it once; ``match`` runs the CUDA kernels and this module replays the
ordered records to the listener, reproducing the reference's early-stop quirks.

Positions are UTF-16 code-unit offsets, exactly like Java ``String`` indices (for BMP-only text
they equal Python ``str`` indices).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._lib import IllegalArgumentException, check


# ----------------------------------------------------------------------------- listener interfaces

class SetMatchListener:
    """SetMatchListener.java:3-8 — ``match(haystack, startPosition, endPosition) -> bool``."""

    def match(self, haystack, startPosition: int, endPosition: int) -> bool:  # pragma: no cover - interface
        raise NotImplementedError


class MapMatchListener:
    """MapMatchListener.java:3-8 — ``match(haystack, startPosition, endPosition, value) -> bool``."""

    def match(self, haystack, startPosition: int, endPosition: int, value) -> bool:  # pragma: no cover
        raise NotImplementedError


class ReadableMatchListener:
    """ReadableMatchListener.java:4-8 — ``match(value) -> bool`` (values only, no positions)."""

    def match(self, value) -> bool:  # pragma: no cover - interface
        raise NotImplementedError


def _callable_of(listener) -> Callable[..., Any]:
    m = getattr(listener, "match", None)
    if m is not None:
        return m
    if callable(listener):
        return listener
    raise TypeError("listener must be callable or have a match(...) method")


class Thresholder:
    """threshold/Thresholder.java:3-5.  Only shapes the reference's node objects (memory/speed, never
    results); accepted and ignored by the GPU build."""

    def isOverThreshold(self, nodeSize: int, nodeLevel: int, keyIntervalSize: int) -> bool:  # pragma: no cover
        raise NotImplementedError


class RangeNodeThreshold(Thresholder):
    """threshold/RangeNodeThreshold.java:7-29 (kept for API compatibility)."""

    def __init__(self, exponent: float = 1, linearFactor: float = 1, maxValue: float = 0.65, constantFactor: float = 2):
        self.exponent, self.linearFactor, self.maxValue, self.constantFactor = exponent, linearFactor, maxValue, constantFactor

    def isOverThreshold(self, nodeSize: int, nodeLevel: int, keyIntervalSize: int) -> bool:
        if keyIntervalSize <= 8:
            return True
        charArrayCost = (nodeSize // 4) + 3
        return nodeSize + charArrayCost > keyIntervalSize * (
            self.maxValue - self.linearFactor / pow(self.constantFactor + nodeLevel, self.exponent))


# ----------------------------------------------------------------------------- helpers

def _utf16(s) -> np.ndarray:
    if isinstance(s, np.ndarray):
        if s.dtype != np.uint16:
            raise TypeError("haystack arrays must be uint16 UTF-16 code units")
        return np.ascontiguousarray(s)
    if isinstance(s, str):
        return np.frombuffer(s.encode("utf-16-le", "surrogatepass"), dtype=np.uint16)
    raise TypeError("haystack must be str or a uint16 numpy array, not %r" % type(s).__name__)


def _pack_keywords(keywords: Iterable) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """Flatten an Iterable of keywords (str / uint16 arrays / None) into (chars, offsets, is_null, n) for
    acgpu_create_from_keywords.  All-``str`` dictionaries are joined and encoded once (a 1M-keyword dictionary packs in
    ~0.2 s instead of ~2 s); UTF-16 lengths differ from ``len(str)`` only when a keyword holds non-BMP characters, which
    the size check detects."""
    kws = keywords if isinstance(keywords, (list, tuple)) else list(keywords)
    n = len(kws)
    if n and all(type(k) is str for k in kws):
        joined = "".join(kws).encode("utf-16-le", "surrogatepass")
        lens = np.fromiter(map(len, kws), dtype=np.int64, count=n)
        if int(lens.sum()) * 2 == len(joined):
            offsets = np.zeros(n + 1, np.int64)
            np.cumsum(lens, out=offsets[1:])
            chars = np.frombuffer(joined, dtype=np.uint16) if joined else np.zeros(1, np.uint16)
            return np.ascontiguousarray(chars, np.uint16), offsets, np.zeros(n, np.uint8), n
    units: List[np.ndarray] = []
    nulls: List[int] = []
    for k in kws:
        if k is None:
            units.append(np.zeros(0, np.uint16))
            nulls.append(1)
        else:
            units.append(_utf16(k))
            nulls.append(0)
    offsets = np.zeros(n + 1, np.int64)
    if n:
        np.cumsum([u.size for u in units], out=offsets[1:])
    chars = np.concatenate(units) if n else np.zeros(0, np.uint16)
    if chars.size == 0:
        chars = np.zeros(1, np.uint16)
    is_null = np.array(nulls, dtype=np.uint8) if n else np.zeros(1, np.uint8)
    return np.ascontiguousarray(chars, np.uint16), offsets, is_null, n


class WordCharacters:
    """WordCharacters.java:4-63."""

    @staticmethod
    def generateWordCharsFlags(wordCharacters: Optional[Sequence[str]] = None,
                               toggleFlags: Optional[Sequence[bool]] = None) -> np.ndarray:
        out = np.zeros(65536, np.uint8)
        if wordCharacters is None:
            mode, ch, tg, n = 0, np.zeros(1, np.uint16), np.zeros(1, np.uint8), 0
        else:
            n = len(wordCharacters)
            ch = np.array([ord(c) for c in wordCharacters], dtype=np.uint16) if n else np.zeros(1, np.uint16)
            if toggleFlags is None:
                mode, tg = 1, np.zeros(1, np.uint8)
            else:
                if len(toggleFlags) < n:
                    raise IndexError("toggleFlags shorter than wordCharacters")  # ArrayIndexOutOfBounds in Java
                mode = 2
                tg = np.array([1 if t else 0 for t in toggleFlags[:n]], dtype=np.uint8) if n else np.zeros(1, np.uint8)
        check(_lib.lib().acgpu_word_chars(mode, ch.ctypes.data, tg.ctypes.data, n, out.ctypes.data))
        return out.astype(bool)

    @staticmethod
    def trim(keyword: str, wordChars: np.ndarray) -> str:
        u = _utf16(keyword)
        idx = np.nonzero(wordChars[u])[0]
        if idx.size == 0:
            return keyword
        return u[idx[0]:idx[-1] + 1].tobytes().decode("utf-16-le", "surrogatepass")


class _Records:
    """Ordered match records copied out of an acgpu_result."""

    __slots__ = ("start", "end", "value")

    def __init__(self, res, is_map: bool):
        n = int(res.n)
        if n and getattr(res, "kind", _lib.MATCHES_RECORDS) == _lib.MATCHES_MASKS:
            # compact wire format (dense AhoCorasickSet streams): expand the per-char hit masks on the host
            pos = np.empty((n, 2), np.int32)
            k = _lib.lib().acgpu_masks_to_records(res.masks, res.n_chars, 0, pos.ctypes.data, n)
            if k != n:
                raise _lib.AcgpuError(_lib.ECUDA, "hit masks hold %d matches, the call reported %d" % (k, n))
            self.start, self.end, self.value = pos[:, 0], pos[:, 1], None
        elif n and not res.pos:
            # values-only stream results (acgpu_stream_set_values_only): the positions stayed on the device
            self.start = self.end = None
            self.value = np.ctypeslib.as_array(res.val, shape=(n,)).copy()
        elif n:
            pos = np.ctypeslib.as_array(res.pos, shape=(n, 2)).copy()
            self.start, self.end = pos[:, 0], pos[:, 1]
            self.value = np.ctypeslib.as_array(res.val, shape=(n,)).copy() if (is_map and res.val) else None
        else:
            self.start = self.end = np.zeros(0, np.int32)
            self.value = np.zeros(0, np.uint32) if is_map else None

    def __len__(self):
        return self.start.size if self.start is not None else self.value.size


class _Matcher:
    _family = -1
    _is_map = False

    def _create(self, keywords: Iterable, values: Optional[Iterable], caseSensitive: bool,
                word_flags: Optional[np.ndarray], device: int = 0):
        if values is not None:
            # Maps zip keywords with values and stop at the shorter (AhoCorasickMap.java:32)
            pairs = list(zip(keywords, values))
            kws = [p[0] for p in pairs]
            self._values = [p[1] for p in pairs]
            n_values = len(pairs)
        else:
            kws = keywords
            self._values = None
            n_values = -1
        chars, offsets, is_null, n = _pack_keywords(kws)
        wc = None
        if word_flags is not None:
            wc = np.ascontiguousarray(np.asarray(word_flags).astype(np.uint8))
            if wc.size != 65536:
                raise ValueError("word character table must have 65536 entries")
        self._caseSensitive = bool(caseSensitive)
        h = C.c_uint64(0)
        self._h = 0
        check(_lib.lib().acgpu_create_from_keywords(
            self._family, chars.ctypes.data, offsets.ctypes.data, is_null.ctypes.data, n, n_values,
            1 if caseSensitive else 0, wc.ctypes.data if wc is not None else None, device, C.byref(h)))
        self._h = h.value
        self._device = device

    @classmethod
    def from_trie(cls, trie, values: Optional[Sequence] = None, device: int = 0):
        """A matcher from an already flattened goto trie (trie_desc.FlatTrie -> acgpu_create): what a Java-side builder hands
        over instead of the keyword strings.  `values`: the Map's value objects, indexed by the trie's value indices."""
        if trie.family != cls._family or trie.is_map != cls._is_map:
            raise ValueError("the trie was flattened for another matcher class")
        self = cls.__new__(cls)
        self._values = list(values) if values is not None else None
        self._caseSensitive = trie.case_sensitive
        if trie.family in (_lib.WHOLEWORD, _lib.WHOLEWORDLONGEST):
            self._wordChars = trie.word_flags if trie.word_flags is not None else WordCharacters.generateWordCharsFlags()
        self._h = 0
        h = C.c_uint64(0)
        d = trie.desc(device)
        check(_lib.lib().acgpu_create(C.byref(d), C.byref(h)))
        self._h = h.value
        self._device = device
        return self

    def close(self):
        h = getattr(self, "_h", 0)
        if h:
            self._h = 0
            _lib.lib().acgpu_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection (not part of the reference API)
    def info(self) -> dict:
        n_nodes, table_bytes = C.c_int64(), C.c_int64()
        n_classes, max_len, cbs = C.c_int32(), C.c_int32(), C.c_int32()
        check(_lib.lib().acgpu_info(self._h, C.byref(n_nodes), C.byref(n_classes), C.byref(max_len), C.byref(cbs),
                                    C.byref(table_bytes)))
        return dict(n_nodes=n_nodes.value, n_classes=n_classes.value, max_len=max_len.value,
                    char_buffer_size=cbs.value, table_bytes=table_bytes.value)

    @property
    def handle(self) -> int:
        return self._h

    def char_classes(self) -> Tuple[np.ndarray, bool]:
        """(classes[65536], has_other): class of every UTF-16 code unit under this matcher's case folding; with has_other,
        class 0 = "occurs in no keyword" (not part of the reference API; used by sharding.plan_sync_shards)."""
        out = np.zeros(65536, np.uint16)
        ho = C.c_int32(0)
        check(_lib.lib().acgpu_char_classes(self._h, out.ctypes.data, C.byref(ho)))
        return out, bool(ho.value)

    def match_records(self, haystack, compact: bool = True) -> _Records:
        """The ordered (start, end[, valueIdx]) stream of one match(String) call, without replay.  compact (default): through
        acgpu_match_utf16_compact - dense AhoCorasickSet streams cross PCIe as 2-byte hit masks and are expanded here;
        compact=False: through acgpu_match_utf16 (records on the wire)."""
        hay = _utf16(haystack)
        if hay.size > 0x7FFFFFFF:
            raise ValueError("haystack longer than a Java String (2^31 - 1 chars)")
        if compact:
            res = _lib.Matches()
            check(_lib.lib().acgpu_match_utf16_compact(self._h, hay.ctypes.data if hay.size else None, hay.size, C.byref(res)))
            try:
                return _Records(res, self._is_map)
            finally:
                _lib.lib().acgpu_free_matches(C.byref(res))
        res = _lib.Result()
        check(_lib.lib().acgpu_match_utf16(self._h, hay.ctypes.data if hay.size else None, hay.size, C.byref(res)))
        try:
            return _Records(res, self._is_map)
        finally:
            _lib.lib().acgpu_free_result(C.byref(res))

    # -- replay: reproduce the reference's listener call sequence, including early-stop quirks
    def _replay_string(self, haystack, rec: _Records, emit: Callable[[int], bool], n_chars: int):
        n = len(rec)
        if self._family == _lib.SHORTEST:
            # ShortestMatchSet.java:196-226: a `false` breaks out of the loop and falls into the post-loop
            # emit, which delivers the same match once more (quirk Q1); the match ending at the end of the
            # haystack is only emitted post-loop and its return value is ignored (Q2).
            end = rec.end
            for i in range(n):
                if int(end[i]) == n_chars:
                    emit(i)
                    return
                if not emit(i):
                    emit(i)
                    return
            return
        for i in range(n):
            if not emit(i):
                return


class StringSet(_Matcher):
    """StringSet.java:3-5."""

    def match(self, haystack, listener) -> None:
        """match(String haystack, SetMatchListener listener) — StringSet.java:4."""
        if haystack is None or listener is None:
            raise TypeError("NullPointerException: haystack and listener must not be None")
        cb = _callable_of(listener)
        hay = _utf16(haystack)
        rec = self.match_records(hay)
        start, end = rec.start, rec.end
        self._replay_string(haystack, rec, lambda i: bool(cb(haystack, int(start[i]), int(end[i]))), hay.size)


class StringMap(_Matcher):
    """StringMap.java:5-9."""

    _is_map = True

    def _value(self, idx: int):
        return self._values[idx]

    def match(self, haystack, listener) -> None:
        """match(String, MapMatchListener) (StringMap.java:8) or match(Readable, ReadableMatchListener)
        (StringMap.java:6), chosen by the haystack type like the Java overloads."""
        if haystack is None or listener is None:
            raise TypeError("NullPointerException: haystack and listener must not be None")
        cb = _callable_of(listener)
        if isinstance(haystack, (str, np.ndarray)):
            hay = _utf16(haystack)
            rec = self.match_records(hay)
            start, end, val = rec.start, rec.end, rec.value
            values = self._values
            self._replay_string(
                haystack, rec, lambda i: bool(cb(haystack, int(start[i]), int(end[i]), values[int(val[i])])), hay.size)
            return
        if not hasattr(haystack, "read"):
            raise TypeError("haystack must be str, a uint16 array or a Readable (object with read(n))")
        from .streaming import match_readable
        match_readable(self, haystack, cb)


def _split_thresholder(args: tuple) -> tuple:
    """The reference overloads accept a trailing Thresholder; it never changes results, so drop it."""
    if args and (isinstance(args[-1], Thresholder) or args[-1] is None):
        return args[:-1]
    return args


# ----------------------------------------------------------------------------- the eight public classes

class AhoCorasickSet(StringSet):
    """AhoCorasickSet.java:11-252 — all (overlapping) occurrences, ordered by end then longest first."""
    _family = _lib.AHOCORASICK

    def __init__(self, keywords: Iterable[Optional[str]], caseSensitive: bool, thresholdStrategy: Optional[Thresholder] = None,
                 device: int = 0):
        self._create(keywords, None, caseSensitive, None, device)


class AhoCorasickMap(StringMap):
    """AhoCorasickMap.java:14-336."""
    _family = _lib.AHOCORASICK

    def __init__(self, keywords: Iterable[Optional[str]], values: Iterable, caseSensitive: bool,
                 thresholdStrategy: Optional[Thresholder] = None, device: int = 0):
        self._create(keywords, values, caseSensitive, None, device)


class LongestMatchSet(StringSet):
    """LongestMatchSet.java:10-265 — leftmost-longest, non-overlapping."""
    _family = _lib.LONGEST

    def __init__(self, keywords, caseSensitive: bool, thresholdStrategy: Optional[Thresholder] = None, device: int = 0):
        self._create(keywords, None, caseSensitive, None, device)


class LongestMatchMap(StringMap):
    """LongestMatchMap.java:14-361."""
    _family = _lib.LONGEST

    def __init__(self, keywords, values, caseSensitive: bool, thresholdStrategy: Optional[Thresholder] = None,
                 device: int = 0):
        self._create(keywords, values, caseSensitive, None, device)


class ShortestMatchSet(StringSet):
    """ShortestMatchSet.java:10-260 — earliest-ending, non-overlapping."""
    _family = _lib.SHORTEST

    def __init__(self, keywords, caseSensitive: bool, thresholdStrategy: Optional[Thresholder] = None, device: int = 0):
        self._create(keywords, None, caseSensitive, None, device)


class ShortestMatchMap(StringMap):
    """ShortestMatchMap.java:14-373."""
    _family = _lib.SHORTEST

    def __init__(self, keywords, values, caseSensitive: bool, thresholdStrategy: Optional[Thresholder] = None,
                 device: int = 0):
        self._create(keywords, values, caseSensitive, None, device)


def _word_flags(rest: tuple) -> np.ndarray:
    """(…, char[] wordCharacters[, boolean[] toggleFlags]) tail of the WholeWord constructors."""
    rest = _split_thresholder(rest)
    if len(rest) == 0:
        return WordCharacters.generateWordCharsFlags()
    if len(rest) == 1:
        return WordCharacters.generateWordCharsFlags(rest[0])
    if len(rest) == 2:
        return WordCharacters.generateWordCharsFlags(rest[0], rest[1])
    raise TypeError("too many constructor arguments")


class WholeWordMatchSet(StringSet):
    """WholeWordMatchSet.java:8-205.  Overloads: (keywords, caseSensitive[, wordCharacters[, toggleFlags]][, Thresholder])."""
    _family = _lib.WHOLEWORD

    def __init__(self, keywords, caseSensitive: bool, *rest, device: int = 0):
        self._wordChars = _word_flags(rest)
        self._create(keywords, None, caseSensitive, self._wordChars, device)

    def getWordChars(self) -> np.ndarray:
        """WholeWordMatchSet.java:134."""
        return self._wordChars


class WholeWordMatchMap(StringMap):
    """WholeWordMatchMap.java:13-339."""
    _family = _lib.WHOLEWORD

    def __init__(self, keywords, values, caseSensitive: bool, *rest, device: int = 0):
        self._wordChars = _word_flags(rest)
        self._create(keywords, values, caseSensitive, self._wordChars, device)

    def getWordChars(self) -> np.ndarray:
        """WholeWordMatchMap.java:242."""
        return self._wordChars


class WholeWordLongestMatchSet(StringSet):
    """WholeWordLongestMatchSet.java:9-260.  Same constructor overloads as WholeWordMatchSet; keywords are trimmed
    but may hold non-word chars inside ("as if"), so no IllegalArgumentException."""
    _family = _lib.WHOLEWORDLONGEST

    def __init__(self, keywords, caseSensitive: bool, *rest, device: int = 0):
        self._wordChars = _word_flags(rest)
        self._create(keywords, None, caseSensitive, self._wordChars, device)

    def getWordChars(self) -> np.ndarray:
        """WholeWordLongestMatchSet.java:180."""
        return self._wordChars


class WholeWordLongestMatchMap(StringMap):
    """WholeWordLongestMatchMap.java:13-420."""
    _family = _lib.WHOLEWORDLONGEST

    def __init__(self, keywords, values, caseSensitive: bool, *rest, device: int = 0):
        self._wordChars = _word_flags(rest)
        self._create(keywords, values, caseSensitive, self._wordChars, device)

    def getWordChars(self) -> np.ndarray:
        return self._wordChars
