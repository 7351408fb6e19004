"""ctypes front end of the CPU ORACLE (oracle/ac_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never from the product package
``ahocorasick_b200`` (tests/test_host_cpu.py::test_product_does_not_import_oracle enforces that).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

AHOCORASICK, LONGEST, SHORTEST, WHOLEWORD, WHOLEWORDLONGEST = 0, 1, 2, 3, 4
FAMILIES = {"ahocorasick": AHOCORASICK, "longest": LONGEST, "shortest": SHORTEST, "wholeword": WHOLEWORD,
            "wholewordlongest": WHOLEWORDLONGEST}


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc if missing or stale (gcc is in the image)."""
    srcs = [os.path.join(_HERE, f) for f in ("ac_oracle.c", "ac_oracle.h", "java_char_tables.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ora_create.restype = C.c_void_p
        L.ora_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                 C.c_int, C.c_void_p, C.c_char_p, C.c_int]
        L.ora_destroy.argtypes = [C.c_void_p]
        L.ora_char_buffer_size.restype = C.c_int32
        L.ora_char_buffer_size.argtypes = [C.c_void_p]
        L.ora_node_count.restype = C.c_int64
        L.ora_node_count.argtypes = [C.c_void_p]
        L.ora_match_collect.restype = C.c_int64
        L.ora_match_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_int64]
        L.ora_match_readable.restype = C.c_int64
        L.ora_match_readable.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p]
        L.ora_match_count.restype = C.c_int64
        L.ora_match_count.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.ora_queue_new.restype = C.c_void_p
        L.ora_queue_free.argtypes = [C.c_void_p]
        L.ora_queue_push.restype = C.c_int
        L.ora_queue_push.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.ora_queue_match_and_clear.restype = C.c_int64
        L.ora_queue_match_and_clear.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
        L.ora_to_lower.restype = C.c_uint16
        L.ora_to_lower.argtypes = [C.c_uint16]
        L.ora_is_letter_or_digit.restype = C.c_int
        L.ora_is_letter_or_digit.argtypes = [C.c_uint16]
        L.ora_word_chars.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        _lib = L
    return _lib


MATCH_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("value", "<i4")])


def utf16(s) -> np.ndarray:
    """str (or an already-encoded uint16 array) -> UTF-16 code units, like a Java String."""
    if isinstance(s, np.ndarray):
        return np.ascontiguousarray(s, dtype=np.uint16)
    return np.frombuffer(s.encode("utf-16-le", "surrogatepass"), dtype=np.uint16).copy()


def pack_keywords(keywords: Sequence) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Keyword list (str / uint16 arrays / None) -> (chars, offsets[n+1], is_null[n])."""
    units = [np.zeros(0, np.uint16) if k is None else utf16(k) for k in keywords]
    offsets = np.zeros(len(units) + 1, np.int64)
    if units:
        np.cumsum([len(u) for u in units], out=offsets[1:])
    chars = np.concatenate(units) if units else np.zeros(0, np.uint16)
    if chars.size == 0:
        chars = np.zeros(1, np.uint16)
    is_null = np.array([k is None for k in keywords], dtype=np.uint8)
    if is_null.size == 0:
        is_null = np.zeros(1, np.uint8)
    return np.ascontiguousarray(chars, np.uint16), offsets, is_null


def word_chars(mode: int = 0, chars: Sequence[str] = (), toggles: Sequence[bool] = ()) -> np.ndarray:
    """WordCharacters.generateWordCharsFlags: mode 0 default, 1 custom-only, 2 default+toggles."""
    out = np.zeros(65536, np.uint8)
    ch = np.array([ord(c) for c in chars], dtype=np.uint16) if len(chars) else np.zeros(1, np.uint16)
    tg = np.array([1 if t else 0 for t in toggles], dtype=np.uint8) if len(toggles) else np.zeros(1, np.uint8)
    lib().ora_word_chars(mode, ch.ctypes.data, tg.ctypes.data, len(chars), out.ctypes.data)
    return out


class OracleError(ValueError):
    """Stands in for the reference's IllegalArgumentException."""


class Matcher:
    """One reference matcher (family x Set/Map), restated on the CPU."""

    def __init__(self, family, keywords: Sequence, n_values: int = -1, case_sensitive: bool = True,
                 word_chars_table: Optional[np.ndarray] = None):
        fam = FAMILIES[family] if isinstance(family, str) else int(family)
        chars, offsets, is_null = pack_keywords(keywords)
        err = C.create_string_buffer(4096)
        wc = None
        if word_chars_table is not None:
            wc = np.ascontiguousarray(word_chars_table, np.uint8)
            assert wc.size == 65536
        self._h = lib().ora_create(fam, chars.ctypes.data, offsets.ctypes.data, is_null.ctypes.data,
                                   len(keywords), n_values, 1 if case_sensitive else 0,
                                   wc.ctypes.data if wc is not None else None, err, len(err))
        if not self._h:
            raise OracleError(err.value.decode("utf-8", "replace"))
        self.family = fam

    def close(self):
        if getattr(self, "_h", None):
            lib().ora_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def char_buffer_size(self) -> int:
        return lib().ora_char_buffer_size(self._h)

    @property
    def node_count(self) -> int:
        return lib().ora_node_count(self._h)

    def match(self, haystack, readable: bool = False, stop_after: int = 0, cap: int = 1 << 16) -> np.ndarray:
        """All listener calls, in order, as a (start, end, value) record array.  `cap`: first guess of the record
        count (a too small guess costs a second pass)."""
        hay = utf16(haystack)
        n = hay.size
        if n == 0:
            hay = np.zeros(1, np.uint16)
        cap = max(1, int(cap))
        while True:
            out = np.zeros(cap, MATCH_DTYPE)
            total = lib().ora_match_collect(self._h, hay.ctypes.data, n, 1 if readable else 0,
                                            stop_after, out.ctypes.data, cap)
            if total <= cap:
                return out[:total]
            cap = int(total)

    def match_readable_schedule(self, haystack, schedule: Sequence[int]) -> np.ndarray:
        """Readable overload with an explicit read-size schedule (quirk Q4)."""
        hay = utf16(haystack)
        n = hay.size
        if n == 0:
            hay = np.zeros(1, np.uint16)
        sched = np.array(list(schedule), dtype=np.int32) if len(schedule) else np.zeros(1, np.int32)
        records: List[Tuple[int, int, int]] = []
        CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int32)

        def cb(_ctx, s, e, v):
            records.append((s, e, v))
            return 1

        cb_obj = CB(cb)  # keep alive for the duration of the call
        lib().ora_match_readable(self._h, hay.ctypes.data, n, sched.ctypes.data, len(schedule),
                                 C.cast(cb_obj, C.c_void_p), None)
        return np.array(records, dtype=MATCH_DTYPE) if records else np.zeros(0, MATCH_DTYPE)

    def count(self, haystack) -> int:
        hay = utf16(haystack)
        return lib().ora_match_count(self._h, hay.ctypes.data if hay.size else None, hay.size)


class MatchQueue:
    """SetMatchQueue, for the reference's MatchQueueTest known answers."""

    def __init__(self):
        self._q = lib().ora_queue_new()

    def push(self, length: int, idx: int) -> bool:
        return bool(lib().ora_queue_push(self._q, length, idx))

    def match_and_clear(self, purge_to: int) -> List[Tuple[int, int]]:
        out = np.zeros(1024, MATCH_DTYPE)
        n = lib().ora_queue_match_and_clear(self._q, purge_to, out.ctypes.data, 1024)
        return [(int(r["end"]), int(r["end"] - r["start"])) for r in out[:n]]

    def __del__(self):
        if getattr(self, "_q", None):
            lib().ora_queue_free(self._q)
            self._q = None
