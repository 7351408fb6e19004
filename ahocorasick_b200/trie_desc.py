"""acgpu_create(const acgpu_automaton_desc*) - the constructor entry for a caller that flattened the dictionary itself
(include/acgpu.h; BASELINE.json north_star (1): "a Java-side trie builder ... uploaded once").

flatten_trie() restates what that Java-side builder does with the reference's constructor arguments - skip nulls, zip
keywords with values and stop at the shorter, last duplicate wins (Shortest: the first), one trie state per distinct prefix, breadth-first
failure links (AhoCorasickSet.java:20-191, AhoCorasickMap.java:24-206) - and returns the arrays of the descriptor.  It is
host-side glue for tests and examples; the device tables are derived from the descriptor inside libacgpu.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import numpy as np

from . import _lib


class AutomatonDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("family", C.c_int32), ("is_map", C.c_int32), ("case_sensitive", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32), ("n_states", C.c_int64), ("parent", C.c_void_p),
                ("edge_char", C.c_void_p), ("terminal", C.c_void_p), ("value", C.c_void_p), ("n_values", C.c_int64),
                ("word_chars", C.c_void_p), ("fail", C.c_void_p)]


class FlatTrie:
    """The goto trie of a dictionary as parallel arrays (state 0 = root, parent[s] < s)."""

    def __init__(self, family: int, is_map: bool, case_sensitive: bool, parent, edge_char, terminal, value, n_values,
                 word_flags=None, fail=None):
        self.family, self.is_map, self.case_sensitive = int(family), bool(is_map), bool(case_sensitive)
        self.parent = np.ascontiguousarray(parent, np.int32)
        self.edge_char = np.ascontiguousarray(edge_char, np.uint16)
        self.terminal = np.ascontiguousarray(terminal, np.uint8)
        self.value = np.ascontiguousarray(value, np.uint32) if is_map else None
        self.n_values = int(n_values) if is_map else 0
        self.word_flags = None if word_flags is None else np.ascontiguousarray(np.asarray(word_flags).astype(np.uint8))
        self.fail = None if fail is None else np.ascontiguousarray(fail, np.int32)

    def desc(self, device: int = 0) -> AutomatonDesc:
        d = AutomatonDesc()
        d.struct_size = C.sizeof(AutomatonDesc)
        d.family, d.is_map, d.case_sensitive, d.device = self.family, int(self.is_map), int(self.case_sensitive), device
        d.n_states = self.parent.size
        d.parent, d.edge_char, d.terminal = self.parent.ctypes.data, self.edge_char.ctypes.data, self.terminal.ctypes.data
        d.value = self.value.ctypes.data if self.value is not None else None
        d.n_values = self.n_values
        d.word_chars = self.word_flags.ctypes.data if self.word_flags is not None else None
        d.fail = self.fail.ctypes.data if self.fail is not None else None
        return d

    def fingerprint(self) -> int:
        """Host only: the fingerprint of the tables libacgpu derives from this trie."""
        fp = C.c_uint64(0)
        d = self.desc()
        _lib.check(_lib.lib().acgpu_desc_fingerprint(C.byref(d), C.byref(fp)))
        return fp.value


def flatten_trie(family: int, keywords: Iterable, values: Optional[Iterable] = None, case_sensitive: bool = True,
                 word_flags=None, with_fail: bool = True) -> FlatTrie:
    """The Java-side builder: keywords (str or None) [zipped with values] -> goto trie arrays (+ failure links)."""
    is_map = values is not None
    if is_map:
        entries = [k for k, _ in zip(keywords, values)]
    else:
        entries = list(keywords)
    parent, edge, term, val = [-1], [0], [0], [0]
    kids = [dict()]
    for idx, kw in enumerate(entries):
        if kw is None:
            continue
        units = np.frombuffer(kw.encode("utf-16-le", "surrogatepass"), dtype=np.uint16)
        if units.size == 0:
            continue
        s = 0
        for u in units.tolist():
            if not case_sensitive and u < 0x80:
                u = ord(chr(u).lower())  # the library folds every unit again with Character.toLowerCase (idempotent)
            nxt = kids[s].get(u)
            if nxt is None:
                nxt = len(parent)
                kids[s][u] = nxt
                parent.append(s); edge.append(u); term.append(0); val.append(0); kids.append(dict())
            s = nxt
        if term[s] and family == _lib.SHORTEST:
            continue  # ShortestMatchMap.java:44-54: the first duplicate keeps its value
        term[s] = 1
        val[s] = idx  # every other family: the last duplicate wins
    fail = None
    if with_fail:
        fail = [0] * len(parent)
        order, head = [0], 0
        while head < len(order):  # breadth first
            s = order[head]; head += 1
            for u, c in kids[s].items():
                order.append(c)
                if s == 0:
                    continue
                t = fail[s]
                while True:
                    n2 = kids[t].get(u)
                    if n2 is not None:
                        fail[c] = n2
                        break
                    if t == 0:
                        break
                    t = fail[t]
    return FlatTrie(family, is_map, case_sensitive, parent, edge, term, val, len(entries), word_flags, fail)
