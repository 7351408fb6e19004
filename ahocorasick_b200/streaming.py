"""match(Readable, ReadableMatchListener) — host side (StringMap.java:6).

The reference pulls ``charBufferSize`` chars at a time (AhoCorasickMap.java:53,213-245) and reports
values only.  Here the fills are batched into large blocks and pushed through libacgpu.so's
acgpu_stream_* entry points (pinned double buffers + cudaMemcpyAsync, automaton context carried on the
device); the records of every block are replayed to the listener in order.  Observable difference
(documented in DESIGN.md): the Readable is consumed up to one block ahead of the listener calls.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, List

import numpy as np

from . import _lib
from ._lib import check

# Device block policy of match(Readable).  The reference consumes the Readable charBufferSize chars at a time and stops
# reading on an early stop; a device block is the unit this path reads ahead of the listener.  Blocks therefore start
# small (a caller that stops after the first match has over-read at most 64 Ki chars) and double up to MAX_BLOCK_CHARS,
# where the fixed cost of a feed (synchronisation + two small copies, ~0.4 ms) is amortised (tools/bench_stream_sweep.py).
FIRST_BLOCK_CHARS = 1 << 16
MAX_BLOCK_CHARS = 1 << 24
BLOCK_CHARS = 1 << 22  # a fixed block size callers / tests may pass explicitly


def read_fills(readable, fill_size: int) -> Iterator[np.ndarray]:
    """The successive CharBuffer fills the reference loop would see: read(fill_size) until EOF."""
    while True:
        chunk = readable.read(fill_size)
        if chunk is None or len(chunk) == 0:
            return
        if isinstance(chunk, str):
            yield np.frombuffer(chunk.encode("utf-16-le", "surrogatepass"), dtype=np.uint16)
        else:
            yield np.ascontiguousarray(chunk, np.uint16)


class DeviceStream:
    """Thin wrapper over acgpu_stream_begin / feed / end."""

    def __init__(self, matcher):
        self._m = matcher
        h = C.c_uint64(0)
        check(_lib.lib().acgpu_stream_begin(matcher.handle, C.byref(h)))
        self._h = h.value
        # ReadableMatchListener sees values only: the positions need not cross PCIe - except for ShortestMatchMap, whose
        # replay compares the match ends with the CharBuffer fill boundaries (quirk Q4)
        if matcher._is_map and matcher._family != _lib.SHORTEST:
            check(_lib.lib().acgpu_stream_set_values_only(self._h, 1))

    def feed(self, chars: np.ndarray):
        from .matchers import _Records
        res = _lib.Result()
        check(_lib.lib().acgpu_stream_feed(self._h, chars.ctypes.data if chars.size else None, chars.size, C.byref(res)))
        try:
            return _Records(res, self._m._is_map)
        finally:
            _lib.lib().acgpu_free_result(C.byref(res))

    def end(self):
        from .matchers import _Records
        res = _lib.Result()
        h, self._h = self._h, 0
        check(_lib.lib().acgpu_stream_end(h, C.byref(res)))
        try:
            return _Records(res, self._m._is_map)
        finally:
            _lib.lib().acgpu_free_result(C.byref(res))

    def abort(self):
        if self._h:
            h, self._h = self._h, 0
            _lib.lib().acgpu_stream_end(h, None)

    def __del__(self):
        try:
            self.abort()
        except Exception:
            pass


def match_readable(matcher, readable, cb, block_chars: int = 0) -> None:
    """block_chars = 0: adaptive blocks (FIRST_BLOCK_CHARS doubling to MAX_BLOCK_CHARS); > 0: fixed block size."""
    adaptive = block_chars <= 0
    if adaptive:
        block_chars = FIRST_BLOCK_CHARS
    cbs = matcher.info()["char_buffer_size"]
    stream = DeviceStream(matcher)
    shortest = matcher._family == _lib.SHORTEST
    values = matcher._values
    boundaries = set()   # absolute stream offsets at which a CharBuffer fill ended (quirk Q4)
    n_read = 0
    try:
        fills = read_fills(readable, cbs)
        eof = False
        while not eof:
            parts: List[np.ndarray] = []
            got = 0
            while got < block_chars:
                f = next(fills, None)
                if f is None:
                    eof = True
                    break
                parts.append(f)
                got += f.size
                n_read += f.size
                if shortest:
                    boundaries.add(n_read)
            rec = stream.feed(np.concatenate(parts)) if parts else None
            if adaptive:
                block_chars = min(2 * block_chars, MAX_BLOCK_CHARS)
            if rec is not None and not _replay(rec, cb, values, shortest, boundaries, n_read, False):
                stream.abort()
                return
        rec = stream.end()
        _replay(rec, cb, values, shortest, boundaries, n_read, True)
    finally:
        stream.abort()


def _replay(rec, cb, values, shortest, boundaries, n_read, final) -> bool:
    """Values-only replay, stop at the first False.  ShortestMatchMap.match(Readable)
    (ShortestMatchMap.java:199-291) emits the pending match at the end of every buffer fill and does not
    clear it, so a match that ends exactly on a fill boundary and is followed by more input is delivered
    twice (quirk Q4)."""
    val = rec.value
    if shortest:
        end = rec.end
        for i in range(len(rec)):
            v = values[int(val[i])]
            if not cb(v):
                return False
            e = int(end[i])
            if e in boundaries and e < n_read:
                if not cb(v):
                    return False
        return True
    for i in range(len(rec)):
        if not cb(values[int(val[i])]):
            return False
    return True
