"""Multi-GPU sharding of the matching path — host-side plan + the one exchange the path has.

The AhoCorasick family needs no automaton state across positions (every end position is matched from its own
left context), so a corpus shards with NO data-path collective:

* a corpus of independent haystacks (a Java String holds < 2^31 chars, so a 64 GB corpus is many match() calls,
  AhoCorasickSet.java:193) is dealt round-robin: ``deal_haystacks``;
* one large haystack is cut by END-position range: rank r reports the matches whose end lies in its range and reads
  ``max_len - 1`` chars of left context: ``plan_range_shards``.  Concatenating the ranks' record streams in rank
  order gives exactly the single-GPU stream (end ascending, longest first).

The WholeWord family shards the same way by word-START range (``plan_word_shards``).  Longest / Shortest carry a
selection chain across positions; one haystack is cut at SYNCHRONISATION points - chars that occur in no keyword reset
every reference automaton - into independent pieces (``plan_sync_shards``; exact, but needs such chars near the even
split, else the haystack stays whole) or - for ANY text, also one without such chars - into runs of whole tiles whose
entry -> exit maps are exchanged and composed (``plan_chain_shards`` / ``compose_chain_maps``, acgpu_chain_shard_*).
WholeWordLongest shards by haystack or at synchronisation points.

The only exchange is an all-gather of per-rank match counts (8 bytes per rank) so that every rank knows its
global record offset: ``exchange_counts`` (NCCL on device tensors, gloo on CPU tensors in the tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple


@dataclass(frozen=True)
class RangeShard:
    rank: int
    emit_from: int   # first end-anchor position (index of the match's last char) this rank reports
    emit_to: int     # one past the last
    read_from: int   # first haystack char this rank needs on its device (left context)

    @property
    def read_to(self) -> int:
        return self.emit_to


def plan_range_shards(n_chars: int, world: int, max_len: int, align: int = 8) -> List[RangeShard]:
    """Cut [0, n_chars) into `world` contiguous end-position ranges (boundaries aligned to `align` chars so every
    shard keeps the kernel's 16-byte aligned loads)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    ctx = max(0, max_len - 1)
    bounds = [0]
    for r in range(1, world):
        b = (n_chars * r // world) // align * align
        bounds.append(max(bounds[-1], min(b, n_chars)))
    bounds.append(n_chars)
    return [RangeShard(r, bounds[r], bounds[r + 1], max(0, bounds[r] - ctx)) for r in range(world)]


@dataclass(frozen=True)
class WordShard:
    rank: int
    emit_from: int   # first word-START position this rank reports
    emit_to: int     # one past the last
    read_from: int   # one char of look-behind (is emit_from a word start?)
    read_to: int     # emit_to + max_len + 1, clipped: a word starting before emit_to is decided by then


def plan_word_shards(n_chars: int, world: int, max_len: int) -> List[WordShard]:
    """WholeWord family: a word is a keyword or not by itself, so one large haystack is cut by word-START
    range with no state across shards; boundaries may fall anywhere, also inside a word (the word belongs to the
    shard holding its first char).  Rank-ordered concatenation of the shards' streams is the single-GPU stream."""
    if world < 1:
        raise ValueError("world must be >= 1")
    bounds = [n_chars * r // world for r in range(world)] + [n_chars]
    return [WordShard(r, bounds[r], bounds[r + 1], max(0, bounds[r] - 1), min(n_chars, bounds[r + 1] + max_len + 1))
            for r in range(world)]


def match_word_shard(matcher, d_haystack_ptr: int, shard: WordShard, d_pos_ptr: int, d_val_ptr, cap: int, stream_ptr=None) -> int:
    """Scan one word-start range (WholeWordMatchSet/Map) of a haystack resident on this rank's GPU; the pointer addresses
    char 0 of the WHOLE haystack, only [shard.read_from, shard.read_to) is touched.  Returns the number of matches."""
    import ctypes as C
    from . import _lib
    tot = C.c_int64(0)
    _lib.check(_lib.lib().acgpu_match_device(matcher.handle, d_haystack_ptr, shard.read_to, shard.emit_from, shard.emit_to,
                                             d_pos_ptr, d_val_ptr, cap, C.byref(tot), stream_ptr))
    return tot.value


@dataclass(frozen=True)
class SyncShard:
    rank: int
    lo: int   # the shard is the independent haystack [lo, hi); positions it reports are relative to lo
    hi: int


def plan_sync_shards(haystack, world: int, classes, has_other: bool, window: int = 1 << 20, word_chars=None):
    """Longest / Shortest (and AhoCorasick): cut ONE haystack into `world` independent pieces at SYNCHRONISATION points.

    A char that occurs in no keyword (class 0 of ``matcher.char_classes()``) sends every reference automaton back to its
    root (all transitions fail), flushes the Longest match queue (LongestMatchSet.java:227) and ends any pending Shortest
    match, so the scan of what follows does not depend on what came before: the rank-ordered concatenation of the pieces'
    streams (positions shifted by ``lo``) is exactly the single stream.  Boundaries are the first such position at or
    after the even split, searched over at most `window` chars; returns None when one boundary has no synchronisation
    point in its window (e.g. every char of the text occurs in some keyword) - the caller then keeps the haystack whole.
    `haystack` is a uint16 numpy array or an int16 / uint16 torch tensor (device tensors copy only the windows).

    WholeWordLongest: pass the matcher's word-character table (``getWordChars()``) as `word_chars`; a synchronisation
    point then also needs the char to be a NON-word char (after a keyword-free letter the rest of its word is no walk
    start in the reference, but it would be one at the start of a piece).  WholeWord itself shards by plan_word_shards."""
    import numpy as np
    if world < 1:
        raise ValueError("world must be >= 1")
    n = int(haystack.shape[0])
    if world == 1:
        return [SyncShard(0, 0, n)]
    if not has_other:
        return None
    is_other = np.asarray(classes) == 0
    if word_chars is not None:
        is_other &= ~np.asarray(word_chars).astype(bool)
    bounds = [0]
    for r in range(1, world):
        b = max(n * r // world, bounds[-1], 1)
        if b >= n:
            bounds.append(n)
            continue
        w = haystack[b - 1:min(n, b - 1 + window)]      # char p - 1 decides whether p is a synchronisation point
        if hasattr(w, "cpu"):
            w = w.cpu().numpy()
        hit = np.flatnonzero(is_other[np.asarray(w).view(np.uint16)])
        if hit.size == 0:
            return None
        bounds.append(b + int(hit[0]))
    bounds.append(n)
    return [SyncShard(r, bounds[r], bounds[r + 1]) for r in range(world)]


def match_sync_shard(matcher, d_haystack_ptr: int, shard: SyncShard, d_pos_ptr: int, d_val_ptr, cap: int, stream_ptr=None) -> int:
    """Scan one piece of plan_sync_shards as a haystack of its own (the pointer addresses char 0 of the WHOLE haystack).
    The records written are relative to shard.lo: add shard.lo to both columns when merging.  Returns the match count."""
    import ctypes as C
    from . import _lib
    tot = C.c_int64(0)
    n = shard.hi - shard.lo
    _lib.check(_lib.lib().acgpu_match_device(matcher.handle, d_haystack_ptr + 2 * shard.lo, n, 0, n, d_pos_ptr, d_val_ptr, cap,
                                             C.byref(tot), stream_ptr))
    return tot.value


@dataclass(frozen=True)
class ChainShard:
    rank: int
    lo: int        # first chain position (haystack position) this rank owns; the rank's window starts here
    hi: int        # one past the last; every inner boundary is a multiple of the tile size
    read_to: int   # the window is hay[lo, read_to): the domain plus the look-ahead

    @property
    def empty(self) -> bool:
        return self.hi <= self.lo


CHAIN_TILE, CHAIN_LOOKAHEAD, CHAIN_ENTRIES = 8192, 256, 16   # acgpu_chain_shard_layout()


def plan_chain_shards(n_chars: int, world: int, tile: int = CHAIN_TILE, lookahead: int = CHAIN_LOOKAHEAD) -> List[ChainShard]:
    """Longest / Shortest: cut ONE haystack into `world` runs of whole tiles.  Unlike plan_sync_shards this needs
    nothing from the text: the selection chain that crosses every boundary is composed from the shards' entry -> exit maps
    (compose_chain_maps).  Ranks beyond the last tile get empty shards."""
    if world < 1:
        raise ValueError("world must be >= 1")
    bounds = [0]
    for r in range(1, world):
        b = (n_chars * r // world + tile // 2) // tile * tile
        if b + lookahead > n_chars:       # an inner shard needs its look-ahead inside the haystack
            b = n_chars
        bounds.append(max(bounds[-1], min(b, n_chars)))
    bounds.append(n_chars)
    for r in range(1, world):             # a boundary that fell back to n_chars closes the haystack: nothing after it
        if bounds[r] == n_chars:
            bounds[r + 1:] = [n_chars] * (world - r)
            break
    return [ChainShard(r, bounds[r], bounds[r + 1], bounds[r + 1] if bounds[r + 1] == n_chars else bounds[r + 1] + lookahead)
            for r in range(world)]


def compose_chain_maps(maps: Sequence[Sequence[int]]) -> Tuple[List[int], List[int]]:
    """maps[r][e] = (exit offset | matches << 8) of shard r entered at offset e (acgpu_chain_shard_begin; an EMPTY shard
    contributes the identity map [0, 1, ..., 15]).  Returns (entry offset of every rank, index of every rank's first record):
    entry_0 = 0, entry_{r+1} = exit offset of maps[r][entry_r].  The map of the LAST shard is never consulted (nothing follows
    it; its window need not be tile-aligned, and then only its entry-0 row is meaningful) - the total is the sum of the
    counts the shards report when they finish."""
    entries, firsts, cur, acc = [], [], 0, 0
    for r, mp in enumerate(maps):
        entries.append(cur)
        firsts.append(acc)
        if r + 1 < len(maps):
            t = int(mp[cur])
            if t < 0 or (t & 0xFF) >= CHAIN_ENTRIES:
                raise ValueError("shard %d has no map row for entry offset %d (its window is not tile-aligned)" % (r, cur))
            cur, acc = t & 0xFF, acc + (t >> 8)
    return entries, firsts


def chain_shard_begin(matcher, d_window_ptr: int, n_window: int, n_domain: int, d_map_ptr: int, stream_ptr=None) -> int:
    """Phase A of a chain shard (entry-independent): start masks of the window, the exit map of every tile, and the shard's
    composed map into d_map_ptr (16 device uint64).  Returns the shard handle for chain_shard_finish."""
    import ctypes as C
    from . import _lib
    h = C.c_uint64(0)
    _lib.check(_lib.lib().acgpu_chain_shard_begin(matcher.handle, d_window_ptr, n_window, n_domain, d_map_ptr, C.byref(h), stream_ptr))
    return h.value


def chain_shard_finish(shard: int, entry: int, pos_base: int, d_pos_ptr: int, d_val_ptr, cap: int, d_total_ptr: int, stream_ptr=None):
    """Phase B: the records of the shard for the chain entering at offset `entry`; positions are window positions + pos_base."""
    from . import _lib
    _lib.check(_lib.lib().acgpu_chain_shard_finish(shard, entry, pos_base, d_pos_ptr, d_val_ptr, cap, d_total_ptr, stream_ptr))


def deal_haystacks(n_haystacks: int, world: int, rank: int) -> List[int]:
    """Indices of the haystacks rank `rank` scans (round-robin)."""
    return list(range(rank, n_haystacks, world))


def exchange_counts(my_count: int, device=None, group=None) -> Tuple[List[int], int, int]:
    """All-gather the per-rank match counts.  Returns (counts per rank, my global record offset, total)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(my_count)], 0, int(my_count)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([int(my_count)], dtype=torch.int64, device=device)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    counts = [int(p.item()) for p in parts]
    return counts, sum(counts[:rank]), sum(counts)


def match_range_shard(matcher, d_haystack_ptr: int, n_chars: int, shard: RangeShard, d_pos_ptr: int, d_val_ptr,
                      cap: int, stream_ptr=None) -> int:
    """Scan one end-position range of a haystack that is resident on this rank's GPU (the pointer addresses char 0
    of the WHOLE haystack; only [shard.read_from, shard.emit_to) is touched).  Returns the number of matches."""
    import ctypes as C
    from . import _lib
    tot = C.c_int64(0)
    _lib.check(_lib.lib().acgpu_match_device(matcher.handle, d_haystack_ptr, shard.emit_to, shard.emit_from, shard.emit_to,
                                             d_pos_ptr, d_val_ptr, cap, C.byref(tot), stream_ptr))
    return tot.value


def merge_rank_streams(streams: Sequence[Sequence]) -> list:
    """Rank-ordered concatenation = the reference's listener order."""
    out: list = []
    for s in streams:
        out.extend(s)
    return out
