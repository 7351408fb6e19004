"""N>1 host logic on CPU: world_size-2 gloo process group, range shards + count exchange + rank-ordered merge
reproduce the single-stream result.  The per-shard matcher here is the oracle (this is a test of the
sharding plan and the exchange, not of the kernels); the GPU twin is tests/test_gpu_parity.py::test_range_shards."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ahocorasick_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    rng = random.Random(4242)
    kws = list({"".join(rng.choice("abc") for _ in range(rng.randint(1, 6))) for _ in range(40)})
    hay = "".join(rng.choice("abc ") for _ in range(5000))
    return kws, hay


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as ora
        kws, hay = _case()
        m = ora.Matcher("ahocorasick", kws)
        max_len = max(len(k) for k in kws)
        shard = sharding.plan_range_shards(len(hay), world, max_len)[rank]
        # the rank only sees its slice [read_from, emit_to); matches are reported by END position
        piece = hay[shard.read_from:shard.emit_to]
        mine = [(int(r["start"]) + shard.read_from, int(r["end"]) + shard.read_from) for r in m.match(piece)]
        mine = [(s, e) for s, e in mine if shard.emit_from < e <= shard.emit_to]
        counts, offset, total = sharding.exchange_counts(len(mine))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            q.put((counts, offset, total, sharding.merge_rank_streams(gathered)))
        else:
            q.put((counts, offset, total, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_range_shards_world2_gloo():
    from oracle import oracle as ora
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    kws, hay = _case()
    want = [(int(r["start"]), int(r["end"])) for r in ora.Matcher("ahocorasick", kws).match(hay)]
    merged = [r[3] for r in results if r[3] is not None][0]
    assert merged == want
    for counts, offset, total, _ in results:
        assert total == len(want) and sum(counts) == total
    assert sorted(r[1] for r in results) == [0, results[0][0][0]]


def test_plan_properties():
    for n, world, max_len in [(0, 1, 5), (7, 4, 3), (1000, 8, 12), (10**9, 8, 12), (5, 8, 12)]:
        shards = sharding.plan_range_shards(n, world, max_len)
        assert len(shards) == world and shards[0].emit_from == 0 and shards[-1].emit_to == n
        for a, b in zip(shards, shards[1:]):
            assert a.emit_to == b.emit_from
        for s in shards:
            assert 0 <= s.read_from <= s.emit_from <= s.emit_to
            assert s.emit_from - s.read_from == min(s.emit_from, max_len - 1)
        for s in shards[1:]:
            assert s.emit_from % 8 == 0 or s.emit_from == n
    assert sharding.deal_haystacks(32, 8, 3) == [3, 11, 19, 27]
    assert sorted(sum((sharding.deal_haystacks(10, 4, r) for r in range(4)), [])) == list(range(10))


# ---------------------------------------------------------------- WholeWord: word-START range shards

def _word_case():
    rng = random.Random(777)
    # sorted: set order depends on the per-process string hash seed, and value indices must agree across ranks
    kws = sorted({"".join(rng.choice("abcd") for _ in range(rng.randint(1, 7))) for _ in range(60)})
    hay = "".join(rng.choice("abcd" * 3 + " ,._") for _ in range(6000))
    return kws, hay


def _word_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as ora
        kws, hay = _word_case()
        m = ora.Matcher("wholeword", kws, n_values=len(kws))
        max_len = max(len(k) for k in kws)
        shard = sharding.plan_word_shards(len(hay), world, max_len)[rank]
        # the rank only sees [read_from, read_to): one char of look-behind, max_len + 1 chars of look-ahead;
        # it reports the words that START in [emit_from, emit_to)
        piece = hay[shard.read_from:shard.read_to]
        mine = [(int(r["start"]) + shard.read_from, int(r["end"]) + shard.read_from, int(r["value"])) for r in m.match(piece)]
        mine = [t for t in mine if shard.emit_from <= t[0] < shard.emit_to]
        counts, offset, total = sharding.exchange_counts(len(mine))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        q.put((counts, offset, total, sharding.merge_rank_streams(gathered) if rank == 0 else None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_word_shards_world2_gloo():
    from oracle import oracle as ora
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_word_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    kws, hay = _word_case()
    want = [(int(r["start"]), int(r["end"]), int(r["value"])) for r in ora.Matcher("wholeword", kws, n_values=len(kws)).match(hay)]
    merged = [r[3] for r in results if r[3] is not None][0]
    assert merged == want and len(want) > 50
    for counts, offset, total, _ in results:
        assert total == len(want) and sum(counts) == total


def test_word_shard_plan_and_semantics_many_worlds():
    """plan_word_shards covers [0, n) without gaps or overlaps for any world size, and the shard semantics (one char of
    look-behind, max_len + 1 of look-ahead, filter by word start) reproduce the single stream even when boundaries cut
    words - checked with the oracle in-process for world sizes up to more shards than words."""
    from oracle import oracle as ora
    kws, hay = _word_case()
    m = ora.Matcher("wholeword", kws)
    max_len = max(len(k) for k in kws)
    want = [(int(r["start"]), int(r["end"])) for r in m.match(hay)]
    for world in (1, 2, 3, 5, 16, 257, 3000):
        shards = sharding.plan_word_shards(len(hay), world, max_len)
        assert len(shards) == world and shards[0].emit_from == 0 and shards[-1].emit_to == len(hay)
        got = []
        for a, b in zip(shards, shards[1:]):
            assert a.emit_to == b.emit_from
        for s in shards:
            assert s.read_from == max(0, s.emit_from - 1) and s.read_to == min(len(hay), s.emit_to + max_len + 1)
            if s.emit_to == s.emit_from:
                continue
            piece = hay[s.read_from:s.read_to]
            got += [(int(r["start"]) + s.read_from, int(r["end"]) + s.read_from) for r in m.match(piece)
                    if s.emit_from <= int(r["start"]) + s.read_from < s.emit_to]
        assert got == want, world


# ---------------------------------------------------------------- Longest / Shortest: synchronisation-point shards

def _classes_for(kws, cs):
    """classes[c] == 0 iff code unit c (after the matcher's case folding) occurs in no keyword - what
    matcher.char_classes() returns on the GPU side, restated with the spec's Java folding."""
    import numpy as np
    import ac_spec as spec
    used = set()
    for k in kws:
        if k:
            used.update(spec.fold(k, cs))
    out = np.zeros(65536, np.uint16)
    for c in range(65536):
        ch = chr(c)
        f = ch if cs or 0xD800 <= c < 0xE000 else spec.fold(ch, False)
        out[c] = 1 if f in used else 0
    return out


@pytest.mark.parametrize("family", ["longest", "shortest", "ahocorasick"])
@pytest.mark.parametrize("cs", [True, False])
def test_sync_point_shards_reproduce_the_single_stream(family, cs):
    """plan_sync_shards cuts a haystack right after chars that occur in no keyword; the reference automata are in their
    root state there, so the pieces are independent haystacks.  Checked against the literal oracle: the concatenation of
    the pieces' streams (shifted by lo) is the stream of the whole haystack - Set positions and Map values, dense and
    sparse separators, any world size; and the plan declines (None) when a window holds no such char."""
    import numpy as np
    from oracle import oracle as ora
    rng = random.Random(31337 + cs)
    for it in range(40):
        alpha = rng.choice(["ab", "abc", "abAB", "abcdeXY"])
        kws = sorted({"".join(rng.choice(alpha) for _ in range(rng.randint(1, 7))) for _ in range(rng.randint(1, 30))})
        seps = rng.choice([" ", " ,.", "z"])                      # never in a keyword (also not after folding)
        p_sep = rng.choice([0.02, 0.15, 0.5])
        hay = "".join(rng.choice(seps) if rng.random() < p_sep else rng.choice(alpha) for _ in range(rng.randint(50, 3000)))
        arr = np.frombuffer(hay.encode("utf-16-le"), dtype=np.uint16)
        classes = _classes_for(kws, cs)
        m = ora.Matcher(family, kws, n_values=len(kws), case_sensitive=cs)
        want = [(int(r["start"]), int(r["end"]), int(r["value"])) for r in m.match(hay)]
        for world in (1, 2, 3, 8, 50):
            shards = sharding.plan_sync_shards(arr, world, classes, True, window=4096)
            if shards is None:
                continue
            assert shards[0].lo == 0 and shards[-1].hi == len(hay) and all(a.hi == b.lo for a, b in zip(shards, shards[1:]))
            for s in shards[1:]:
                assert s.lo == len(hay) or classes[arr[s.lo - 1]] == 0
            got = []
            for s in shards:
                got += [(int(r["start"]) + s.lo, int(r["end"]) + s.lo, int(r["value"])) for r in m.match(hay[s.lo:s.hi])]
            assert got == want, (family, cs, kws, hay, world)
    # no synchronisation point in the window / every char is a keyword char -> the plan declines
    arr = np.frombuffer(("ab" * 500).encode("utf-16-le"), dtype=np.uint16)
    assert sharding.plan_sync_shards(arr, 2, _classes_for(["a", "b"], True), True) is None
    assert sharding.plan_sync_shards(arr, 2, np.ones(65536, np.uint16), False) is None
    assert [(s.lo, s.hi) for s in sharding.plan_sync_shards(arr, 1, np.ones(65536, np.uint16), False)] == [(0, 1000)]


@pytest.mark.parametrize("cs", [True, False])
def test_sync_point_shards_wholewordlongest(cs):
    """WholeWordLongest (keywords may hold non-word chars, "as if"): synchronisation points are chars that are in no
    keyword AND are non-word chars; pieces cut there reproduce the oracle's single stream.  A keyword-free LETTER is not
    a synchronisation point (the rest of its word would become a walk start) - the plan must skip it."""
    import numpy as np
    import ac_spec as spec
    from oracle import oracle as ora
    rng = random.Random(4711 + cs)
    word_chars = ora.word_chars(0).astype(bool)
    n_cut = 0
    for it in range(60):
        alpha = rng.choice(["ab", "abc", "abAB"])
        words = sorted({"".join(rng.choice(alpha) for _ in range(rng.randint(1, 4))) for _ in range(rng.randint(1, 12))})
        kws = words + [a + " " + b for a, b in zip(words[::2], words[1::2])]         # phrases: ' ' IS a keyword char
        fill = alpha * 3 + "  " + ",;" + "z"                                            # ',' ';' sync points, 'z' a keyword-free letter
        hay = "".join(rng.choice(fill) for _ in range(rng.randint(50, 2500)))
        arr = np.frombuffer(hay.encode("utf-16-le"), dtype=np.uint16)
        classes = _classes_for(kws, cs)
        assert classes[ord("z")] == 0 and classes[ord(",")] == 0
        ok_cut = ",;" + (" " if classes[ord(" ")] == 0 else "")   # ' ' only when no phrase made it a keyword char
        m = ora.Matcher("wholewordlongest", kws, n_values=len(kws), case_sensitive=cs)
        want = [(int(r["start"]), int(r["end"]), int(r["value"])) for r in m.match(hay)]
        for world in (2, 3, 8, 40):
            shards = sharding.plan_sync_shards(arr, world, classes, True, window=4096, word_chars=word_chars)
            if shards is None:
                continue
            for s in shards[1:]:
                assert s.lo == len(hay) or hay[s.lo - 1] in ok_cut
            got = []
            for s in shards:
                got += [(int(r["start"]) + s.lo, int(r["end"]) + s.lo, int(r["value"])) for r in m.match(hay[s.lo:s.hi])]
            assert got == want, (cs, kws, hay, world)
            n_cut += 1
    assert n_cut > 100


def _chain_worker(rank, world, port, q):
    """world-size-2 gloo: every rank owns one chain shard (modelled by the oracle: the stream of the haystack that starts at
    the shard's entry, cut where the chain leaves the shard), the 16-entry maps ride in ONE all-gather, every rank
    composes the same entries."""
    import os
    import sys
    import torch
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(here)]
    from oracle import oracle as ora
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kws = ["ab", "ba", "aba", "bab", "abab"]
        n = 8192 * 4 + 99
        hay = ("ab" * n)[:n]                                  # periodic: no synchronisation point, chains never merge
        om = ora.Matcher("longest", kws)
        shards = sharding.plan_chain_shards(n, world)
        sh = shards[rank]

        def run(entry):
            """(records with chain position in [lo + entry, hi), exit offset) - the shard's semantics, by the oracle"""
            recs = om.match(hay[sh.lo + entry:], cap=n)
            out, exit_off = [], 0
            for r in recs:
                s, e = int(r["start"]) + sh.lo + entry, int(r["end"]) + sh.lo + entry
                if s >= sh.hi:
                    break
                out.append((s, e))
                exit_off = max(0, e - sh.hi)
            return out, exit_off

        mp = []
        for e in range(sharding.CHAIN_ENTRIES):
            recs, ex = run(e)
            mp.append(ex | (len(recs) << 8))
        mine = torch.tensor(mp, dtype=torch.int64)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        entries, firsts = sharding.compose_chain_maps([p.tolist() for p in parts])
        recs, _ = run(entries[rank])
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(recs)], dtype=torch.int64))
        q.put((rank, entries, firsts, int(sum(int(c.item()) for c in counts)), recs))
    finally:
        dist.destroy_process_group()


def test_chain_maps_compose_over_gloo():
    import multiprocessing as mp
    from oracle import oracle as ora
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randrange(2000)
    world = 2
    ps = [ctx.Process(target=_chain_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=60) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    kws = ["ab", "ba", "aba", "bab", "abab"]
    n = 8192 * 4 + 99
    hay = ("ab" * n)[:n]
    want = [(int(r["start"]), int(r["end"])) for r in ora.Matcher("longest", kws).match(hay, cap=n)]
    assert res[0][1] == res[1][1] and res[0][3] == res[1][3] == len(want)      # every rank composed the same picture
    got = res[0][4] + res[1][4]
    assert got == want
    assert res[1][2][1] == len(res[0][4])                                       # rank 1's first record index
