"""GPU parity at REAL config size (VERDICT r01 "Next 1"): the full BASELINE dictionaries against the CPU oracle on
>= 16 MB slices, record positions near 2^30 and 2^31 on 2*10^9-char device haystacks, a 16 MB prefix of every one of
the 32 configs[4] haystacks.  Bit-exact ordered (start, end, value) streams.

Windows of a long haystack are compared with the oracle run on the window alone:
  * AhoCorasick: the matches ending in (a, b] depend only on h[a - max_len, b)       (AhoCorasickSet.java:193-252);
  * WholeWord:   cut at non-word chars - a word is a keyword or not by itself        (WholeWordMatchSet.java:47-132);
  * Longest / Shortest: cut at SYNCHRONISATION points - a char that occurs in no keyword sends the reference automaton
    to its root, flushes the Longest queue (LongestMatchSet.java:227) and ends a pending Shortest match, so what
    follows is an independent haystack.
"""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as ora  # noqa: E402

import ahocorasick_b200 as ac  # noqa: E402
from ahocorasick_b200 import _lib  # noqa: E402
import workloads as W  # noqa: E402

SLICE = 8_000_000  # chars = 16 MB of UTF-16


def _slice(spec, n=SLICE):
    """make_haystack(spec, n) evaluated on the device (bit-identical generator, seconds faster than numpy)."""
    import torch
    return W.make_haystack_torch(spec, n, device=torch.device("cuda", 0)).cpu().numpy().view(np.uint16)


def _same(rec, want, values=True, note=""):
    assert len(rec) == len(want), (note, len(rec), len(want))
    assert np.array_equal(rec.start, want["start"]) and np.array_equal(rec.end, want["end"]), note
    if values:
        assert np.array_equal(rec.value.astype(np.int64), want["value"].astype(np.int64)), note


def test_config1_full_dictionary_16mb_slice():
    """configs[1]: AhoCorasickMap(dict, dict, false), ALL 100 000 keywords, case-insensitive, mixed-case text with
    Latin-1 / Greek / Cyrillic letters (AhoCorasickMap.java:208-336)."""
    c = W.config(1)
    kws = c["keywords"]
    assert len(kws) == 100_000
    hay = _slice(c["spec"])
    want = ora.Matcher("ahocorasick", kws, n_values=len(kws), case_sensitive=False).match(hay, cap=SLICE)
    rec = ac.AhoCorasickMap(kws, list(range(len(kws))), False).match_records(hay)
    _same(rec, want, note="config1")
    assert len(want) > 1_000_000


@pytest.mark.parametrize("family,is_map", [("longest", True), ("shortest", False), ("shortest", True), ("longest", False)])
def test_config2_full_dictionary_16mb_slice(family, is_map):
    """configs[2]: LongestMatchMap / ShortestMatchSet, ALL 100 000 nested keywords (LongestMatchSet.java:192-265,
    ShortestMatchSet.java:182-260)."""
    c = W.config(2)
    kws = c["keywords"]
    assert len(kws) == 100_000
    hay = _slice(c["spec"])
    want = ora.Matcher(family, kws, n_values=len(kws) if is_map else -1).match(hay, cap=SLICE // 2)
    cls = getattr(ac, ("Longest" if family == "longest" else "Shortest") + "Match" + ("Map" if is_map else "Set"))
    m = cls(kws, list(range(len(kws))), True) if is_map else cls(kws, True)
    _same(m.match_records(hay), want, values=is_map, note=(family, is_map))
    assert len(want) > 500_000


@pytest.mark.parametrize("family,is_map,text", [("longest", True, "config2"), ("shortest", False, "config2"), ("longest", False, "dense"),
                                                ("shortest", True, "dense")])
def test_config2_host_call_pipelines_chain_shards(family, is_map, text):
    """acgpu_match_utf16 on a Longest / Shortest haystack of several 2^23-char chunks: the host call cuts the haystack into
    chain shards (upload of chunk k + 2 next to the record download of chunk k; HostCall::run_chain) and must return the
    stream of the oracle's single loop - also on text without any synchronisation point ("dense": every chunk hands a chain
    position inside a keyword over to the next one)."""
    c = W.config(2, scale=0.05)
    kws = c["keywords"]
    n = (1 << 24) + (1 << 23) + 12_345          # three chunks, the last one not a multiple of anything
    if text == "config2":
        hay = _slice(c["spec"], n)
    else:
        rng = np.random.default_rng(77)
        hay = (rng.integers(0, 3, n) + ord("a")).astype(np.uint16)
        kws = sorted({"".join(chr(ord("a") + int(x)) for x in rng.integers(0, 3, int(rng.integers(1, 12)))) for _ in range(400)})
    want = ora.Matcher(family, kws, n_values=len(kws) if is_map else -1).match(hay, cap=n)
    cls = getattr(ac, ("Longest" if family == "longest" else "Shortest") + "Match" + ("Map" if is_map else "Set"))
    m = cls(kws, list(range(len(kws))), True) if is_map else cls(kws, True)
    _same(m.match_records(hay), want, values=is_map, note=(family, is_map, text))
    assert len(want) > 1_000_000


def test_config3_full_dictionary_16mb_slice_string_and_readable():
    """configs[3]: WholeWordMatchSet(kw, true, {'_','='}, {false,true}) with ALL 50 000 keywords via String, and the Map
    via Readable (WholeWordMatchSet.java:47-132, WholeWordMatchMap.java:55-153)."""
    c = W.config(3)
    kws = c["keywords"]
    assert len(kws) == 50_000
    hay = _slice(c["spec"])
    table = ora.word_chars(2, *c["word_chars"])
    om = ora.Matcher("wholeword", kws, n_values=len(kws), word_chars_table=table)
    want = om.match(hay, cap=SLICE // 8)
    _same(ac.WholeWordMatchSet(kws, True, *c["word_chars"]).match_records(hay), want, values=False, note="config3 set")
    gm = ac.WholeWordMatchMap(kws, list(range(len(kws))), True, *c["word_chars"])
    _same(gm.match_records(hay), want, note="config3 map")
    assert len(want) > 10_000

    class Reader:  # Readable over the array, odd read sizes
        def __init__(self, arr):
            self.arr, self.at = arr, 0

        def read(self, n):
            out = self.arr[self.at:self.at + n]
            self.at += out.size
            return out

    got = []
    gm.match(Reader(hay), lambda v: got.append(v) or True)
    assert np.array_equal(np.array(got, dtype=np.int64), om.match(hay, readable=True, cap=SLICE // 8)["value"].astype(np.int64))


def test_config4_16mb_prefix_of_each_of_the_32_haystacks():
    """configs[4]: the 1M-keyword AhoCorasickSet against the oracle on a 16 MB prefix of EVERY haystack of the
    corpus (seeds 2005 .. 2036); the oracle runs one match() per host thread."""
    c = W.config(4)
    kws = c["keywords"]
    om = ora.Matcher("ahocorasick", kws)
    gm = ac.AhoCorasickSet(kws, True)
    hays = [_slice(W.HaystackSpec("lower", 2005 + i, kws)) for i in range(32)]
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        wants = list(ex.map(lambda h: om.match(h, cap=SLICE), hays))
    total = 0
    for i, (hay, want) in enumerate(zip(hays, wants)):
        _same(gm.match_records(hay), want, values=False, note="haystack %d" % i)
        total += len(want)
    assert total > 32 * 6_000_000


# ---------------------------------------------------------------- positions near 2^30 and 2^31 (2 * 10^9-char haystacks)

N_BIG = 2_000_000_000   # configs[2] / configs[3] haystack length: one Java String at the int limit
WIN = 1_000_000
OFFSETS = (0, (1 << 30) - WIN // 2, N_BIG - WIN)   # start, across 2^30, the last window (ends at 2*10^9 < 2^31)


def _device_run(m, hay, n, is_map):
    """Whole-haystack device match: (d_pos int32 [k, 2], d_val int32 [k] or None)."""
    import torch
    lib = _lib.lib()
    tot = C.c_int64(0)
    _lib.check(lib.acgpu_match_device(m.handle, hay.data_ptr(), n, 0, n, None, None, 0, C.byref(tot), None))
    k = tot.value
    d_pos = torch.empty((max(k, 1), 2), dtype=torch.int32, device="cuda")
    d_val = torch.empty(max(k, 1), dtype=torch.int32, device="cuda") if is_map else None
    _lib.check(lib.acgpu_match_device(m.handle, hay.data_ptr(), n, 0, n, d_pos.data_ptr(), d_val.data_ptr() if is_map else None, k,
                                      C.byref(tot), None))
    torch.cuda.synchronize()
    assert tot.value == k
    return d_pos[:k], (d_val[:k] if is_map else None)


def _window_records(d_pos, d_val, lo_start, hi_end):
    """Records with start >= lo_start and end <= hi_end of an end-ascending (AhoCorasick) or start-ascending stream."""
    import torch
    end = d_pos[:, 1].contiguous()
    i0 = int(torch.searchsorted(end, torch.tensor([lo_start], dtype=torch.int32, device="cuda"), right=True))
    i1 = int(torch.searchsorted(end, torch.tensor([hi_end], dtype=torch.int32, device="cuda"), right=True))
    pos = d_pos[i0:i1].cpu().numpy().astype(np.int64)
    val = d_val[i0:i1].cpu().numpy().astype(np.int64) if d_val is not None else None
    keep = pos[:, 0] >= lo_start
    return pos[keep], (val[keep] if val is not None else None)


def _cut_points(hay, n, is_cut, a, b):
    """First cut char at or after a, last cut char before b (is_cut: bool[65536] lookup) - positions of the chars."""
    import torch
    lut = torch.from_numpy(is_cut).cuda()
    span = 4096
    w = hay[a:min(n, a + span)].to(torch.int64) & 0xFFFF
    first = a + int(torch.nonzero(lut[w])[0])
    lo = max(0, b - span)
    w = hay[lo:b].to(torch.int64) & 0xFFFF
    last = lo + int(torch.nonzero(lut[w])[-1])
    return first, last


def test_ahocorasick_positions_up_to_2e9():
    """AhoCorasickSet, 1M keywords, ONE 2*10^9-char haystack: 1.7*10^9 records; the windows at 0, across 2^30 and at the very
    end (ends up to 2 000 000 000 < 2^31) equal the oracle's stream on the window (shifted)."""
    import torch
    c = W.config(4)
    kws = c["keywords"]
    spec = W.HaystackSpec("lower", 2077, kws)
    hay = W.make_haystack_torch(spec, N_BIG, device=torch.device("cuda", 0))
    m = ac.AhoCorasickSet(kws, True)
    d_pos, _ = _device_run(m, hay, N_BIG, False)
    assert d_pos.shape[0] > 1_500_000_000
    assert int(d_pos[-1, 1]) <= N_BIG and int(d_pos[-1, 1]) > N_BIG - 64
    om = ora.Matcher("ahocorasick", kws)
    for off in OFFSETS:
        a, b = off, min(N_BIG, off + WIN)
        ctx = min(a, 16)
        w = (hay[a - ctx:b].cpu().numpy()).view(np.uint16)
        want = om.match(w, cap=WIN)
        ws, we = want["start"].astype(np.int64) + (a - ctx), want["end"].astype(np.int64) + (a - ctx)
        keep = we > a                       # matches that END inside (a, b]; their starts may reach into the context
        pos, _ = _window_records(d_pos, None, a - 64, b)
        pos = pos[pos[:, 1] > a]
        assert pos.shape[0] == int(keep.sum()) and pos.shape[0] > 500_000, off
        assert np.array_equal(pos[:, 0], ws[keep]) and np.array_equal(pos[:, 1], we[keep]), off
    # the device range entry at the same offsets (multi-GPU shards): identical records
    lib = _lib.lib()
    for off in OFFSETS[1:]:
        a, b = off, min(N_BIG, off + WIN)
        tot = C.c_int64(0)
        d_w = torch.empty((WIN * 2, 2), dtype=torch.int32, device="cuda")
        # positions are indices of a keyword's LAST char: ends in (a, b] <=> last chars in [a, b)
        _lib.check(lib.acgpu_match_device(m.handle, hay.data_ptr(), N_BIG, a, b, d_w.data_ptr(), None, WIN * 2, C.byref(tot), None))
        torch.cuda.synchronize()
        pos, _ = _window_records(d_pos, None, a - 64, b)
        pos = pos[pos[:, 1] > a]
        assert tot.value == pos.shape[0]
        assert np.array_equal(d_w[:tot.value].cpu().numpy().astype(np.int64), pos), off


@pytest.mark.parametrize("family,is_map", [("longest", True), ("shortest", False)])
def test_config2_positions_up_to_2e9(family, is_map):
    """configs[2] at FULL size: LongestMatchMap / ShortestMatchSet, 100 000 nested keywords, 2*10^9 chars (`int idx`,
    LongestMatchSet.java:197).  Windows cut at synchronisation points (spaces: in no keyword) equal the oracle's stream."""
    import torch
    c = W.config(2)
    kws = c["keywords"]
    hay = W.make_haystack_torch(c["spec"], N_BIG, device=torch.device("cuda", 0))
    cls = getattr(ac, ("Longest" if family == "longest" else "Shortest") + "Match" + ("Map" if is_map else "Set"))
    m = cls(kws, list(range(len(kws))), True) if is_map else cls(kws, True)
    classes, has_other = m.char_classes()
    assert has_other and classes[32] == 0
    d_pos, d_val = _device_run(m, hay, N_BIG, is_map)
    assert d_pos.shape[0] > 200_000_000
    start, end = d_pos[:, 0], d_pos[:, 1]
    assert bool((start[1:] >= end[:-1]).all()) and bool((end > start).all()) and int(end[-1]) <= N_BIG
    om = ora.Matcher(family, kws, n_values=len(kws) if is_map else -1)
    for off in OFFSETS:
        a, b = _cut_points(hay, N_BIG, classes == 0, off, min(N_BIG, off + WIN))
        lo = a + 1 if off else 0            # the piece after the cut char (or the haystack's start)
        w = hay[lo:b].cpu().numpy().view(np.uint16)
        want = om.match(w, cap=WIN)
        pos, val = _window_records(d_pos, d_val, lo, b)
        assert pos.shape[0] == len(want) and len(want) > 100_000, (off, pos.shape[0], len(want))
        assert np.array_equal(pos[:, 0], want["start"].astype(np.int64) + lo), off
        assert np.array_equal(pos[:, 1], want["end"].astype(np.int64) + lo), off
        if is_map:
            assert np.array_equal(val, want["value"].astype(np.int64)), off


def test_config3_positions_up_to_2e9():
    """configs[3] at FULL size: WholeWordMatchSet with the toggle word chars, 50 000 keywords, 2*10^9 chars; windows cut at
    non-word chars equal the oracle's stream."""
    import torch
    c = W.config(3)
    kws = c["keywords"]
    hay = W.make_haystack_torch(c["spec"], N_BIG, device=torch.device("cuda", 0))
    m = ac.WholeWordMatchSet(kws, True, *c["word_chars"])
    table = ora.word_chars(2, *c["word_chars"])
    d_pos, _ = _device_run(m, hay, N_BIG, False)
    assert d_pos.shape[0] > 2_000_000
    om = ora.Matcher("wholeword", kws, word_chars_table=table)
    for off in OFFSETS:
        a, b = _cut_points(hay, N_BIG, table == 0, off, min(N_BIG, off + WIN))
        lo = a + 1 if off else 0
        w = hay[lo:b].cpu().numpy().view(np.uint16)
        want = om.match(w, cap=WIN // 4)
        pos, _ = _window_records(d_pos, None, lo, b)
        assert pos.shape[0] == len(want) and len(want) > 1_000, (off, pos.shape[0], len(want))
        assert np.array_equal(pos[:, 0], want["start"].astype(np.int64) + lo), off
        assert np.array_equal(pos[:, 1], want["end"].astype(np.int64) + lo), off


@pytest.mark.parametrize("cs", [True, False])
def test_sync_point_shards_wholewordlongest_on_device(cs):
    """SURVEY 8e for WholeWordLongest on the DEVICE (the CPU twin is tests/test_sharding_gloo.py): pieces cut at chars that
    are in no keyword and are non-word chars, each scanned as a haystack of its own, concatenate to the oracle's stream."""
    import random
    import torch
    from ahocorasick_b200.sharding import match_sync_shard, plan_sync_shards
    rng = random.Random(811 + cs)
    words = sorted({"".join(rng.choice("abcAB") for _ in range(rng.randint(1, 6))) for _ in range(300)})
    kws = words + [a + " " + b for a, b in zip(words[::2], words[1::2])]     # phrases: ' ' is a keyword char
    fill = "abcAB" * 3 + "   " + ",;" + "z"                                    # ',' ';' synchronise, 'z' is a keyword-free letter
    hay = "".join(rng.choice(fill) for _ in range(600_000))
    arr = np.frombuffer(hay.encode("utf-16-le"), dtype=np.uint16)
    m = ac.WholeWordLongestMatchMap(kws, list(range(len(kws))), cs)
    classes, has_other = m.char_classes()
    want = ora.Matcher("wholewordlongest", kws, n_values=len(kws), case_sensitive=cs).match(hay, cap=1 << 20)
    want_pos = np.stack([want["start"], want["end"]], axis=1).astype(np.int64)
    d_hay = torch.from_numpy(arr.astype(np.int16)).cuda()
    cap = len(want) + 16
    d_pos = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
    d_val = torch.empty(cap, dtype=torch.int32, device="cuda")
    for world in (2, 3, 8):
        shards = plan_sync_shards(d_hay, world, classes, has_other, word_chars=m.getWordChars())
        assert shards is not None and len(shards) == world
        for s in shards[1:]:
            assert hay[s.lo - 1] in ",;"
        pos_parts, val_parts = [], []
        for sh in shards:
            k = match_sync_shard(m, d_hay.data_ptr(), sh, d_pos.data_ptr(), d_val.data_ptr(), cap)
            torch.cuda.synchronize()
            pos_parts.append(d_pos[:k].cpu().numpy().astype(np.int64) + sh.lo)
            val_parts.append(d_val[:k].cpu().numpy().astype(np.int64))
        got = np.concatenate(pos_parts, axis=0)
        assert got.shape == want_pos.shape and np.array_equal(got, want_pos), world
        assert np.array_equal(np.concatenate(val_parts), want["value"].astype(np.int64)), world
    assert len(want) > 10_000
