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
