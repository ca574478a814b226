"""world_size-2 gloo tests (CPU) of the batch-partition host logic: sharding, result gathering, max-over-ranks
timing and the one-time broadcast of an encoded schema store."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from promptcache_b200 import parallel
from tests.helpers import HostLM

SCHEMA = ('<schema name="s"><module name="a">one two three</module><union scaffold="u1"><module name="u1">four five</module>'
          '<module name="u2">six seven eight nine</module></union><module name="b">ten</module></schema>')


def test_shard_and_merge_roundtrip():
    for n, w in [(0, 2), (1, 2), (7, 2), (64, 8), (5, 8)]:
        shards = [parallel.shard_indices(n, r, w) for r in range(w)]
        assert sorted(i for s in shards for i in s) == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
        assert parallel.merge_sharded([[f"p{i}" for i in s] for s in shards], n) == [f"p{i}" for i in range(n)]
    assert parallel.gather_results({"x": 1}) == [{"x": 1}] and parallel.max_over_ranks(3.5) == 3.5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Out:
    def __init__(self, text, ms):
        self.new_text, self.elapsed_time = text, ms


class _StubCache:
    last_spliced_rows = 0

    def __len__(self):
        return 7


class _StubEngine:
    """Duck-typed CacheEngine / GenerationEngine pair: serve_batch only touches the public calls."""

    def __init__(self, lm):
        self.lm, self.prompt_cache = lm, _StubCache()

    def process(self, prompt, return_full_position_ids=False):
        self.prompt_cache.last_spliced_rows = 10 * int(prompt[1:])
        return [1, 2], [7, 8], 1.5, prompt


class _StubGen:
    def generate(self, ids, pos, params, cache, stream_interval=2, use_full_position_ids=False):
        yield _Out(cache + "!", float(cache[1:]))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from promptcache_b200.cache_engine import SchemaCache, TokenSequenceCache
        from promptcache_b200.schema import Schema
        lm = HostLM()
        schema = Schema(SCHEMA, lm)
        sc = SchemaCache(schema, lm, no_cache=True, target_device="cpu")
        if rank == 0:  # pretend rank 0 encoded the schema: fill its store with recognisable values
            for path in sc.scaffold_paths():
                for seq in schema.get_scaffold(path).select(path).all_token_sequences():
                    t = torch.full((2, 2, 3, len(seq), 8), float(seq.offset), dtype=torch.float16)
                    t += torch.arange(len(seq)).view(1, 1, 1, -1, 1) / 16
                    sc.cache_l1[id(seq)] = TokenSequenceCache(seq, t)
        moved = parallel.broadcast_schema(sc, src=0)
        assert moved["segments"] == 4 and moved["bytes"] == sum(e.packed.numel() * 2 for e in sc.cache_l1.values())
        # an empty store on the source is announced, not raised before the collective (no deadlock)
        empty = SchemaCache(Schema(SCHEMA.replace('name="s"', 'name="e"'), lm), lm, no_cache=True, target_device="cpu")
        assert parallel.broadcast_schema(empty, src=0) == {"segments": 0, "bytes": 0} and not empty.cache_l1
        served = parallel.serve_batch(_StubEngine(lm), _StubGen(), [f"p{i}" for i in range(5)], params=None)
        assert [(s.index, s.rank, s.text) for s in served] == [(i, i % world, f"p{i}!") for i in range(5)]
        assert all(s.ttft_ms == 1.5 + s.index for s in served) and served[3].spliced_rows == 30
        store = sorted((e.token_sequence.offset, len(e), float(e.packed.sum())) for e in sc.cache_l1.values())
        mine = [f"r{rank}-p{i}" for i in parallel.shard_indices(5)]
        gathered = parallel.gather_results(mine)
        slowest = parallel.max_over_ranks(10.0 + rank)
        ret[rank] = (store, parallel.merge_sharded(gathered, 5), slowest, parallel.world(), parallel.rank())
    finally:
        dist.destroy_process_group()


def test_two_rank_broadcast_gather_and_timing():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert r0[0] == r1[0] and len(r0[0]) == 4          # both ranks hold the same 4 segments (a, u1, u2, b)
    assert [o for o, _, _ in r0[0]] == [0, 3, 3, 7]    # union members share offset 3, b sits after the longest member
    assert r0[1] == r1[1] == ["r0-p0", "r1-p1", "r0-p2", "r1-p3", "r0-p4"]
    assert r0[2] == r1[2] == 11.0                      # max over ranks
    assert (r0[3], r0[4], r1[4]) == (2, 0, 1)
