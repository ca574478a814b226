"""Batch partition across the GPUs of one node.

Prompts are independent, so the hot path shards with NO data-path collective: prompt b is served by rank
b mod G, every rank holds the full weights and its own HBM copy of the schema's segment store.  NCCL (over
NVLink 5 / NVSwitch) is used for two things only:
  * `broadcast_schema`  one-time broadcast of the segment store rank `src` encoded, so that `add_schema`
                        runs the scaffold forwards once per node instead of once per GPU;
  * `gather_results`    all-gather of the per-rank outputs / timings at the end of a batch.
The reference has no distributed code at all; it runs one OS process per GPU over a dataset
(eval_acc.py:240-262, scripts/run_benchmarks.py:48-62) -- this module is the in-process equivalent.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_indices(n_items: int, rank_: Optional[int] = None, world_: Optional[int] = None) -> List[int]:
    """Indices of the prompts rank `rank_` serves: b mod G == rank (round-robin keeps shards within one item)."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    return list(range(r, n_items, w))


def gather_results(local: Any) -> List[Any]:
    """All ranks' `local` objects, in rank order (a list of length 1 without a process group)."""
    if world() == 1:
        return [local]
    out: List[Any] = [None] * world()
    dist.all_gather_object(out, local)
    return out


def merge_sharded(per_rank: Sequence[Sequence[Any]], n_items: int) -> List[Any]:
    """Inverse of shard_indices: interleave the per-rank result lists back into prompt order."""
    w = len(per_rank)
    merged: List[Any] = [None] * n_items
    for r, items in enumerate(per_rank):
        for k, item in enumerate(items):
            merged[r + k * w] = item
    return merged


def max_over_ranks(value_ms: float, device=None) -> float:
    """Timing rule: a multi-GPU step takes as long as its slowest rank."""
    if world() == 1:
        return float(value_ms)
    t = torch.tensor([value_ms], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def _default_device(device):
    """NCCL moves CUDA tensors only and the splice needs HBM-resident segments: default to the current CUDA device
    (gloo test rigs on CUDA-less hosts get "cpu")."""
    if device is not None:
        return torch.device(device)
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_schema(schema_cache, src: int = 0, device=None) -> dict:
    """Make every rank hold the segment store rank `src` encoded.

    `schema_cache` is a SchemaCache built with the same schema on every rank (`no_cache=True` on ranks other
    than `src`, so only `src` ran the scaffold forwards).  The store is ONE arena [L,2,Hkv,R,D] whose row layout every
    rank derives identically from the schema (`SchemaCache.layout`), so it moves with a single ncclBroadcast (C5: 8 GiB)
    and the receivers just register row slices.  Returns {"segments", "bytes"} of what was moved."""
    if world() == 1:
        return {"segments": 0, "bytes": 0}
    device = _default_device(device if device is not None else getattr(schema_cache, "target_device", None))
    plan, rows, total = schema_cache.layout()
    seqs = {}
    for _, _, target in plan:
        for seq in target:
            seqs.setdefault(id(seq), seq)
    # shape header from the source rank; an empty store is announced instead of raising before the collective
    meta: List[Any] = [None]
    if rank() == src:
        first = next(iter(schema_cache.cache_l1.values()), None)
        if first is None:
            meta = [("empty",)]
        else:
            if schema_cache.arena is None:  # entries built one by one (tests, older stores): pack them into an arena first
                L, _, hkv, _, d = first.packed.shape
                arena = torch.empty(L, 2, hkv, max(total, 1), d, dtype=first.packed.dtype, device=first.packed.device)
                old = dict(schema_cache.cache_l1)
                schema_cache.arena = arena
                for sid, seq in seqs.items():
                    if sid in old:
                        schema_cache.adopt(seq, rows[sid]).packed.copy_(old[sid].packed)
            a = schema_cache.arena
            meta = [(int(a.shape[0]), int(a.shape[2]), int(a.shape[4]), str(a.dtype).replace("torch.", ""), int(a.shape[3]))]
    dist.broadcast_object_list(meta, src=src)
    if meta[0][0] == "empty":
        return {"segments": 0, "bytes": 0}
    L, hkv, d, dtype_name, total_rows = meta[0]
    if rank() != src:
        schema_cache.arena = torch.empty(L, 2, hkv, total_rows, d, dtype=getattr(torch, dtype_name), device=device)
    dist.broadcast(schema_cache.arena, src=src)
    if rank() != src:
        for sid, seq in seqs.items():
            schema_cache.adopt(seq, rows[sid])
    return {"segments": len(seqs), "bytes": schema_cache.arena.numel() * schema_cache.arena.element_size()}


class ServedPrompt:
    """What `serve_batch` returns per prompt (in prompt order on every rank)."""

    __slots__ = ("index", "rank", "text", "cache_ms", "prefill_ms", "n_cached", "q_len", "spliced_rows")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))

    @property
    def ttft_ms(self) -> float:
        return self.cache_ms + self.prefill_ms


def serve_batch(cache_engine, gen_engine, prompts: Sequence, params, stream_interval: int = 2) -> List[ServedPrompt]:
    """Serve a batch of independent prompts across the ranks of the process group: prompt b runs on rank b mod G
    through the same public calls a single-GPU user makes (CacheEngine.process -> GenerationEngine.generate), the
    results are all-gathered and returned in prompt order on every rank.  No data-path collective: the in-process
    equivalent of the reference's one-OS-process-per-GPU dataset partition (eval_acc.py:240-262,
    scripts/run_benchmarks.py:48-62)."""
    lm = cache_engine.lm
    full_ids = bool(getattr(lm, "use_full_position_ids", False))
    mine = []
    for b in shard_indices(len(prompts)):
        ids, pos, cache_ms, cache = cache_engine.process(prompts[b], return_full_position_ids=full_ids)
        outs = list(gen_engine.generate(ids, pos, params, cache, stream_interval=stream_interval, use_full_position_ids=full_ids))
        mine.append(dict(index=b, rank=rank(), text=outs[-1].new_text if outs else "", cache_ms=float(cache_ms),
                         prefill_ms=float(outs[0].elapsed_time) if outs else 0.0, n_cached=len(cache_engine.prompt_cache),
                         q_len=len(ids), spliced_rows=int(cache_engine.prompt_cache.last_spliced_rows)))
    merged = merge_sharded(gather_results(mine), len(prompts))
    return [ServedPrompt(**m) for m in merged]
